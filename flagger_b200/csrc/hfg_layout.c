/*
 * hfg_layout.c -- host side of the data layout for the CUDA E-step (plain C, like the reference's host code).
 *
 * Turns the reference's per-chunk CoverageInfo sequences (flat u16 arrays at the C-ABI, include/hfg.h) into the
 * run-constant, segment-transposed packed observation words the kernel streams.  Everything that depends only on
 * the data -- validity masks (hmm_utils.c:2229-2264), region-change flags (hmm.c:398-400), the contig-end factor
 * beta (hmm.c:301-316) -- is evaluated ONCE here instead of 3x per window per iteration as in the reference.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

#include "hfg_internal.h"

#define HFG_PACK_THREADS 16

#include <time.h>
static double lay_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * t.tv_sec + 1e-6 * t.tv_nsec;
}

#define HFG_LHD static inline
#include "hfg_layout_inl.h"

double hfg_beta(const hfg_config *cfg, const hfg_chunk_desc *ch, int i) {
    return hfg_beta_of(cfg->adjust_contig_ends, cfg->min_read_fraction_at_ends, cfg->mean_read_length, ch->ctg_len, ch->s,
                       ch->e, ch->window_len, i);
}

static uint32_t validity_mask(const hfg_config *cfg, uint16_t cov, uint16_t mapq, uint16_t clip) {
    return hfg_validity_mask(cfg->max_high_mapq_ratio, cfg->min_high_mapq_ratio, cfg->min_highly_clipped_ratio, cov, mapq, clip);
}

void hfg_layout_free(hfg_layout *l) {
    if (!l) return;
    free(l->obsT);
    free(l->wkeyT);
    free(l->kdesc);
    free(l->kbeta);
    free(l->klist);
    free(l->tile_key);
    free(l->tile_begin);
    free(l->tile_cnt);
    free(l->edge_head);
    free(l->edge_tail);
    free(l->seg_start);
    free(l->seg_len);
    free(l->seg_chunk);
    free(l->seg_edge_begin);
    free(l->edge_beta);
    free(l->chunk_offset);
    memset(l, 0, sizeof(*l));
}

/* ---- segmentation ------------------------------------------------------------------------------------------------
 * A "run" is a maximal stretch of windows of one chunk with one region index; segments are the runs cut into pieces of
 * at most smax windows.  Contig-end ("edge") windows -- beta different from the interior constant -- form a prefix and a
 * suffix of their chunk: beta(i) = (min(mid,U) - max(mid-Lr+1,Lo))/Lr is concave piecewise linear in mid, mid is
 * non-decreasing in i, so the windows that reach the plateau (Lr-1)/Lr are contiguous. */

typedef struct Run {
    int32_t chunk;
    int32_t first; /* window index inside the chunk */
    int32_t len;
} Run;

/* segment slots per statistics worker (the E-step kernel's threads per tile worker; set by hfg_api.cu before a build) */
int32_t hfg_layout_tile_div = 1;

static int64_t segments_for(const Run *runs, int64_t n_runs, int smax) {
    int64_t n = 0;
    for (int64_t i = 0; i < n_runs; i++) n += (runs[i].len + smax - 1) / smax;
    return n;
}

typedef struct PackJob {
    const hfg_config *cfg;
    const hfg_chunk_desc *chunks;
    const uint16_t *cov, *mapq, *clip;
    const uint8_t *region;
    const int32_t *edge_head, *edge_tail; /* per chunk: number of leading / trailing edge windows */
    hfg_layout *out;
    int32_t seg_begin, seg_end;
} PackJob;

/* fills the packed words (and edge factors) of segments [seg_begin, seg_end) -- independent across segments */
static void *pack_segments(void *arg) {
    PackJob *jb = arg;
    hfg_layout *out = jb->out;
    const int capacity = out->capacity;
    for (int32_t seg = jb->seg_begin; seg < jb->seg_end; seg++) {
        const int c = out->seg_chunk[seg];
        const hfg_chunk_desc *ch = &jb->chunks[c];
        const int64_t o = ch->offset;
        const int L = ch->n_windows, a = (int) (out->seg_start[seg] - o), len = out->seg_len[seg];
        int64_t e = out->seg_edge_begin[seg];
        for (int k = 0; k < len; k++) {
            const int w = a + k;
            const int64_t g = o + w;
            const int is_edge = w < jb->edge_head[c] || w >= L - jb->edge_tail[c];
            const uint32_t word = hfg_pack_word(validity_mask(jb->cfg, jb->cov[g], jb->mapq[g], jb->clip[g]), jb->cov[g],
                                                w > 0 ? jb->cov[g - 1] : 0, jb->region[g], w > 0 ? jb->region[g - 1] : 0, w,
                                                L, is_edge);
            if (is_edge) {
                /* (beta, beta0/beta, sqrt(beta0/beta)) */
                const double b = hfg_beta(jb->cfg, ch, w);
                out->edge_beta[3 * e] = b;
                out->edge_beta[3 * e + 1] = out->beta0 / b;
                out->edge_beta[3 * e + 2] = sqrt(out->beta0 / b);
                e++;
            }
            out->obsT[(size_t) k * capacity + seg] = word;
        }
    }
    return NULL;
}


/* ---- observation keys ----------------------------------------------------------------------------------------------
 * Everything the kernel needs to know about a window except its position: the packed word without the chunk-end bit,
 * plus the bits of beta for contig-end windows.  Distinct combinations are numbered through an open-addressing hash
 * table (a 3 Gbp assembly at 40x has ~10^4 of them for ~10^6 windows). */

typedef struct KeySlot {
    uint64_t beta_bits;
    uint32_t word;
    int32_t id; /* -1 = empty */
} KeySlot;

typedef struct KeyTab {
    KeySlot *slots;
    uint32_t mask; /* capacity - 1 */
    int32_t n;
    /* per key, in discovery order */
    uint32_t *word;
    uint64_t *beta_bits;
    int32_t *count;
    int32_t *first; /* smallest global window index carrying the key */
    int32_t cap_keys;
} KeyTab;

static uint32_t key_hash(uint32_t word, uint64_t bb) {
    uint64_t h = ((uint64_t) word * 0x9E3779B97F4A7C15ull) ^ (bb * 0xC2B2AE3D27D4EB4Full);
    return (uint32_t) (h >> 32);
}

static int keytab_init(KeyTab *t, uint32_t capacity) {
    memset(t, 0, sizeof(*t));
    t->slots = malloc(sizeof(KeySlot) * (size_t) capacity);
    t->cap_keys = (int32_t) (capacity / 2);
    t->word = malloc(sizeof(uint32_t) * (size_t) t->cap_keys);
    t->beta_bits = malloc(sizeof(uint64_t) * (size_t) t->cap_keys);
    t->count = malloc(sizeof(int32_t) * (size_t) t->cap_keys);
    t->first = malloc(sizeof(int32_t) * (size_t) t->cap_keys);
    if (!t->slots || !t->word || !t->beta_bits || !t->count || !t->first) return 0;
    for (uint32_t i = 0; i < capacity; i++) t->slots[i].id = -1;
    t->mask = capacity - 1;
    return 1;
}

static void keytab_free(KeyTab *t) {
    free(t->slots);
    free(t->word);
    free(t->beta_bits);
    free(t->count);
    free(t->first);
    memset(t, 0, sizeof(*t));
}

/* doubles the table (keys keep their ids) */
static int keytab_grow(KeyTab *t) {
    const uint32_t capacity = (t->mask + 1) * 2;
    KeySlot *ns = malloc(sizeof(KeySlot) * (size_t) capacity);
    uint32_t *nw = realloc(t->word, sizeof(uint32_t) * (size_t) (capacity / 2));
    if (nw) t->word = nw;
    uint64_t *nb = realloc(t->beta_bits, sizeof(uint64_t) * (size_t) (capacity / 2));
    if (nb) t->beta_bits = nb;
    int32_t *nc = realloc(t->count, sizeof(int32_t) * (size_t) (capacity / 2));
    if (nc) t->count = nc;
    int32_t *nf = realloc(t->first, sizeof(int32_t) * (size_t) (capacity / 2));
    if (nf) t->first = nf;
    if (!ns || !nw || !nb || !nc || !nf) {
        free(ns);
        return 0;
    }
    for (uint32_t i = 0; i < capacity; i++) ns[i].id = -1;
    for (int32_t k = 0; k < t->n; k++) {
        uint32_t h = key_hash(t->word[k], t->beta_bits[k]) & (capacity - 1);
        while (ns[h].id >= 0) h = (h + 1) & (capacity - 1);
        ns[h].word = t->word[k];
        ns[h].beta_bits = t->beta_bits[k];
        ns[h].id = k;
    }
    free(t->slots);
    t->slots = ns;
    t->mask = capacity - 1;
    t->cap_keys = (int32_t) (capacity / 2);
    return 1;
}

/* id of (word, beta_bits) seen at global window g, inserting it when new; -1 when out of memory */
static inline int32_t keytab_lookup(KeyTab *t, uint32_t word, uint64_t bb, int32_t g) {
    uint32_t h = key_hash(word, bb) & t->mask;
    for (;;) {
        KeySlot *s = &t->slots[h];
        if (s->id < 0) break;
        if (s->word == word && s->beta_bits == bb) {
            t->count[s->id]++;
            return s->id;
        }
        h = (h + 1) & t->mask;
    }
    if (t->n == t->cap_keys) {
        if (!keytab_grow(t)) return -1;
        h = key_hash(word, bb) & t->mask;
        while (t->slots[h].id >= 0) h = (h + 1) & t->mask;
    }
    const int32_t id = t->n++;
    t->slots[h].word = word;
    t->slots[h].beta_bits = bb;
    t->slots[h].id = id;
    t->word[id] = word;
    t->beta_bits[id] = bb;
    t->count[id] = 1;
    t->first[id] = g;
    return id;
}

/* Final key order: by region, then hottest first (their matrices share cache lines), ties by the first window that
 * carries the key -- a function of the data alone, whatever the slicing.  One 64-bit sort key per table entry
 * (region 6 bits | 2^28-1-count 28 bits | first window 28 bits), LSD radix sort with the old id as payload. */
typedef struct KeyOrder {
    uint64_t sort_key;
    int32_t old_id, count;
} KeyOrder;

static int key_order_sort(KeyOrder *a, int32_t n) {
    KeyOrder *tmp = malloc(sizeof(KeyOrder) * (size_t) (n > 0 ? n : 1));
    if (!tmp) return 0;
    KeyOrder *src = a, *dst = tmp;
    for (int shift = 0; shift < 64; shift += 8) {
        size_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (int32_t i = 0; i < n; i++) hist[((src[i].sort_key >> shift) & 0xff) + 1]++;
        if (hist[((src[0].sort_key >> shift) & 0xff) + 1] == (size_t) n) continue; /* all equal in this digit */
        for (int d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (int32_t i = 0; i < n; i++) dst[hist[(src[i].sort_key >> shift) & 0xff]++] = src[i];
        KeyOrder *sw = src;
        src = dst;
        dst = sw;
    }
    if (src != a) memcpy(a, src, sizeof(KeyOrder) * (size_t) n);
    free(tmp);
    return 1;
}

/* does the pair (previous window -> this window) enter the statistics?  (hmm.c:638-642: pairs 1->2 .. L-2->L-1) */
static inline int key_has_stats(uint32_t word) { return !(word & (HFG_OBS_CHUNK_START | HFG_OBS_SECOND)); }

/* id of (word, beta_bits) with `cnt` more windows, inserting it when new */
static int32_t keytab_add(KeyTab *t, uint32_t word, uint64_t bb, int32_t cnt, int32_t first) {
    const int32_t id = keytab_lookup(t, word, bb, first);
    if (id >= 0) {
        t->count[id] += cnt - 1;
        if (first < t->first[id]) t->first[id] = first;
    }
    return id;
}

/* one slice of consecutive segments (= consecutive windows) per thread */
typedef struct KeyJob {
    hfg_layout *out;
    int32_t seg_begin, seg_end;
    int32_t *wkid;     /* [W] local key id of every window of the slice (shared array, disjoint ranges) */
    KeyTab kt;         /* the slice's own table */
    int32_t *final_id; /* [kt.n] local id -> final key id */
    int32_t *fill;     /* [kt.n] local id -> next free entry of klist for this slice's windows of the key */
    int ok;
} KeyJob;

/* pass 1: number the keys of the slice in discovery order */
static void *keys_number_slice(void *arg) {
    KeyJob *jb = arg;
    hfg_layout *out = jb->out;
    const int capacity = out->capacity;
    jb->ok = keytab_init(&jb->kt, 1u << 14);
    if (!jb->ok) return NULL;
    for (int32_t seg = jb->seg_begin; seg < jb->seg_end; seg++) {
        int64_t e = out->seg_edge_begin[seg];
        const int64_t g0 = out->seg_start[seg];
        for (int k = 0; k < out->seg_len[seg]; k++) {
            const uint32_t word = out->obsT[(size_t) k * capacity + seg] & ~HFG_OBS_CHUNK_END;
            uint64_t bb = 0;
            if (word & HFG_OBS_EDGE) memcpy(&bb, &out->edge_beta[3 * e++], sizeof(bb));
            const int32_t id = keytab_lookup(&jb->kt, word, bb, (int32_t) (g0 + k));
            if (id < 0) {
                jb->ok = 0;
                return NULL;
            }
            jb->wkid[g0 + k] = id;
        }
    }
    return NULL;
}

/* pass 4: key words (segment-transposed) and the per-key window lists of the slice (ascending: windows in order) */
static void *keys_scatter_slice(void *arg) {
    KeyJob *jb = arg;
    hfg_layout *out = jb->out;
    const int capacity = out->capacity;
    for (int32_t seg = jb->seg_begin; seg < jb->seg_end; seg++) {
        const int64_t g0 = out->seg_start[seg];
        for (int k = 0; k < out->seg_len[seg]; k++) {
            const uint32_t word = out->obsT[(size_t) k * capacity + seg];
            const int32_t lid = jb->wkid[g0 + k];
            uint32_t kw = (uint32_t) jb->final_id[lid];
            if (word & HFG_OBS_CHUNK_START) kw |= HFG_KEY_CHUNK_START;
            if (word & HFG_OBS_CHUNK_END) kw |= HFG_KEY_CHUNK_END;
            out->wkeyT[(size_t) k * capacity + seg] = kw;
            if (key_has_stats(word)) out->klist[jb->fill[lid]++] = (int32_t) (g0 + k);
        }
    }
    return NULL;
}

static void run_key_jobs(KeyJob *jobs, int n, void *(*fn)(void *)) {
    pthread_t tids[HFG_PACK_THREADS];
    for (int t = 1; t < n; t++)
        if (pthread_create(&tids[t], NULL, fn, &jobs[t]) != 0) {
            fn(&jobs[t]); /* could not spawn: do the slice here */
            tids[t] = 0;
        }
    fn(&jobs[0]);
    for (int t = 1; t < n; t++)
        if (tids[t]) pthread_join(tids[t], NULL);
}

/* builds wkeyT, the key tables, the per-key window lists and their tiles from obsT / edge_beta */
static int build_keys(hfg_layout *out) {
    const int64_t W = out->n_windows;
    const int capacity = out->capacity;
    int n_jobs = (int) (W / 65536) + 1;
    if (n_jobs > HFG_PACK_THREADS) n_jobs = HFG_PACK_THREADS;
    if (n_jobs > out->n_seg) n_jobs = out->n_seg > 0 ? out->n_seg : 1;
    KeyJob jobs[HFG_PACK_THREADS];
    KeyTab kt;
    memset(&kt, 0, sizeof(kt));
    memset(jobs, 0, sizeof(jobs));
    int32_t *wkid = malloc(sizeof(int32_t) * (size_t) W);
    KeyOrder *order = NULL;
    int32_t *new_id = NULL, *kbegin = NULL, *running = NULL;
    int32_t P = 0;
    int64_t n_list = 0, n_tiles = 0;
    int ok = 0;
    if (!wkid) goto done;
    double tk[6];
    tk[0] = lay_ms();
    /* pass 1 (parallel): per-slice key tables */
    for (int t = 0; t < n_jobs; t++) {
        jobs[t].out = out;
        jobs[t].seg_begin = (int32_t) ((int64_t) out->n_seg * t / n_jobs);
        jobs[t].seg_end = (int32_t) ((int64_t) out->n_seg * (t + 1) / n_jobs);
        jobs[t].wkid = wkid;
    }
    run_key_jobs(jobs, n_jobs, keys_number_slice);
    for (int t = 0; t < n_jobs; t++)
        if (!jobs[t].ok) goto done;
    tk[1] = lay_ms();
    /* merge the slice tables (O(#keys) each) */
    if (!keytab_init(&kt, 1u << 15)) goto done;
    for (int t = 0; t < n_jobs; t++) {
        KeyTab *lt = &jobs[t].kt;
        jobs[t].final_id = malloc(sizeof(int32_t) * (size_t) (lt->n > 0 ? lt->n : 1));
        jobs[t].fill = malloc(sizeof(int32_t) * (size_t) (lt->n > 0 ? lt->n : 1));
        if (!jobs[t].final_id || !jobs[t].fill) goto done;
        for (int32_t k = 0; k < lt->n; k++) {
            const int32_t gid = keytab_add(&kt, lt->word[k], lt->beta_bits[k], lt->count[k], lt->first[k]);
            if (gid < 0) goto done;
            jobs[t].final_id[k] = gid; /* merged id for now */
        }
    }
    P = kt.n;
    if (P > HFG_KEY_MAX) goto done;
    tk[2] = lay_ms();
    /* pass 2: final numbering by (region, count descending) */
    order = malloc(sizeof(KeyOrder) * (size_t) P);
    new_id = malloc(sizeof(int32_t) * (size_t) P);
    kbegin = malloc(sizeof(int32_t) * ((size_t) P + 1));
    running = malloc(sizeof(int32_t) * (size_t) P);
    out->kdesc = malloc(sizeof(uint32_t) * (size_t) P);
    out->kbeta = malloc(sizeof(double) * 3 * (size_t) P);
    if (!order || !new_id || !kbegin || !running || !out->kdesc || !out->kbeta) goto done;
    for (int32_t k = 0; k < P; k++)
        order[k] = (KeyOrder){((uint64_t) HFG_OBS_REGION(kt.word[k]) << 56) | ((uint64_t) (HFG_KEY_MAX - kt.count[k]) << 28) |
                                  (uint64_t) kt.first[k],
                              k, kt.count[k]};
    if (!key_order_sort(order, P)) goto done;
    for (int32_t p = 0; p < P; p++) {
        const int32_t o = order[p].old_id;
        new_id[o] = p;
        out->kdesc[p] = kt.word[o];
        double b = out->beta0;
        if (kt.word[o] & HFG_OBS_EDGE) memcpy(&b, &kt.beta_bits[o], sizeof(b));
        out->kbeta[3 * p] = b;
        out->kbeta[3 * p + 1] = (kt.word[o] & HFG_OBS_EDGE) ? out->beta0 / b : 1.0;
        out->kbeta[3 * p + 2] = (kt.word[o] & HFG_OBS_EDGE) ? sqrt(out->beta0 / b) : 1.0;
        kbegin[p] = (int32_t) n_list;
        if (key_has_stats(kt.word[o])) n_list += order[p].count;
    }
    /* tile length: HFG_TILE, or the shortest one for which every thread of the grid gets at most ONE tile (a block whose
     * share is a few tiles over its thread count would spend a second round on them) */
    int tile_len = HFG_TILE;
    const int tile_workers = capacity / (hfg_layout_tile_div > 0 ? hfg_layout_tile_div : 1);
    for (;; tile_len++) {
        n_tiles = 0;
        for (int32_t p = 0; p < P; p++)
            if (key_has_stats(out->kdesc[p])) n_tiles += (order[p].count + tile_len - 1) / tile_len;
        if (n_tiles <= (int64_t) tile_workers - tile_workers / 16 || tile_len >= 4 * HFG_TILE) break;
    }
    kbegin[P] = (int32_t) n_list;
    out->n_keys = P;
    out->n_list = n_list;
    out->n_tiles = (int32_t) n_tiles;
    out->tile_len = tile_len;
    out->klist = malloc(sizeof(int32_t) * (size_t) (n_list > 0 ? n_list : 1));
    out->tile_key = malloc(sizeof(int32_t) * (size_t) (n_tiles > 0 ? n_tiles : 1));
    out->tile_begin = malloc(sizeof(int32_t) * (size_t) (n_tiles > 0 ? n_tiles : 1));
    out->tile_cnt = malloc(sizeof(int32_t) * (size_t) (n_tiles > 0 ? n_tiles : 1));
    tk[3] = lay_ms();
    out->wkeyT = calloc((size_t) out->smax * capacity, sizeof(uint32_t));
    if (!out->klist || !out->tile_key || !out->tile_begin || !out->tile_cnt || !out->wkeyT) goto done;
    /* pass 3: tiles, grouped by region because the keys are */
    {
        int32_t t = 0;
        for (int r = 0; r <= HFG_MAX_REGIONS; r++) out->region_tile_begin[r] = -1;
        for (int32_t p = 0; p < P; p++) {
            const int r = (int) HFG_OBS_REGION(out->kdesc[p]);
            if (out->region_tile_begin[r] < 0) out->region_tile_begin[r] = t;
            for (int32_t b = kbegin[p]; b < kbegin[p + 1]; b += tile_len, t++) {
                out->tile_key[t] = p;
                out->tile_begin[t] = b;
                out->tile_cnt[t] = kbegin[p + 1] - b < tile_len ? kbegin[p + 1] - b : tile_len;
            }
        }
        out->region_tile_begin[HFG_MAX_REGIONS] = t;
        for (int r = HFG_MAX_REGIONS - 1; r >= 0; r--)
            if (out->region_tile_begin[r] < 0) out->region_tile_begin[r] = out->region_tile_begin[r + 1];
    }
    /* where each slice's windows of a key go in the key's list: the slices are consecutive stretches of the genome, so
     * slice t's windows follow those of the slices before it */
    memcpy(running, kbegin, sizeof(int32_t) * (size_t) P);
    for (int t = 0; t < n_jobs; t++) {
        KeyTab *lt = &jobs[t].kt;
        for (int32_t k = 0; k < lt->n; k++) {
            const int32_t p = new_id[jobs[t].final_id[k]];
            jobs[t].final_id[k] = p;
            jobs[t].fill[k] = running[p];
            if (key_has_stats(lt->word[k])) running[p] += lt->count[k];
        }
    }
    tk[4] = lay_ms();
    /* pass 4 (parallel): key words and lists */
    run_key_jobs(jobs, n_jobs, keys_scatter_slice);
    tk[5] = lay_ms();
    if (getenv("HFG_TIMING"))
        fprintf(stderr, "[hfg] keys: number %.2f ms (%d slices), merge %.2f, sort+tables %.2f, alloc+tiles+offsets %.2f, scatter %.2f\n",
                tk[1] - tk[0], n_jobs, tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4]);
    ok = 1;
done:
    for (int t = 0; t < n_jobs; t++) {
        keytab_free(&jobs[t].kt);
        free(jobs[t].final_id);
        free(jobs[t].fill);
    }
    keytab_free(&kt);
    free(wkid);
    free(order);
    free(new_id);
    free(kbegin);
    free(running);
    return ok;
}

/* number of edge windows among the first w windows of chunk c */
static int64_t edges_before(int head, int tail, int L, int w) {
    int64_t n = w < head ? w : head;
    if (w > L - tail) n += w - (L - tail);
    return n;
}

int hfg_layout_build(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                     const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                     int32_t capacity, int32_t granule, hfg_layout *out, char *err, size_t errlen) {
    return hfg_layout_build_ex(cfg, n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region, capacity, granule, 0, out, err,
                               errlen);
}

/* segments_only != 0: stop after the segment table and the per-chunk contig-end extents (the device builds the packed
 * words, the keys, their lists and tiles itself: hfg_layout_dev.cuh); only `region` is read then. */
int hfg_layout_build_ex(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                        const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                        int32_t capacity, int32_t granule, int segments_only, hfg_layout *out, char *err, size_t errlen) {
    memset(out, 0, sizeof(*out));
    const double t_lay0 = lay_ms();
    int64_t W = 0;
    for (int32_t c = 0; c < n_chunks; c++) {
        if (chunks[c].n_windows <= 0 || chunks[c].offset != W || chunks[c].window_len <= 0) {
            snprintf(err, errlen, "chunk %d: n_windows must be > 0 and offsets contiguous in list order", c);
            return HFG_ERR_INVALID;
        }
        W += chunks[c].n_windows;
    }
    if (W <= 0 || W >= 0x7fffffffLL) {
        snprintf(err, errlen, "number of windows (%lld) out of range", (long long) W);
        return HFG_ERR_INVALID;
    }
    const double beta0 = !cfg->adjust_contig_ends ? 1.0
                         : (cfg->mean_read_length > 0 ? (double) (cfg->mean_read_length - 1) / cfg->mean_read_length : 0.25);

    /* pass 1: runs (one byte scan over the region indices) and the edge prefix / suffix of every chunk */
    int64_t run_cap = 2 * (int64_t) n_chunks + 1024, n_runs = 0, n_edge = 0;
    Run *runs = malloc(sizeof(Run) * (size_t) run_cap);
    int32_t *edge_head = malloc(sizeof(int32_t) * (size_t) n_chunks), *edge_tail = malloc(sizeof(int32_t) * (size_t) n_chunks);
    int64_t *chunk_edge_base = malloc(sizeof(int64_t) * ((size_t) n_chunks + 1));
    if (!runs || !edge_head || !edge_tail || !chunk_edge_base) goto nomem;
    for (int32_t c = 0; c < n_chunks; c++) {
        const uint8_t *r = region + chunks[c].offset;
        const int L = chunks[c].n_windows;
        for (int i = 0; i < L;) {
            if (r[i] >= cfg->n_regions) {
                snprintf(err, errlen, "window %lld has region %d >= n_regions %d", (long long) (chunks[c].offset + i), r[i],
                         cfg->n_regions);
                free(runs); free(edge_head); free(edge_tail); free(chunk_edge_base);
                return HFG_ERR_INVALID;
            }
            int j = i + 1;
            {
                /* end of the run of equal region indices, eight windows per comparison */
                uint64_t pat = r[i];
                pat |= pat << 8;
                pat |= pat << 16;
                pat |= pat << 32;
                while (j + 8 <= L) {
                    uint64_t v;
                    memcpy(&v, r + j, 8);
                    if (v != pat) break;
                    j += 8;
                }
                while (j < L && r[j] == r[i]) j++;
            }
            if (n_runs == run_cap) {
                run_cap *= 2;
                Run *nr = realloc(runs, sizeof(Run) * (size_t) run_cap);
                if (!nr) goto nomem;
                runs = nr;
            }
            runs[n_runs].chunk = c;
            runs[n_runs].first = i;
            runs[n_runs].len = j - i;
            n_runs++;
            i = j;
        }
        int head = 0, tail = 0;
        while (head < L) {
            const double b = hfg_beta(cfg, &chunks[c], head);
            if (memcmp(&b, &beta0, sizeof(double)) == 0) break;
            head++;
        }
        while (tail < L - head) {
            const double b = hfg_beta(cfg, &chunks[c], L - 1 - tail);
            if (memcmp(&b, &beta0, sizeof(double)) == 0) break;
            tail++;
        }
        edge_head[c] = head;
        edge_tail[c] = tail;
        chunk_edge_base[c] = n_edge;
        n_edge += head + tail;
    }
    chunk_edge_base[n_chunks] = n_edge;

    /* smallest smax whose segment count fits the persistent grid */
    int smax = (int) ((W + capacity - 1) / capacity);
    if (smax < 1) smax = 1;
    if (n_runs > capacity) {
        /* a segment never straddles a chunk boundary or a region change, so there are at least n_runs of them whatever
         * their length: no smax can fit */
        free(runs); free(edge_head); free(edge_tail); free(chunk_edge_base);
        hfg_layout_free(out);
        snprintf(err, errlen, "%lld runs of equal region index (chunk boundaries included) exceed the %d segment slots of the grid",
                 (long long) n_runs, capacity);
        return HFG_ERR_INVALID;
    }
    /* (one window at a time while segments are short: a step of two windows at ten leaves a sixth of the SMs without work) */
    while (segments_for(runs, n_runs, smax) > capacity) smax += smax < 64 ? 1 : (smax + 7) / 8;
    const int64_t n_seg = segments_for(runs, n_runs, smax);
    /* keep only as many slots (whole CTAs of `granule` threads) as the segments of this length need: a grid padded with
     * idle CTAs only makes the grid-wide barriers slower */
    if (granule > 0) capacity = (int32_t) ((n_seg + granule - 1) / granule) * granule;

    out->n_windows = W;
    out->n_chunks = n_chunks;
    out->capacity = capacity;
    out->smax = smax;
    out->n_seg = (int32_t) n_seg;
    out->beta0 = beta0;
    out->n_edge = n_edge;
    out->obsT = segments_only ? NULL : calloc((size_t) smax * capacity, sizeof(uint32_t));
    out->seg_start = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_len = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_chunk = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_edge_begin = calloc((size_t) capacity + 1, sizeof(int32_t));
    out->chunk_offset = calloc((size_t) n_chunks + 1, sizeof(int64_t));
    out->edge_beta = malloc(sizeof(double) * 3 * (size_t) (n_edge > 0 ? n_edge : 1));
    if ((!segments_only && !out->obsT) || !out->seg_start || !out->seg_len || !out->seg_chunk || !out->seg_edge_begin ||
        !out->edge_beta || !out->chunk_offset)
        goto nomem;

    /* pass 2: the segment table (O(#segments)) */
    {
        int32_t seg = 0;
        for (int64_t i = 0; i < n_runs; i++) {
            const int c = runs[i].chunk, L = chunks[c].n_windows;
            for (int a = 0; a < runs[i].len; a += smax, seg++) {
                const int w = runs[i].first + a;
                out->seg_start[seg] = (int32_t) (chunks[c].offset + w);
                out->seg_len[seg] = runs[i].len - a < smax ? runs[i].len - a : smax;
                out->seg_chunk[seg] = c;
                out->seg_edge_begin[seg] = (int32_t) (chunk_edge_base[c] + edges_before(edge_head[c], edge_tail[c], L, w));
            }
        }
        for (int32_t j = seg; j <= capacity; j++) out->seg_edge_begin[j] = (int32_t) n_edge;
    }
    for (int32_t c = 0; c <= n_chunks; c++) out->chunk_offset[c] = c < n_chunks ? chunks[c].offset : W;

    if (W > HFG_KEY_MAX) {
        free(runs); free(edge_head); free(edge_tail); free(chunk_edge_base);
        hfg_layout_free(out);
        snprintf(err, errlen, "number of windows (%lld) above the %d this build addresses per device", (long long) W, HFG_KEY_MAX);
        return HFG_ERR_INVALID;
    }
    if (segments_only) {
        out->edge_head = edge_head; /* ownership moves to the layout */
        out->edge_tail = edge_tail;
        free(runs); free(chunk_edge_base);
        if (getenv("HFG_TIMING")) fprintf(stderr, "[hfg] layout: segments %.2f ms (host); keys on the device\n", lay_ms() - t_lay0);
        return HFG_OK;
    }
    /* pass 3: pack the observation words, in parallel over segments (pthreads, as the reference's own parser) */
    {
        int n_threads = (int) (W / 65536) + 1;
        if (n_threads > HFG_PACK_THREADS) n_threads = HFG_PACK_THREADS;
        PackJob jobs[HFG_PACK_THREADS];
        pthread_t tids[HFG_PACK_THREADS];
        for (int t = 0; t < n_threads; t++) {
            jobs[t] = (PackJob){cfg, chunks, cov, cov_high_mapq, cov_high_clip, region, edge_head, edge_tail, out,
                                (int32_t) (n_seg * t / n_threads), (int32_t) (n_seg * (t + 1) / n_threads)};
            if (t > 0 && pthread_create(&tids[t], NULL, pack_segments, &jobs[t]) != 0) {
                pack_segments(&jobs[t]); /* could not spawn: do the slice here */
                tids[t] = 0;
            }
        }
        pack_segments(&jobs[0]);
        for (int t = 1; t < n_threads; t++)
            if (tids[t]) pthread_join(tids[t], NULL);
    }
    const double t_keys0 = lay_ms();
    if (!build_keys(out)) goto nomem;
    if (getenv("HFG_TIMING"))
        fprintf(stderr, "[hfg] layout: segments+pack %.2f ms, keys %.2f ms (%d keys, %d tiles)\n", t_keys0 - t_lay0,
                lay_ms() - t_keys0, out->n_keys, out->n_tiles);
    free(runs); free(edge_head); free(edge_tail); free(chunk_edge_base);
    return HFG_OK;
nomem:
    free(runs); free(edge_head); free(edge_tail); free(chunk_edge_base);
    hfg_layout_free(out);
    snprintf(err, errlen, "out of host memory building the layout");
    return HFG_ERR_NOMEM;
}

void hfg_classes_build(const hfg_config *cfg, const double *alpha, hfg_classes *out) {
    memset(out, 0, sizeof(*out));
    int d = 0;
    for (int s = 0; s < HFG_NS; s++) {
        out->is_gaussian[s] = !(cfg->model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && s == HFG_STATE_ERR);
        out->first_class_of_state[s] = d;
        /* slot s (0..3) must be the alpha == 0 class of state s; so lay the four alpha-0 classes out first */
        out->class_state[d] = s;
        out->class_alpha[d] = 0.0;
        d++;
    }
    for (int s = 0; s < HFG_NS; s++) {
        out->n_class_of_state[s] = 1;
        for (int pre = 0; pre < HFG_NS; pre++) {
            const double a = out->is_gaussian[s] ? alpha[pre * HFG_NS + s] : 0.0; /* TruncExp ignores alpha */
            int slot = -1;
            if (a == 0.0) slot = s;
            for (int k = HFG_NS; k < d && slot < 0; k++)
                if (out->class_state[k] == s && out->class_alpha[k] == a) slot = k;
            if (slot < 0) {
                slot = d++;
                out->class_state[slot] = s;
                out->class_alpha[slot] = a;
                out->n_class_of_state[s]++;
            }
            out->cls[pre][s] = slot;
        }
    }
    out->n_classes = d;
}

/* Host-only self-check of the layout builder (no GPU): rebuilds the layout for `capacity` segment slots and verifies
 * that every window appears exactly once, in order, with the right packed fields, and that no segment straddles a chunk
 * or a region change.  summary = {n_seg, smax, n_edge, windows}.  Returns HFG_OK or HFG_ERR_INVALID. */
int hfg_debug_layout_check(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                           const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                           int32_t capacity, int64_t *summary) {
    hfg_layout l;
    char err[256];
    int rc = hfg_layout_build(cfg, n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region, capacity, 0, &l, err,
                              sizeof(err));
    if (rc != HFG_OK) return rc;
    int64_t next = 0, edges = 0;
    int bad = 0;
    for (int32_t j = 0; j < l.n_seg && !bad; j++) {
        const int c = l.seg_chunk[j];
        if (l.seg_start[j] != next || l.seg_len[j] < 1 || l.seg_len[j] > l.smax) bad = 1;
        if (l.seg_edge_begin[j] != edges) bad = 1;
        for (int k = 0; k < l.seg_len[j] && !bad; k++) {
            const int64_t g = l.seg_start[j] + k;
            const int w = (int) (g - chunks[c].offset);
            const uint32_t word = l.obsT[(size_t) k * capacity + j];
            if (w < 0 || w >= chunks[c].n_windows) bad = 1;
            if (!(word & HFG_OBS_VALID) || HFG_OBS_X(word) != (uint8_t) cov[g] || HFG_OBS_REGION(word) != region[g]) bad = 1;
            if (HFG_OBS_REGION(word) != region[l.seg_start[j]]) bad = 1; /* one region per segment */
            if (HFG_OBS_PX(word) != (w > 0 ? (uint8_t) cov[g - 1] : 0)) bad = 1;
            if (((word & HFG_OBS_CHUNK_START) != 0) != (w == 0) || (w == 0 && k != 0)) bad = 1;
            if (((word & HFG_OBS_SECOND) != 0) != (w == 1)) bad = 1;
            if (((word & HFG_OBS_CHUNK_END) != 0) != (w == chunks[c].n_windows - 1)) bad = 1;
            if (((word & HFG_OBS_REGION_CHANGE) != 0) != (w > 0 && region[g] != region[g - 1])) bad = 1;
            if ((word & HFG_OBS_REGION_CHANGE) && k != 0) bad = 1; /* a region change starts a segment */
            if (HFG_OBS_MASK(word) != validity_mask(cfg, cov[g], cov_high_mapq[g], cov_high_clip[g])) bad = 1;
            if (word & HFG_OBS_EDGE) {
                const double b = hfg_beta(cfg, &chunks[c], w);
                if (l.edge_beta[3 * edges] != b || b == l.beta0) bad = 1;
                edges++;
            } else if (hfg_beta(cfg, &chunks[c], w) != l.beta0) bad = 1;
        }
        for (int k = l.seg_len[j]; k < l.smax && !bad; k++)
            if (l.obsT[(size_t) k * capacity + j] != 0) bad = 1; /* padding */
        next += l.seg_len[j];
    }
    for (int32_t j = l.n_seg; j < capacity && !bad; j++)
        if (l.seg_len[j] != 0) bad = 1;
    if (next != l.n_windows || edges != l.n_edge || l.n_seg > capacity) bad = 1;
    /* keys: every window's key reproduces its packed word and beta; every pair that enters the statistics is listed
     * exactly once under its key, ascending; the tiles partition the lists and are grouped by region */
    if (!bad) {
        int64_t e = 0, listed = 0;
        uint8_t *seen = calloc((size_t) l.n_windows, 1);
        for (int32_t j = 0; j < l.n_seg && !bad; j++) {
            for (int k = 0; k < l.seg_len[j] && !bad; k++) {
                const uint32_t word = l.obsT[(size_t) k * capacity + j], kw = l.wkeyT[(size_t) k * capacity + j];
                const uint32_t p = HFG_KEY_ID(kw);
                if ((int32_t) p >= l.n_keys) { bad = 1; break; }
                if (l.kdesc[p] != (word & ~HFG_OBS_CHUNK_END)) bad = 1;
                if (((kw & HFG_KEY_CHUNK_START) != 0) != ((word & HFG_OBS_CHUNK_START) != 0)) bad = 1;
                if (((kw & HFG_KEY_CHUNK_END) != 0) != ((word & HFG_OBS_CHUNK_END) != 0)) bad = 1;
                const double b = (word & HFG_OBS_EDGE) ? l.edge_beta[3 * e] : l.beta0;
                if (l.kbeta[3 * p] != b || l.kbeta[3 * p + 1] != l.beta0 / b) bad = 1;
                if (word & HFG_OBS_EDGE) e++;
                if (key_has_stats(word)) listed++;
            }
        }
        if (listed != l.n_list) bad = 1;
        int32_t at = 0, last_region = -1;
        for (int32_t t = 0; t < l.n_tiles && !bad; t++) {
            const int32_t p = l.tile_key[t];
            if (p < 0 || p >= l.n_keys || l.tile_begin[t] != at || l.tile_cnt[t] < 1 || l.tile_cnt[t] > 4 * HFG_TILE) { bad = 1; break; }
            const int32_t r = (int32_t) HFG_OBS_REGION(l.kdesc[p]);
            if (r < last_region || t < l.region_tile_begin[r] || t >= l.region_tile_begin[r + 1]) bad = 1;
            last_region = r;
            for (int i = 0; i < l.tile_cnt[t] && !bad; i++) {
                const int32_t g = l.klist[at + i];
                if (g < 1 || g >= l.n_windows || seen[g]) { bad = 1; break; }
                if (i > 0 && l.klist[at + i - 1] >= g) bad = 1;
                seen[g] = 1;
            }
            at += l.tile_cnt[t];
        }
        if (at != l.n_list) bad = 1;
        /* the listed windows carry the tile's key */
        for (int32_t j = 0; j < l.n_seg && !bad; j++)
            for (int k = 0; k < l.seg_len[j]; k++) {
                const uint32_t word = l.obsT[(size_t) k * capacity + j];
                if ((seen[l.seg_start[j] + k] != 0) != (key_has_stats(word) != 0)) bad = 1;
            }
        for (int32_t t = 0; t < l.n_tiles && !bad; t++)
            for (int i = 0; i < l.tile_cnt[t]; i++) {
                /* locate the window's segment by binary search over seg_start */
                const int32_t g = l.klist[l.tile_begin[t] + i];
                int32_t lo = 0, hi = l.n_seg - 1;
                while (lo < hi) {
                    const int32_t mid = (lo + hi + 1) / 2;
                    if (l.seg_start[mid] <= g) lo = mid; else hi = mid - 1;
                }
                if ((int32_t) HFG_KEY_ID(l.wkeyT[(size_t) (g - l.seg_start[lo]) * capacity + lo]) != l.tile_key[t]) bad = 1;
            }
        free(seen);
    }
    if (summary) {
        summary[0] = l.n_seg;
        summary[1] = l.smax;
        summary[2] = l.n_edge;
        summary[3] = l.n_windows;
        summary[4] = l.n_keys;
        summary[5] = l.n_tiles;
    }
    hfg_layout_free(&l);
    return bad ? HFG_ERR_INVALID : HFG_OK;
}

/* EM_computeAdjustmentBeta through the C-ABI, for tests */
double hfg_debug_beta(const hfg_config *cfg, const hfg_chunk_desc *chunk, int window) { return hfg_beta(cfg, chunk, window); }
