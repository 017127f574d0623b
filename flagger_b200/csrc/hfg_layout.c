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

/* submodules/common/common.c:142-148: min/max are int functions; double arguments are truncated at the call */
static int imin_(int a, int b) { return a < b ? a : b; }
static int imax_(int a, int b) { return a < b ? b : a; }

double hfg_beta(const hfg_config *cfg, const hfg_chunk_desc *ch, int i) {
    if (!cfg->adjust_contig_ends) return 1.0;
    const double frac = cfg->min_read_fraction_at_ends;
    const int Lr = cfg->mean_read_length;
    const int mid = imin_((int) (ch->s + (double) ch->window_len * (i + 0.5)),
                          (int) ((ch->s + (double) ch->window_len * i + ch->e) / 2));
    const int lo = imax_(mid - Lr + 1, (int) (-(1 - frac) * Lr));
    const int hi = imin_(mid, (int) (ch->ctg_len - frac * Lr));
    const double b = (double) (hi - lo) / Lr;
    return b <= 0.25 ? 0.25 : b;
}

static uint32_t validity_mask(const hfg_config *cfg, uint16_t cov, uint16_t mapq, uint16_t clip) {
    const double rm = (double) mapq / (0.1 + cov);
    const double rc = (double) clip / (0.1 + cov);
    uint32_t m = 0;
    if (rm > cfg->max_high_mapq_ratio) m |= 1u;          /* Dup invalid */
    if (rm < cfg->min_high_mapq_ratio) m |= 2u;          /* Col invalid */
    if (!(rc < cfg->min_highly_clipped_ratio)) m |= 4u;  /* END column valid */
    return m;
}

void hfg_layout_free(hfg_layout *l) {
    if (!l) return;
    free(l->obsT);
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
            uint32_t word = HFG_OBS_VALID;
            word |= (uint32_t) (uint8_t) jb->cov[g];
            if (w > 0) word |= (uint32_t) (uint8_t) jb->cov[g - 1] << 8;
            word |= (uint32_t) jb->region[g] << 16;
            word |= validity_mask(jb->cfg, jb->cov[g], jb->mapq[g], jb->clip[g]) << 22;
            if (w > 0 && jb->region[g] != jb->region[g - 1]) word |= HFG_OBS_REGION_CHANGE;
            if (w == 0) word |= HFG_OBS_CHUNK_START;
            if (w == 1) word |= HFG_OBS_SECOND;
            if (w == L - 1) word |= HFG_OBS_CHUNK_END;
            if (w < jb->edge_head[c] || w >= L - jb->edge_tail[c]) {
                /* (beta, beta0/beta, sqrt(beta0/beta)) */
                const double b = hfg_beta(jb->cfg, ch, w);
                word |= HFG_OBS_EDGE;
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

/* number of edge windows among the first w windows of chunk c */
static int64_t edges_before(int head, int tail, int L, int w) {
    int64_t n = w < head ? w : head;
    if (w > L - tail) n += w - (L - tail);
    return n;
}

int hfg_layout_build(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                     const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                     int32_t capacity, int32_t granule, hfg_layout *out, char *err, size_t errlen) {
    memset(out, 0, sizeof(*out));
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
            while (j < L && r[j] == r[i]) j++;
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
    while (segments_for(runs, n_runs, smax) > capacity) smax += (smax + 7) / 8;
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
    out->obsT = calloc((size_t) smax * capacity, sizeof(uint32_t));
    out->seg_start = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_len = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_chunk = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_edge_begin = calloc((size_t) capacity + 1, sizeof(int32_t));
    out->chunk_offset = calloc((size_t) n_chunks + 1, sizeof(int64_t));
    out->edge_beta = malloc(sizeof(double) * 3 * (size_t) (n_edge > 0 ? n_edge : 1));
    if (!out->obsT || !out->seg_start || !out->seg_len || !out->seg_chunk || !out->seg_edge_begin || !out->edge_beta ||
        !out->chunk_offset)
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
    if (summary) {
        summary[0] = l.n_seg;
        summary[1] = l.smax;
        summary[2] = l.n_edge;
        summary[3] = l.n_windows;
    }
    hfg_layout_free(&l);
    return bad ? HFG_ERR_INVALID : HFG_OK;
}

/* EM_computeAdjustmentBeta through the C-ABI, for tests */
double hfg_debug_beta(const hfg_config *cfg, const hfg_chunk_desc *chunk, int window) { return hfg_beta(cfg, chunk, window); }
