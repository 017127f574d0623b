/*
 * hfg_cov_reader.c -- the data formats upstream of the hot path: `.cov` / `.cov.gz` -> chunks of windows, and the
 * `.bin` chunk dump.
 *
 * Same window semantics as the reference's chunk builder, restated for speed:
 *   reference                                                   here
 *   ChunksCreator_createCovIndex (chunk.c:240-294)              chunk layout computed from the contig length alone
 *   ChunksCreator_parseOneChunk  (chunk.c:506-547): every chunk  ONE sequential pass over the (gz) file
 *     job re-opens the file and gzseeks from its start
 *   Chunk_addTrack (chunk.c:444-481): a loop over every BASE     one step per (block x window) overlap: run-length aware,
 *     of a block with 3 atof + a Splitter allocation each        O(blocks + windows) instead of O(bases)
 *   Chunk_addWindow (chunk.c:393-441)                            identical: mean -> round -> clip 250; OR of annotation
 *                                                                flags; region / truth / prediction = mode, ties -> lowest
 * Integer-valued coverages (what bam2cov writes) make n*value exact, so the window sums are bit-identical to the
 * reference's base-by-base accumulation; non-integer values fall back to repeated addition to stay so.
 * No `<input>.index` side file is written (chunk.c:154-162 writes one; it is only a cache of the layout).
 * ONE deliberate difference: a single run-length block that crosses two or more chunk boundaries.  The reference's index
 * builder adds at most one chunk per track line, so it drops (or leaves empty) the chunks in between -- a 12 345-base block
 * with -C 3000 gives it 1 chunk / 30 windows (the assert that would catch it is compiled out of release builds).  Here the
 * layout follows from the contig length alone: 4 chunks / 124 windows, every base in a window
 * (tests/test_cov_reader.py::test_one_block_across_several_chunk_boundaries).  bam2cov output (blocks of a few hundred
 * bases against 20 Mb chunks) never meets the case.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <zlib.h>

#include "../../include/hfg_io.h"
#include "hfg_inflate.h"

#define MAX_COVERAGE 250.0 /* chunk.c:8 */
#define REGION_BINS 101    /* Int_getModeValue1DArray(.., 0, 100), chunk.c:388-390 */
#define LABEL_BINS 12      /* Int_getModeValue1DArray(.., -1, 10), chunk.c:376-385 */

typedef struct Growable {
    hfg_cov_data *d;
    int64_t win_cap;
    int32_t chunk_cap;
} Growable;

static int grow_windows(Growable *g, int64_t need) {
    if (need <= g->win_cap) return 1;
    int64_t cap = g->win_cap ? g->win_cap : 4096;
    while (cap < need) cap *= 2;
    hfg_cov_data *d = g->d;
#define GROW(field, type)                                                  \
    do {                                                                   \
        type *p_ = realloc(d->field, sizeof(type) * (size_t) cap);         \
        if (!p_) return 0;                                                 \
        d->field = p_;                                                     \
    } while (0)
    GROW(cov, uint16_t);
    GROW(cov_high_mapq, uint16_t);
    GROW(cov_high_clip, uint16_t);
    GROW(annotation_flag, uint64_t);
    GROW(region, uint8_t);
    GROW(truth, int8_t);
    GROW(prediction, int8_t);
#undef GROW
    g->win_cap = cap;
    return 1;
}

static int grow_chunks(Growable *g, int32_t need) {
    if (need <= g->chunk_cap) return 1;
    int32_t cap = g->chunk_cap ? g->chunk_cap : 256;
    while (cap < need) cap *= 2;
    hfg_chunk_desc *c = realloc(g->d->chunks, sizeof(hfg_chunk_desc) * (size_t) cap);
    if (!c) return 0;
    g->d->chunks = c;
    char(*n)[HFG_CONTIG_NAME_MAX] = realloc(g->d->contig_names, (size_t) cap * HFG_CONTIG_NAME_MAX);
    if (!n) return 0;
    g->d->contig_names = n;
    g->chunk_cap = cap;
    return 1;
}

void hfg_score_cache_free(void *cache); /* hfg_summary.c */

void hfg_cov_free(hfg_cov_data *d) {
    if (!d) return;
    hfg_score_cache_free(d->score_cache);
    for (int i = 0; i < d->n_annotations; i++) free(d->annotation_names ? d->annotation_names[i] : NULL);
    free(d->annotation_names);
    free(d->region_coverages);
    free(d->chunks);
    free(d->contig_names);
    free(d->cov);
    free(d->cov_high_mapq);
    free(d->cov_high_clip);
    free(d->annotation_flag);
    free(d->region);
    free(d->truth);
    free(d->prediction);
    free(d);
}

/* accumulator of the window being filled (Chunk.windowSum*, windowAnnotationFlag, window*Array) */
typedef struct Window {
    int n; /* bases so far (windowItr + 1) */
    double sum_cov, sum_mapq, sum_clip;
    uint64_t flag;
    int32_t region_count[REGION_BINS], truth_count[LABEL_BINS], pred_count[LABEL_BINS];
    /* lowest / highest bin touched since the last reset: a window nearly always sees one region and one label, so the
     * mode and the reset look at that one bin instead of all 101 + 12 + 12 */
    int r_lo, r_hi, t_lo, t_hi, p_lo, p_hi;
} Window;

static void window_init(Window *w) {
    memset(w, 0, sizeof(*w));
    w->r_lo = REGION_BINS;
    w->t_lo = w->p_lo = LABEL_BINS;
    w->r_hi = w->t_hi = w->p_hi = -1;
}

static void window_reset(Window *w) {
    for (int i = w->r_lo; i <= w->r_hi; i++) w->region_count[i] = 0;
    for (int i = w->t_lo; i <= w->t_hi; i++) w->truth_count[i] = 0;
    for (int i = w->p_lo; i <= w->p_hi; i++) w->pred_count[i] = 0;
    w->n = 0;
    w->sum_cov = w->sum_mapq = w->sum_clip = 0.0;
    w->flag = 0;
    w->r_lo = REGION_BINS;
    w->t_lo = w->p_lo = LABEL_BINS;
    w->r_hi = w->t_hi = w->p_hi = -1;
}

static inline void window_count(int32_t *counts, int *lo, int *hi, int bin, int n) {
    counts[bin] += n;
    if (bin < *lo) *lo = bin;
    if (bin > *hi) *hi = bin;
}

/* sum += n copies of v, bit-identical to adding v n times */
static void add_n(double *sum, double v, int n) {
    if (v == floor(v) && fabs(v) < 1e6) {
        *sum += v * n; /* integers: every partial sum is exact */
    } else {
        for (int i = 0; i < n; i++) *sum += v;
    }
}

/* the most frequent value, ties -> the lowest (Int_getModeValue1DArray); bins outside [lo, hi] are empty, and at least one
 * base was counted, so scanning the touched range gives what scanning all bins gives */
static int mode_of(const int32_t *counts, int lo, int hi, int min_value) {
    int best = lo;
    for (int i = lo + 1; i <= hi; i++)
        if (counts[best] < counts[i]) best = i;
    return min_value + best;
}

static int bin_of(int value, int min_value, int bins) {
    int idx = value < min_value ? 0 : value - min_value;
    return idx >= bins ? bins - 1 : idx;
}

static uint16_t clip_round(double v) {
    const double r = round(v);
    return (uint16_t) (MAX_COVERAGE < r ? MAX_COVERAGE : r);
}

/* Chunk_addWindow, chunk.c:393-441 */
static int emit_window(Growable *g, Window *w, int window_len, int start_only) {
    hfg_cov_data *d = g->d;
    if (!grow_windows(g, d->n_windows + 1)) return 0;
    const int64_t i = d->n_windows++;
    double c, m, k;
    if (start_only) {
        c = w->sum_cov * window_len / w->n;
        m = w->sum_mapq * window_len / w->n;
        k = w->sum_clip * window_len / w->n;
    } else {
        c = w->sum_cov / w->n;
        m = w->sum_mapq / w->n;
        k = w->sum_clip / w->n;
    }
    d->cov[i] = clip_round(c);
    d->cov_high_mapq[i] = clip_round(m);
    d->cov_high_clip[i] = clip_round(k);
    const int region = mode_of(w->region_count, w->r_lo, w->r_hi, 0);
    d->region[i] = (uint8_t) region;
    /* CoverageInfo_setRegionIndex, ptBlock.c:300-304 */
    d->annotation_flag[i] = (w->flag & 0x03FFFFFFFFFFFFFFULL) | ((uint64_t) region << 58);
    d->truth[i] = (int8_t) mode_of(w->truth_count, w->t_lo, w->t_hi, -1);
    d->prediction[i] = (int8_t) mode_of(w->pred_count, w->p_lo, w->p_hi, -1);
    d->chunks[d->n_chunks - 1].n_windows++;
    window_reset(w);
    return 1;
}

static int starts_with(const char *s, const char *p) { return strncmp(s, p, strlen(p)) == 0; }

static const char *nth_field(const char *line, char sep, int n) {
    const char *p = line;
    for (int i = 0; i < n && p; i++) {
        p = strchr(p, sep);
        if (p) p++;
    }
    return p;
}

/* Number tokens of a data line.  bam2cov writes plain decimal integers; those are converted by hand (a run of at most 15
 * digits is exact in a double and equals what atof / atol / atoi return), anything else goes through the libc routine the
 * reference uses, so the values are the reference's for every input. */
static inline long tok_long(const char *p) {
    const char *q = p;
    long v = 0;
    while ((unsigned) (*q - '0') <= 9u && q - p < 18) v = v * 10 + (*q++ - '0');
    if (q == p || q - p >= 18) return atol(p); /* sign, blanks, overlong: libc */
    return v;
}

static inline int tok_int(const char *p) {
    const char *q = p;
    int neg = 0;
    if (*q == '-') {
        neg = 1;
        q++;
    }
    const char *d0 = q;
    int v = 0;
    while ((unsigned) (*q - '0') <= 9u && q - d0 < 9) v = v * 10 + (*q++ - '0');
    if (q == d0 || q - d0 >= 9) return atoi(p);
    return neg ? -v : v;
}

static inline double tok_double(const char *p) {
    const char *q = p;
    long v = 0;
    while ((unsigned) (*q - '0') <= 9u && q - p < 15) v = v * 10 + (*q++ - '0');
    /* a pure digit run that ends the token: exact.  A '.', exponent, sign, blank, "nan", more digits: strtod as atof */
    if (q == p || *q != '\0') return atof(p);
    return (double) v;
}

static int fail_io(char *err, size_t errlen, const char *fmt, const char *a, long b) {
    snprintf(err, errlen, fmt, a, b);
    return HFG_ERR_INVALID;
}

/* Line source: the file is read (and inflated) in large blocks and lines are handed out in place, NUL-terminated, without
 * a copy.  zlib's gzgets costs little, but the per-line strlen / strchr / strtod around it were two thirds of the parse
 * time (profiles/host_timing_r1d.txt). */
typedef struct LineSrc {
    gzFile fp;
    char *buf;
    size_t cap, len, pos;
    int eof;
    /* read-ahead: a second thread inflates the next blocks while this one parses (inflating is ~60 % of the time a
     * .cov.gz takes; it cannot be split, a plain gzip stream is sequential, but it can run beside the parser) */
    pthread_t thread;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    char *slot[2];
    long got[2];  /* bytes in the slot; 0: end of file; < 0: error */
    int full[2];
    int stop, threaded;
    unsigned long produced, consumed;
    /* gzip input goes through the own decoder (hfg_inflate.c: about twice zlib's speed on the Huffman-only streams the
     * reference writes); plain text, or HFG_ZLIB_INFLATE=1, through zlib's gzread */
    hfg_inflate *inflate;
    size_t block; /* most bytes one fill delivers */
    int bad_gzip;
} LineSrc;

#define SRC_BLOCK ((size_t) 4 << 20)

static void *src_producer(void *arg) {
    LineSrc *s = arg;
    for (;;) {
        const int k = (int) (s->produced & 1);
        pthread_mutex_lock(&s->mu);
        while (s->full[k] && !s->stop) pthread_cond_wait(&s->cv, &s->mu);
        const int stop = s->stop;
        pthread_mutex_unlock(&s->mu);
        if (stop) break;
        const long got = s->inflate ? hfg_inflate_next(s->inflate, (uint8_t *) s->slot[k])
                                    : (long) gzread(s->fp, s->slot[k], (unsigned) SRC_BLOCK);
        pthread_mutex_lock(&s->mu);
        s->got[k] = got;
        s->full[k] = 1;
        s->produced++;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mu);
        if (got <= 0) break;
    }
    return NULL;
}

/* 1 on success; the reader works without the thread (inline gzread) when it cannot be started */
static int src_open(LineSrc *s, gzFile fp, const char *path) {
    memset(s, 0, sizeof(*s));
    s->fp = fp;
    s->block = SRC_BLOCK;
    if (!getenv("HFG_ZLIB_INFLATE")) s->inflate = hfg_inflate_open(path, SRC_BLOCK); /* NULL for anything but gzip */
    if (s->inflate) s->block = hfg_inflate_piece_capacity(s->inflate);
    s->cap = 2 * s->block + 2;
    s->buf = malloc(s->cap);
    if (!s->buf) return 0;
    s->slot[0] = malloc(s->block);
    s->slot[1] = malloc(s->block);
    if (s->slot[0] && s->slot[1] && pthread_mutex_init(&s->mu, NULL) == 0) {
        if (pthread_cond_init(&s->cv, NULL) == 0) {
            if (pthread_create(&s->thread, NULL, src_producer, s) == 0) s->threaded = 1;
            else pthread_cond_destroy(&s->cv);
        }
        if (!s->threaded) pthread_mutex_destroy(&s->mu);
    }
    return 1;
}

static void src_close(LineSrc *s) {
    if (s->threaded) {
        pthread_mutex_lock(&s->mu);
        s->stop = 1;
        pthread_cond_broadcast(&s->cv);
        pthread_mutex_unlock(&s->mu);
        pthread_join(s->thread, NULL);
        pthread_cond_destroy(&s->cv);
        pthread_mutex_destroy(&s->mu);
    }
    free(s->slot[0]);
    free(s->slot[1]);
    free(s->buf);
    hfg_inflate_close(s->inflate);
}

/* appends the next block behind s->len (room for s->block + 1 bytes is there); returns the byte count, 0 at the end,
 * < 0 on a read / inflate error */
static long src_fill(LineSrc *s) {
    if (!s->threaded)
        return s->inflate ? hfg_inflate_next(s->inflate, (uint8_t *) s->buf + s->len)
                          : (long) gzread(s->fp, s->buf + s->len, (unsigned) SRC_BLOCK);
    const int k = (int) (s->consumed & 1);
    pthread_mutex_lock(&s->mu);
    while (!s->full[k]) pthread_cond_wait(&s->cv, &s->mu);
    pthread_mutex_unlock(&s->mu);
    const long got = s->got[k];
    if (got > 0) memcpy(s->buf + s->len, s->slot[k], (size_t) got);
    pthread_mutex_lock(&s->mu);
    s->full[k] = 0;
    s->consumed++;
    pthread_cond_broadcast(&s->cv);
    pthread_mutex_unlock(&s->mu);
    return got;
}

/* next line -> *line (NUL-terminated in place, CR / LF stripped), *n = its length.  0 at the end of the file, -1 when out
 * of memory. */
static int src_next(LineSrc *s, char **line, size_t *n) {
    for (;;) {
        char *start = s->buf + s->pos;
        char *nl = s->len > s->pos ? memchr(start, '\n', s->len - s->pos) : NULL;
        if (nl || (s->eof && s->len > s->pos)) {
            char *end = nl ? nl : s->buf + s->len; /* the last line may lack its newline; buf has room for the NUL */
            s->pos = (size_t) (end - s->buf) + 1;
            while (end > start && (end[-1] == '\r' || end[-1] == '\n')) end--;
            *end = '\0';
            *line = start;
            *n = (size_t) (end - start);
            return 1;
        }
        if (s->eof) return 0;
        /* keep the unfinished line, refill behind it */
        const size_t rest = s->len - s->pos;
        if (s->pos > 0) memmove(s->buf, start, rest);
        s->pos = 0;
        s->len = rest;
        if (s->cap - s->len < s->block + 1) {
            char *nb = realloc(s->buf, s->cap * 2);
            if (!nb) return -1;
            s->buf = nb;
            s->cap *= 2;
        }
        const long got = src_fill(s);
        if (got < 0) s->bad_gzip = 1;
        if (got <= 0) s->eof = 1;
        else s->len += (size_t) got;
    }
}

/* digits at *pp -> value, *pp moved past them; 0 digits or more than 15 leave *pp where it was (caller falls back) */
static inline long scan_digits(char **pp) {
    char *q = *pp;
    long v = 0;
    while ((unsigned) (*q - '0') <= 9u) v = v * 10 + (*q++ - '0');
    if (q == *pp || q - *pp > 15) return -1;
    *pp = q;
    return v;
}

int hfg_read_cov(const char *path, int32_t chunk_len, int32_t window_len, hfg_cov_data **out, char *err,
                 size_t errlen) {
    if (!path || !out || chunk_len <= 0 || window_len <= 0) return fail_io(err, errlen, "hfg_read_cov: bad argument%s%ld", "", 0);
    gzFile fp = gzopen(path, "rb"); /* transparently reads plain text as well */
    if (!fp) return fail_io(err, errlen, "cannot open %s%.0ld", path, 0);
    gzbuffer(fp, 1 << 20);
    hfg_cov_data *d = calloc(1, sizeof(*d));
    Growable g = {d, 0, 0};
    LineSrc src;
    const int src_ok = src_open(&src, fp, path);
    char *line = NULL;
    Window *win = malloc(sizeof(Window));
    int status = HFG_OK;
    if (!d || !src_ok || !win) {
        status = HFG_ERR_NOMEM;
        goto done;
    }
    d->chunk_len = chunk_len;
    d->window_len = window_len;
    window_init(win);

    char ctg[HFG_CONTIG_NAME_MAX] = "";
    long ctg_len = 0, next_base = 0; /* next base of the contig that must come */
    int have_chunk = 0;
    long line_no = 0;
    size_t n = 0;
    int more;
    while ((more = src_next(&src, &line, &n)) > 0) {
        line_no++;
        if (n == 0) continue;
        if (line[0] == '#') { /* header keys, track_reader.c:220-457 */
            if (starts_with(line, "#annotation:len:")) {
                d->n_annotations = atoi(line + 16);
                d->annotation_names = calloc((size_t) (d->n_annotations > 0 ? d->n_annotations : 1), sizeof(char *));
                for (int i = 0; i < d->n_annotations; i++) d->annotation_names[i] = strdup("NA");
            } else if (starts_with(line, "#annotation:name:")) {
                const int idx = atoi(line + 17);
                const char *name = nth_field(line, ':', 3);
                if (name && idx >= 0 && idx < d->n_annotations) {
                    free(d->annotation_names[idx]);
                    d->annotation_names[idx] = strdup(name);
                }
            } else if (starts_with(line, "#region:len:")) {
                d->n_regions = atoi(line + 12);
                d->region_coverages = calloc((size_t) (d->n_regions > 0 ? d->n_regions : 1), sizeof(int32_t));
            } else if (starts_with(line, "#region:coverage:")) {
                const int idx = atoi(line + 17);
                const char *v = nth_field(line, ':', 3);
                if (v && idx >= 0 && idx < d->n_regions) d->region_coverages[idx] = atoi(v);
            } else if (starts_with(line, "#label:len:")) {
                d->n_labels = atoi(line + 11);
            } else if (starts_with(line, "#truth:true")) {
                d->truth_available = 1;
            } else if (starts_with(line, "#prediction:true")) {
                d->prediction_available = 1;
            } else if (starts_with(line, "#start-only:true")) {
                d->start_only = 1;
            } else if (starts_with(line, "#avg_alignment_len:")) {
                d->avg_alignment_len = atoi(line + 19);
            }
            continue;
        }
        if (line[0] == '>') { /* ">name length" */
            if (have_chunk && next_base != ctg_len) {
                status = fail_io(err, errlen, "%s: contig ended before its declared length (line %ld)", ctg, line_no);
                goto done;
            }
            char *sp = strchr(line, ' ');
            if (!sp) {
                status = fail_io(err, errlen, "malformed contig line%s at line %ld", "", line_no);
                goto done;
            }
            *sp = '\0';
            snprintf(ctg, sizeof(ctg), "%s", line + 1);
            ctg_len = atol(sp + 1);
            next_base = 0;
            have_chunk = 0;
            continue;
        }
        /* data line: start end cov cov_high_mapq cov_high_clip annot[,annot..] region [truth [prediction]] (1-based) */
        long s, e;
        double v_cov, v_mapq, v_clip;
        uint64_t flag = 0; /* CoverageInfo_getAnnotationFlagFromArray, ptBlock.c:225-236 */
        int region_v, truth_v = -1, pred_v = -1;
        int fast = 0;
        {
            /* the canonical line -- unsigned decimal integers, tabs, comma-separated annotation indices, optional labels
             * that may be -1 -- is scanned in one pass; anything else takes the general route below */
            char *p = line;
            long v[5];
            int k = 0;
            for (; k < 5; k++) {
                if ((v[k] = scan_digits(&p)) < 0 || *p != '\t') break;
                p++;
            }
            if (k == 5) {
                for (;;) {
                    const long a = scan_digits(&p);
                    if (a < 0) break;
                    if (a > 0 && a <= 64) flag |= 1ULL << (a - 1);
                    else if (a > 64) { p = NULL; break; }
                    if (*p != ',') break;
                    p++;
                }
                long rv = -1;
                if (p && *p == '\t' && (p++, (rv = scan_digits(&p)) >= 0) && rv < 1000000) {
                    int ok = 1, nlab = 0;
                    long lab[2] = {-1, -1};
                    while (ok && *p == '\t' && nlab < 2) {
                        p++;
                        int neg = 0;
                        if (*p == '-') {
                            neg = 1;
                            p++;
                        }
                        const long t = scan_digits(&p);
                        if (t < 0 || t > 1000000) ok = 0;
                        else lab[nlab++] = neg ? -t : t;
                    }
                    /* a third tab keeps the general route's reading of the ninth field ("rest of the line") */
                    if (ok && *p == '\0') {
                        fast = 1;
                        s = v[0] - 1;
                        e = v[1] - 1;
                        v_cov = (double) v[2];
                        v_mapq = (double) v[3];
                        v_clip = (double) v[4];
                        region_v = (int) rv;
                        if (nlab > 0) truth_v = (int) lab[0];
                        if (nlab > 1) pred_v = (int) lab[1];
                    }
                }
            }
        }
        if (!fast) {
            char *f[9];
            int nf = 0;
            flag = 0;
            for (char *p = line; p && nf < 9;) {
                f[nf++] = p;
                p = strchr(p, '\t');
                if (p) *p++ = '\0';
            }
            if (nf < 7) {
                status = fail_io(err, errlen, "malformed data line%s at line %ld", "", line_no);
                goto done;
            }
            s = tok_long(f[0]) - 1;
            e = tok_long(f[1]) - 1;
            v_cov = tok_double(f[2]);
            v_mapq = tok_double(f[3]);
            v_clip = tok_double(f[4]);
            for (char *p = f[5]; p && *p;) {
                const int a = tok_int(p);
                if (a > 0) flag |= 1ULL << (a - 1);
                p = strchr(p, ',');
                if (p) p++;
            }
            region_v = tok_int(f[6]);
            truth_v = nf >= 8 ? tok_int(f[7]) : -1;
            pred_v = nf >= 9 ? tok_int(f[8]) : -1;
        }
        if (ctg[0] == '\0') {
            status = fail_io(err, errlen, "malformed data line%s at line %ld", "", line_no);
            goto done;
        }
        if (s != next_base || e < s || e >= ctg_len) {
            status = fail_io(err, errlen, "%s: blocks must tile the contig in order (line %ld)", ctg, line_no);
            goto done;
        }
        const int rbin = bin_of(region_v, 0, REGION_BINS);
        const int tbin = bin_of(truth_v, -1, LABEL_BINS);
        const int pbin = bin_of(pred_v, -1, LABEL_BINS);
        long pos = s;
        while (pos <= e) {
            if (!have_chunk || pos > d->chunks[d->n_chunks - 1].e) {
                /* next chunk of this contig (ChunksCreator_createCovIndex, chunk.c:260-287): canonical length, the last
                 * one absorbs a remainder shorter than a full chunk */
                if (!grow_chunks(&g, d->n_chunks + 1)) {
                    status = HFG_ERR_NOMEM;
                    goto done;
                }
                hfg_chunk_desc *c = &d->chunks[d->n_chunks];
                memset(c, 0, sizeof(*c));
                c->ctg_len = (int32_t) ctg_len;
                c->s = (int32_t) pos;
                c->e = (int32_t) (ctg_len < (pos - 1) + 2L * chunk_len ? ctg_len - 1 : (pos - 1) + chunk_len);
                if (pos == 0) c->e = (int32_t) (ctg_len < 2L * chunk_len ? ctg_len - 1 : chunk_len - 1);
                c->window_len = window_len;
                c->offset = d->n_windows;
                snprintf(d->contig_names[d->n_chunks], HFG_CONTIG_NAME_MAX, "%s", ctg);
                d->n_chunks++;
                have_chunk = 1;
            }
            const hfg_chunk_desc *c = &d->chunks[d->n_chunks - 1];
            /* bases of this block that fall into the window being filled */
            long room = window_len - win->n;
            long upto = pos + room - 1;
            if (upto > e) upto = e;
            if (upto > c->e) upto = c->e;
            const int nb = (int) (upto - pos + 1);
            add_n(&win->sum_cov, v_cov, nb);
            add_n(&win->sum_mapq, v_mapq, nb);
            add_n(&win->sum_clip, v_clip, nb);
            win->flag |= flag;
            window_count(win->region_count, &win->r_lo, &win->r_hi, rbin, nb);
            window_count(win->truth_count, &win->t_lo, &win->t_hi, tbin, nb);
            window_count(win->pred_count, &win->p_lo, &win->p_hi, pbin, nb);
            win->n += nb;
            pos = upto + 1;
            if (win->n == window_len || upto == c->e) { /* full window, or the short last window of the chunk */
                if (!emit_window(&g, win, window_len, d->start_only)) {
                    status = HFG_ERR_NOMEM;
                    goto done;
                }
            }
        }
        next_base = e + 1;
    }
    if (more < 0) {
        status = HFG_ERR_NOMEM;
        goto done;
    }
    if (src.bad_gzip) {
        if (src.inflate) snprintf(err, errlen, "%s: corrupt gzip data (%s)", path, hfg_inflate_error(src.inflate));
        else snprintf(err, errlen, "%s: read error", path);
        status = HFG_ERR_INVALID;
        goto done;
    }
    if (have_chunk && next_base != ctg_len) {
        status = fail_io(err, errlen, "%s: file ended before the contig's declared length%.0ld", ctg, 0);
        goto done;
    }
    if (d->n_annotations <= 0) {
        status = fail_io(err, errlen, "no '#annotation:len:' in the header of %s%.0ld", path, 0); /* track_reader.c:239-249 */
        goto done;
    }
done:
    src_close(&src); /* joins the read-ahead thread before the file is closed */
    gzclose(fp);
    free(win);
    if (status != HFG_OK) {
        if (status == HFG_ERR_NOMEM) snprintf(err, errlen, "out of memory reading %s", path);
        hfg_cov_free(d);
        d = NULL;
    }
    *out = d;
    return status;
}

/* ChunksCreator_parseChunksFromBinaryFile, chunk.c:713-828 (little-endian, C bool = 1 byte) */
int hfg_read_bin(const char *path, hfg_cov_data **out, char *err, size_t errlen) {
    if (!path || !out) return fail_io(err, errlen, "hfg_read_bin: bad argument%s%ld", "", 0);
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail_io(err, errlen, "cannot open %s%.0ld", path, 0);
    hfg_cov_data *d = calloc(1, sizeof(*d));
    Growable g = {d, 0, 0};
    int status = HFG_OK;
    int32_t v;
#define RD(ptr, size, count)                                   \
    do {                                                       \
        if (fread(ptr, size, count, fp) != (size_t) (count)) { \
            status = HFG_ERR_INVALID;                          \
            goto done;                                         \
        }                                                      \
    } while (0)
    RD(&d->n_annotations, 4, 1);
    if (d->n_annotations < 0 || d->n_annotations > 64) {
        status = HFG_ERR_INVALID;
        goto done;
    }
    d->annotation_names = calloc((size_t) (d->n_annotations > 0 ? d->n_annotations : 1), sizeof(char *));
    for (int i = 0; i < d->n_annotations; i++) {
        RD(&v, 4, 1);
        if (v <= 0 || v > 4096) {
            status = HFG_ERR_INVALID;
            goto done;
        }
        d->annotation_names[i] = malloc((size_t) v);
        RD(d->annotation_names[i], 1, v);
    }
    RD(&d->n_regions, 4, 1);
    if (d->n_regions < 0 || d->n_regions > 4096) {
        status = HFG_ERR_INVALID;
        goto done;
    }
    d->region_coverages = calloc((size_t) (d->n_regions > 0 ? d->n_regions : 1), sizeof(int32_t));
    RD(d->region_coverages, 4, d->n_regions);
    RD(&d->n_labels, 4, 1);
    {
        uint8_t b[3];
        RD(b, 1, 3);
        d->truth_available = b[0];
        d->prediction_available = b[1];
        d->start_only = b[2];
    }
    RD(&d->avg_alignment_len, 4, 1);
    RD(&d->chunk_len, 4, 1);
    RD(&d->window_len, 4, 1);
    while (fread(&v, 4, 1, fp) == 1) {
        if (v <= 0 || v > HFG_CONTIG_NAME_MAX || !grow_chunks(&g, d->n_chunks + 1)) {
            status = v <= 0 || v > HFG_CONTIG_NAME_MAX ? HFG_ERR_INVALID : HFG_ERR_NOMEM;
            goto done;
        }
        hfg_chunk_desc *c = &d->chunks[d->n_chunks];
        memset(c, 0, sizeof(*c));
        RD(d->contig_names[d->n_chunks], 1, v);
        d->contig_names[d->n_chunks][HFG_CONTIG_NAME_MAX - 1] = '\0';
        RD(&c->ctg_len, 4, 1);
        RD(&c->s, 4, 1);
        RD(&c->e, 4, 1);
        RD(&c->n_windows, 4, 1);
        if (c->n_windows < 0) {
            status = HFG_ERR_INVALID;
            goto done;
        }
        c->window_len = d->window_len;
        c->offset = d->n_windows;
        if (!grow_windows(&g, d->n_windows + c->n_windows)) {
            status = HFG_ERR_NOMEM;
            goto done;
        }
        const int64_t o = d->n_windows;
        const int L = c->n_windows;
        RD(d->cov + o, 2, L);
        RD(d->cov_high_mapq + o, 2, L);
        RD(d->cov_high_clip + o, 2, L);
        RD(d->annotation_flag + o, 8, L);
        RD(d->truth + o, 1, L);
        RD(d->prediction + o, 1, L);
        for (int i = 0; i < L; i++) d->region[o + i] = (uint8_t) (d->annotation_flag[o + i] >> 58); /* ptBlock.c:294-298 */
        d->n_windows += L;
        d->n_chunks++;
    }
#undef RD
done:
    fclose(fp);
    if (status != HFG_OK) {
        snprintf(err, errlen, status == HFG_ERR_NOMEM ? "out of memory reading %s" : "%s is not a valid chunk dump", path);
        hfg_cov_free(d);
        d = NULL;
    }
    *out = d;
    return status;
}
