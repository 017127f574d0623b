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

#include "hfg_internal.h"

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

/* Cost of a window in units of an ordinary window: contig-end (edge) windows take the generic emission path
 * with per-window factors; the kernel folds them into the tabulated constants, so they cost the same as any other
 * window (HFG_EDGE_COST 1).  The cost-based cut is kept so that a costlier special path can be balanced by data. */
#define HFG_EDGE_COST 1

/* end (exclusive) of the segment that starts at window i of a chunk: same region, accumulated cost <= smax */
static int segment_end(const uint8_t *region, const uint8_t *cost, int L, int i, int smax) {
    int j = i + 1, acc = cost[i] & 0x7f; /* bit 7 of cost[] flags an edge window, the low bits are the cost */
    while (j < L && region[j] == region[i] && acc + (cost[j] & 0x7f) <= smax) acc += cost[j++] & 0x7f;
    return j;
}

static int64_t count_segments(int32_t n_chunks, const hfg_chunk_desc *chunks, const uint8_t *region,
                              const uint8_t *cost, int smax) {
    int64_t n = 0;
    for (int32_t c = 0; c < n_chunks; c++) {
        const int64_t o = chunks[c].offset;
        const int L = chunks[c].n_windows;
        for (int i = 0; i < L; n++) i = segment_end(region + o, cost + o, L, i, smax);
    }
    return n;
}

int hfg_layout_build(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                     const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                     int32_t capacity, hfg_layout *out, char *err, size_t errlen) {
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
    for (int64_t i = 0; i < W; i++) {
        if (region[i] >= cfg->n_regions) {
            snprintf(err, errlen, "window %lld has region %d >= n_regions %d", (long long) i, region[i], cfg->n_regions);
            return HFG_ERR_INVALID;
        }
    }
    const double beta0 = !cfg->adjust_contig_ends ? 1.0
                         : (cfg->mean_read_length > 0 ? (double) (cfg->mean_read_length - 1) / cfg->mean_read_length : 0.25);
    uint8_t *cost = malloc((size_t) W);
    if (!cost) {
        snprintf(err, errlen, "out of host memory building the layout");
        return HFG_ERR_NOMEM;
    }
    int64_t total_cost = 0;
    for (int32_t c = 0; c < n_chunks; c++) {
        for (int i = 0; i < chunks[c].n_windows; i++) {
            const double b = hfg_beta(cfg, &chunks[c], i);
            cost[chunks[c].offset + i] = memcmp(&b, &beta0, sizeof(double)) != 0 ? (0x80 | HFG_EDGE_COST) : 1;
            total_cost += cost[chunks[c].offset + i] & 0x7f;
        }
    }
    /* smallest cost budget whose segment count fits the persistent grid */
    int smax = (int) ((total_cost + capacity - 1) / capacity);
    if (smax < HFG_EDGE_COST) smax = HFG_EDGE_COST;
    while (count_segments(n_chunks, chunks, region, cost, smax) > capacity) smax += (smax + 7) / 8;
    const int64_t n_seg = count_segments(n_chunks, chunks, region, cost, smax);

    out->n_windows = W;
    out->n_chunks = n_chunks;
    out->capacity = capacity;
    out->smax = smax;
    out->n_seg = (int32_t) n_seg;
    out->beta0 = beta0;
    out->obsT = calloc((size_t) smax * capacity, sizeof(uint32_t));
    out->seg_start = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_len = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_chunk = calloc((size_t) capacity, sizeof(int32_t));
    out->seg_edge_begin = calloc((size_t) capacity + 1, sizeof(int32_t));
    out->chunk_offset = calloc((size_t) n_chunks + 1, sizeof(int64_t));
    int64_t edge_cap = 1024, n_edge = 0;
    out->edge_beta = malloc(sizeof(double) * 3 * (size_t) edge_cap);
    if (!out->obsT || !out->seg_start || !out->seg_len || !out->seg_chunk || !out->seg_edge_begin ||
        !out->edge_beta || !out->chunk_offset) {
        free(cost);
        hfg_layout_free(out);
        snprintf(err, errlen, "out of host memory building the layout");
        return HFG_ERR_NOMEM;
    }

    int32_t seg = 0;
    for (int32_t c = 0; c < n_chunks; c++) {
        const hfg_chunk_desc *ch = &chunks[c];
        const int64_t o = ch->offset;
        const int L = ch->n_windows;
        out->chunk_offset[c] = o;
        for (int a = 0; a < L;) {
            const int end = segment_end(region + o, cost + o, L, a, smax);
            const int len = end - a;
            out->seg_start[seg] = (int32_t) (o + a);
            out->seg_len[seg] = len;
            out->seg_chunk[seg] = c;
            out->seg_edge_begin[seg] = (int32_t) n_edge;
            for (int k = 0; k < len; k++) {
                const int w = a + k; /* window index inside the chunk */
                const int64_t g = o + w;
                uint32_t word = HFG_OBS_VALID;
                word |= (uint32_t) (uint8_t) cov[g];
                if (w > 0) word |= (uint32_t) (uint8_t) cov[g - 1] << 8;
                word |= (uint32_t) region[g] << 16;
                word |= validity_mask(cfg, cov[g], cov_high_mapq[g], cov_high_clip[g]) << 22;
                if (w > 0 && region[g] != region[g - 1]) word |= HFG_OBS_REGION_CHANGE;
                if (w == 0) word |= HFG_OBS_CHUNK_START;
                if (w == 1) word |= HFG_OBS_SECOND;
                if (w == L - 1) word |= HFG_OBS_CHUNK_END;
                if (cost[g] & 0x80) {
                    word |= HFG_OBS_EDGE;
                    if (n_edge == edge_cap) {
                        edge_cap *= 2;
                        double *nb = realloc(out->edge_beta, sizeof(double) * 3 * (size_t) edge_cap);
                        if (!nb) {
                            free(cost);
                            hfg_layout_free(out);
                            snprintf(err, errlen, "out of host memory building the layout");
                            return HFG_ERR_NOMEM;
                        }
                        out->edge_beta = nb;
                    }
                    {
                        /* (beta, beta0/beta, sqrt(beta0/beta)) */
                        const double b = hfg_beta(cfg, ch, w);
                        out->edge_beta[3 * n_edge] = b;
                        out->edge_beta[3 * n_edge + 1] = beta0 / b;
                        out->edge_beta[3 * n_edge + 2] = sqrt(beta0 / b);
                        n_edge++;
                    }
                }
                out->obsT[(size_t) k * capacity + seg] = word;
            }
            seg++;
            a = end;
        }
    }
    free(cost);
    out->chunk_offset[n_chunks] = W;
    for (int32_t j = seg; j <= capacity; j++) out->seg_edge_begin[j] = (int32_t) n_edge;
    out->n_edge = n_edge;
    return HFG_OK;
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
    int rc = hfg_layout_build(cfg, n_chunks, chunks, cov, cov_high_mapq, cov_high_clip, region, capacity, &l, err,
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
