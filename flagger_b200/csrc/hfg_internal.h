/*
 * hfg_internal.h -- shared between the host C files and the CUDA translation unit of libhfg.
 * Not part of the public ABI (that is include/hfg.h).
 */
#ifndef HFG_INTERNAL_H
#define HFG_INTERNAL_H

#include <stdint.h>
#include "../../include/hfg.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HFG_NS HFG_NUM_STATES
#define HFG_PI 3.14159      /* the reference's PI (submodules/common/common.h:15) -- NOT M_PI, on purpose */
#define HFG_TERM_PROB 1e-4  /* Transition.terminationProb (hmm_utils.c:2112) */
#define HFG_MAX_CLASSES 16  /* distinct (state, alpha) emission classes per window */

/* ---- packed observation word (one per window, stored segment-transposed on the device) ----------------
 *  bits  0..7   x      = (uint8_t) coverage                      (hmm.c:345,384)
 *  bits  8..15  px     = x of the previous window of the chunk (0 at a chunk start, hmm.c:338)
 *  bits 16..21  region = CoverageInfo_getRegionIndex             (ptBlock.c:294-304)
 *  bits 22..24  validity mask: bit22 Dup invalid, bit23 Col invalid, bit24 END column valid
 *                                                                 (hmm_utils.c:2229-2264)
 *  bit  25      region differs from the previous window -> transition is the constant 1/5 (hmm.c:398-400)
 *  bit  26      first window of a chunk (EM_fillFirstColumnForward, hmm.c:333-364)
 *  bit  27      second window of a chunk: pair 0->1 is skipped by the statistics (hmm.c:638-642)
 *  bit  28      edge window: beta differs from the interior constant, value in the edge list (hmm.c:301-316)
 *  bit  29      last window of a chunk (EM_fillLastColumnBackward, hmm.c:452-467)
 *  bit  31      slot holds a window (0 = padding of a short segment)
 */
#define HFG_OBS_X(w) ((w) & 0xffu)
#define HFG_OBS_PX(w) (((w) >> 8) & 0xffu)
#define HFG_OBS_REGION(w) (((w) >> 16) & 0x3fu)
#define HFG_OBS_MASK(w) (((w) >> 22) & 0x7u)
#define HFG_OBS_REGION_CHANGE (1u << 25)
#define HFG_OBS_CHUNK_START (1u << 26)
#define HFG_OBS_SECOND (1u << 27)
#define HFG_OBS_EDGE (1u << 28)
#define HFG_OBS_CHUNK_END (1u << 29)
#define HFG_OBS_VALID (1u << 31)

/* ---- observation keys -----------------------------------------------------------------------------------
 * Observations are small integers, so the per-window transfer matrix M_i = T_i (.) E_i takes few distinct values: it
 * depends only on (x, px, region, validity mask, region change, chunk start, beta).  The host numbers the distinct
 * combinations ("keys", hottest first inside a region); the kernel evaluates emissions and M once per KEY and per
 * E-step, the windows gather M by key id, and the pair statistics are accumulated per key (lists of the windows of
 * each key, cut into tiles of HFG_TILE windows, or longer ones when that
 * leaves every thread of the grid at most one tile).  Per-window key word, stored segment-transposed:
 *  bits 0..27  key id      bit 30  last window of a chunk      bit 31  first window of a chunk */
#define HFG_KEY_ID(w) ((w) & 0x0fffffffu)
#define HFG_KEY_MAX 0x0fffffff
#define HFG_KEY_CHUNK_END (1u << 30)
#define HFG_KEY_CHUNK_START (1u << 31)
#define HFG_TILE 16

/* Host-built, run-constant device layout: the genome is ONE sequence of windows cut into segments that never
 * straddle a chunk or a region change; segment j is owned by global thread j of the E-step kernel and its k-th
 * window lives at obsT[k * capacity + j] (coalesced across threads). */
typedef struct hfg_layout {
    int64_t n_windows;
    int32_t n_chunks;
    int32_t capacity;  /* threads of the persistent grid = segments slots */
    int32_t smax;      /* windows per segment slot */
    int32_t n_seg;
    double beta0;      /* interior beta: (Lr-1)/Lr, or 1 when contig ends are not adjusted */
    uint32_t *obsT;          /* [smax][capacity] packed observation words (host only: the keys are built from them) */
    uint32_t *wkeyT;         /* [smax][capacity] key word of every window (device) */
    int32_t n_keys;
    uint32_t *kdesc;         /* [n_keys] packed observation word of the key (without the chunk-end bit) */
    double *kbeta;           /* [n_keys][3] (beta, beta0/beta, sqrt(beta0/beta)); (beta0, 1, 1) for interior keys */
    int64_t n_list;
    int32_t *klist;          /* [n_list] global window indices grouped by key, ascending inside a key; windows whose pair
                                is skipped by the statistics (chunk starts, second windows) are not listed */
    int32_t n_tiles;
    int32_t *tile_key, *tile_begin, *tile_cnt; /* [n_tiles] consecutive entries of klist (at most 4 * HFG_TILE), all of one key */
    int32_t region_tile_begin[HFG_MAX_REGIONS + 1]; /* tiles are grouped by region */
    int32_t *seg_start;      /* [capacity] global index of the segment's first window (0 for idle slots) */
    int32_t *seg_len;        /* [capacity] 0 for idle slots */
    int32_t *seg_chunk;      /* [capacity] */
    int32_t *seg_edge_begin; /* [capacity + 1] first entry of edge_beta that belongs to segment j or later */
    double *edge_beta;       /* [n_edge][3] (beta, beta0/beta, sqrt(beta0/beta)) of every edge window, in window order */
    int64_t n_edge;
    int64_t *chunk_offset;   /* [n_chunks + 1] */
    int32_t *edge_head, *edge_tail; /* [n_chunks] leading / trailing contig-end windows of every chunk (segments-only build) */
    int32_t tile_len;        /* windows per statistics tile */
} hfg_layout;

/* Builds the layout (host memory, malloc'd; free with hfg_layout_free) for at most `capacity` segment slots; with
 * granule > 0 the slot count is then trimmed to the multiple of `granule` (threads per CTA) the segments need
 * (out->capacity).  Returns hfg_status. */
int hfg_layout_build(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                     const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                     int32_t capacity, int32_t granule, hfg_layout *out, char *err, size_t errlen);
int hfg_layout_build_ex(const hfg_config *cfg, int32_t n_chunks, const hfg_chunk_desc *chunks, const uint16_t *cov,
                        const uint16_t *cov_high_mapq, const uint16_t *cov_high_clip, const uint8_t *region,
                        int32_t capacity, int32_t granule, int segments_only, hfg_layout *out, char *err, size_t errlen);
void hfg_layout_free(hfg_layout *l);
/* Segment slots per statistics worker: the tile length is the shortest one that leaves every WORKER at most one tile
 * (1: a worker per slot; 4: the E-step kernel folds the statistics with four lanes per tile).  Process-wide; every context
 * uses the same kernel generation, so the same value. */
extern int32_t hfg_layout_tile_div;

/* EM_computeAdjustmentBeta (hmm.c:301-316) for window i of a chunk. */
double hfg_beta(const hfg_config *cfg, const hfg_chunk_desc *ch, int i);

/* Emission classes: per state the distinct alpha values of its column of the alpha matrix (hmm.c:388).
 * cls[pre][s] indexes the per-window emission row; the row also serves chunk starts (alpha = 0), whose four
 * values are kept in slots 0..3. */
typedef struct hfg_classes {
    int32_t n_classes;                    /* D, >= 4 */
    int32_t cls[HFG_NS][HFG_NS];          /* [pre][s] -> slot */
    int32_t class_state[HFG_MAX_CLASSES]; /* slot -> state */
    double class_alpha[HFG_MAX_CLASSES];  /* slot -> alpha */
    int32_t n_class_of_state[HFG_NS];
    int32_t first_class_of_state[HFG_NS];
    int32_t is_gaussian[HFG_NS];
} hfg_classes;

void hfg_classes_build(const hfg_config *cfg, const double *alpha, hfg_classes *out);

/* negative-binomial model, host side (hfg_nb.c) */
void hfg_nb_init_component(double mean, double *theta, double *lambda);
int hfg_nb_mstep_region(const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st, double tol);

#ifdef __cplusplus
}
#endif
#endif
