/*
 * hfg_estep.cuh -- the E-step kernel of libhfg (sm_100a, fp64, no tensor cores: a 4-state trellis is not a
 * dense contraction).
 *
 * What it replaces: the per-chunk worker of the reference, EM_runForward + EM_runBackward + EM_updateEstimators +
 * the EM_getMostProbableState loop (submodules/hmm/hmm.c:423-434, 535-545, 638-650, 715-737), for ALL chunks of
 * one EM_runOneIterationForList call (hmm.c:739-780), in ONE persistent cooperative launch.
 *
 * Shape of the computation:
 *   The reference walks every chunk serially (5-10k dependent 4x4 steps) and evaluates 16 emission pdfs per window
 *   four times per iteration.  Here
 *   (1) observations are small integers, so the window transfer matrix M_i = T_i (.) E_i takes few distinct values
 *       (~10^4 "keys" (x, px, region, mask, beta) for ~10^6 windows, numbered once on the host): emissions and M are
 *       evaluated once per KEY per E-step into a table that stays in L2/L1, windows gather their matrix by key id;
 *   (2) the whole genome is one sequence of windows cut into segments (<= smax windows, never straddling a chunk or
 *       a region change); global thread j owns segment j.  A chunk start is just a window whose transfer matrix is
 *       rank-1 (all rows = the start column), so chunks need no special casing in the scan;
 *   (3) the pair statistics are linear in xi_i[pre][s] = f^_{i-1}[pre] M_i[pre][s] b_i[s], and everything else in
 *       them depends only on the key: sum_i xi_i = M_key (.) sum_i f^_{i-1} (x) b_i.  The sweeps only store f^ and
 *       the normalised b; the outer products are summed per key (tiles of <= 16 windows), and the estimator updates of
 *       the reference run once per tile instead of once per window.
 *     phase T   per key: emission classes, transfer matrix -> table (grid barrier).
 *     phase A   per thread: product of the window matrices of the segment (power-of-two rescaled: exact scaling).
 *     phase B   matrix scans: warp shuffles -> warps of a block through shared memory -> blocks through global
 *               memory and a grid barrier.  Yields for every segment the forward message entering it and the direction
 *               of the backward message entering it.
 *     phase C1  the reference's scaled forward recurrence inside the segment (f^, c_i and sum log c_i).
 *     phase C2  backward recurrence fused with the posterior-argmax decode.  b is carried as a direction and
 *               normalised per window so that sum_{pre,s} xi_i = 1, the invariant of the reference's scaling
 *               (sum_s f^_i[s] b^_i[s] c_i = terminationProb with counts divided by terminationProb, hmm.c:452-467,
 *               613-614) (grid barrier).
 *     phase S   per tile: outer-product sum, then the reference's estimator updates with the tile's pooled counts.
 *     phase D   deterministic reduction of the statistics: per-thread -> per-block (fixed order) -> grid (fixed
 *               order, block 0 after the last grid barrier).  No floating-point atomics anywhere.
 *
 * Results differ from the reference only by rounding (~1e-16 relative per operation): t*e is rounded once per key
 * instead of (f*t)*e per window, messages enter the segments from a scan, libdevice-free exp (<= 1-2 ulp from glibc),
 * statistics summed in tile/tree order.
 */
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hfg_internal.h"

#define HFG_HD static __device__
#define HFG_FIT_RATE hfg_fit_rate_warp
#include "hfg_mstep_inl.h"
#undef HFG_HD

namespace cg = cooperative_groups;

/* threads per CTA: 512 when the per-thread statistics columns fit shared memory, else 256 (more mixture components /
 * regions); one persistent CTA per SM either way */
#define HFG_THREADS_MAX 512
#define HFG_THREADS_MIN 256

/* Per-region derived tables in shared memory (doubles):
 *   [0,128)    conditional transition  Tc[mask][pre*4+s]          (Transition_getProbConditional, hmm_utils.c:2278-2292)
 *   [128,144)  the constant 1/(N+1) of a region-change window, laid out as a ninth mask
 *   [144,148)  start probabilities trans[4][s]
 *   [148,152)  termination probabilities trans[s][4]
 *   [152,156)  lambda, truncPoint, lambda/beta0, 1-exp(-(lambda/beta0)*(beta0*truncPoint))
 *   then 6 arrays of G doubles over the flattened Gaussian components g: mu, var, w, var*beta0, 1/(var*beta0),
 *   w/sqrt(var*beta0*2*PI). */
#define RT_TC 0
#define RT_UNI 128 /* 16 x 1/(N+1): the transition "matrix" of a region-change window (hmm.c:398-400), as mask index 8 */
#define RT_START 144
#define RT_TERM 148
#define RT_TEXP 152
#define RT_GAUSS 156
#define HFG_INV_TERM 1e4 /* 1 / terminationProb */
#define HFG_MAX_PEERS 8
#define HFG_PC_STRIDE 16 /* phase-clock slots per block */
#define HFG_MAX_TASKS 160
/* per-(region, task) table after the Gaussian arrays: (1-a)*mu, a, 1/(var*beta0), w/sqrt(var*beta0*2*PI) */
#define RT_TASK(G) (RT_GAUSS + 6 * (G))
#define RT_STRIDE2(G, NT) (RT_TASK(G) + 4 * (NT)) /* doubles per region */

struct EstepArgs {
    /* run-constant layout */
    const uint32_t *wkeyT;     /* [smax][capacity] key word of every window */
    const int32_t *seg_start, *seg_len;
    int32_t capacity, smax, n_regions;
    double beta0;
    /* observation keys and the per-key window lists of the statistics */
    int32_t n_keys;
    const uint32_t *kdesc;     /* [n_keys] packed observation word */
    const double *kbeta;       /* [n_keys][3] beta, beta0/beta, sqrt(beta0/beta) */
    const int32_t *klist;      /* window indices grouped by key */
    const int32_t *tile_key, *tile_begin, *tile_cnt;
    const int32_t *region_tile_begin; /* [HFG_MAX_REGIONS + 1] */
    /* model structure */
    int32_t n_classes;              /* D */
    int32_t cls[HFG_NS][HFG_NS];    /* [pre][s] -> slot */
    double alpha[HFG_NS][HFG_NS];   /* [pre][s]; 0 for non-Gaussian states */
    int32_t class_state[HFG_MAX_CLASSES];
    double class_alpha[HFG_MAX_CLASSES];
    int32_t is_gauss[HFG_NS], ncomp[HFG_NS], gbase[HFG_NS];
    int32_t G;                      /* total Gaussian components */
    double inv_one_minus_alpha[HFG_NS][HFG_NS]; /* 1 / (1 - alpha[pre][s]) */
    int32_t first_pre_of_class[HFG_NS][HFG_NS];  /* [pre][s]: smallest preState sharing the class of (pre, s) */
    uint32_t slots_start, slots_other;           /* emission slots evaluated at a chunk start / elsewhere */
    /* flat list of the Gaussian component evaluations of an ordinary window, ordered by (class, component): the kernel
     * runs them four at a time so that four independent exp() chains are in flight per thread */
    int32_t n_tasks;
    uint8_t task_class[HFG_MAX_TASKS], task_comp[HFG_MAX_TASKS];
    int32_t texp_slot;                           /* emission slot of the truncated-exponential state, or -1 */
    /* per-call inputs / scratch / outputs (device) */
    const hfg_region_params *params;
    double *tabM;  /* [n_keys][16]  transfer matrix of every key (row-major [pre][s]) */
    double *scrFT; /* [smax][4][capacity]  scaled forward f^, segment-transposed (written in C1, read by the same thread in C2) */
    double *scrXB; /* [W][8]        per-window record of the statistics: f^ of the previous window, then the backward message
                                    normalised so that sum_{pre,s} f^_{i-1}[pre] M_i[pre][s] b_i[s] = 1 */
    double *block_tot;   /* [grid][16] */
    int32_t *block_reset; /* [grid] */
    double *partials;    /* [R][NSTAT][grid] */
    double *out;         /* [R * sizeof(hfg_region_stats)/8 + 2]: stats | loglik | error flags (as double) */
    double *out_host;    /* optional mirror of `out` in mapped pinned host memory (blocking calls: no read-back copy) */
    double *seg_loglik;  /* [capacity] */
    int8_t *labels;      /* [W] */
    int8_t *labels_host; /* optional: mapped pinned host buffer; every block copies its own labels there while the
                            statistics phase runs (no device-to-host copy after the kernel) */
    int32_t n_seg;
    int64_t n_windows;
    double *posteriors;  /* [W][4] or NULL */
    int32_t *err_flags;  /* bit0 scale underflow, bit1 NaN */
    int32_t forward_only;
    long long *phase_clock; /* [grid][HFG_PC_STRIDE] clock64() of thread 0 at the phase boundaries + SM id (instrumentation) */
    /* multi-GPU: in-kernel sum all-reduce of [stats | loglik | flags] over peer memory (NVLink P2P).  Every rank owns a
     * mailbox  [2 epochs][HFG_MAX_PEERS senders][out_doubles]  followed by  [2][HFG_MAX_PEERS]  arrival counters;
     * peer_box[r] is rank r's mailbox as mapped into this process (peer_box[rank] is the local one). */
    int32_t n_ranks, rank, out_doubles;
    double *peer_box[HFG_MAX_PEERS];
    unsigned long long *epoch; /* device counter of exchanges done so far (identical on every rank) */
    /* device-resident EM loop (hfg_em_begin / hfg_em_enqueue): the parameters stay on the device and block 0 runs the
     * M-step (HMM_estimateParameters) on the reduced statistics in the tail of the kernel, so that successive
     * iterations are back-to-back launches with no host in between.  em_mode 0 = plain E-step, 1 = E-step + M-step
     * (skipped entirely once the stop flag is up), 2 = final inference pass (no M-step). */
    int32_t em_mode, model_type, em_max_logliks;
    double em_tol;
    hfg_region_params *em_params; /* the same memory as `params` */
    int32_t *em_state;            /* [0] stop (converged or failed), [1] E-steps run, [2] error flags, [3] converged */
    double *em_logliks;           /* log-likelihood of every E-step run */
    /* negative-binomial instantiation only (hfg_estep_kernel<THREADS, true>; NOT YET VALIDATED ON HARDWARE, DESIGN.md
     * section 7).  Appended last so that the offsets the other instantiations read do not move. */
    int32_t *ticket;        /* quad kernel: arrival counter of the grid reduction (zero between launches) */
    double *tabMT;          /* quad kernel: [n_keys][16] the transposed transfer matrices (row s = column s of tabM) */
    const uint32_t *wposT;  /* quad kernel: [smax][capacity] position of the window in the key lists (klist), or 0xffffffff */
    /* third-generation kernel (hfg_estep_v3.cuh) */
    const int32_t *hot_key; /* [n_hot] key id of every slot of the shared-memory matrix table */
    int32_t n_hot;          /* slots in use */
    const int32_t *hot_range; /* [n_regions][3] first hot key of the region, number of hot keys, first slot */
    int32_t lab_bytes;      /* bytes of shared memory for the staged labels (multiple of 16; 0: labels go straight to global memory) */
    double *scan_stash;     /* [32][capacity] exclusive prefix / suffix product of every thread inside its warp */
    int32_t dbg;            /* instrumentation switch (HFG_DBG) */
    int32_t work_doubles;   /* doubles of shared memory behind the region tables (the M-step work area of the kernel tail) */
    const double *nb_table; /* [R][4][HFG_NB_XSTRIDE] pmf of (region, state, x), evaluated on the host (hfg_nb.c) */
    double *nb_tile_col;    /* [n_tiles][4] pair mass of every statistics tile by state: the host folds it into the
                               (region, state, x) histogram the model's estimators are fed from */
    /* negative binomial, device-resident loop (hfg_nb_dev.cuh) */
    const int32_t *nb_bin_begin; /* [R * 250 + 1] first entry of every (region, coverage bin) in nb_bin_tiles */
    const int32_t *nb_bin_tiles; /* [n_tiles] the tiles of every bin, in tile order */
    const double *nb_lgx1;       /* [251] lgamma(x + 1) (host libm values) */
    double *nb_hist;             /* [R][4][256] pair mass by (region, state, coverage bin): folded by the whole grid */
    /* blocking calls, fast path (hfg_api.cu::run_blocking): a single-region model travels in the kernel arguments instead of
     * through an upload, and the tail announces the results with one store the host polls instead of synchronising the
     * stream */
    unsigned long long *done_flag; /* mapped pinned host word, or NULL */
    unsigned long long done_seq;   /* the value that says "this call's results are in host memory" */
    int32_t params_inline;         /* read region 0's parameters from inl_params, not from `params` */
    hfg_region_params inl_params;
};
#define HFG_NB_XSTRIDE 256

/* statistic columns per (block, region): 16 transition counts, lambda num/den, then per Gaussian component
 * (meanNum, den, varNum), then the log-likelihood (kept in region 0's row). */
__host__ __device__ inline int hfg_nstat(int G) { return 18 + 3 * G + 1; }
/* rows of the shared staging area: statistics, or the 32-row scan stash, whichever is larger */
__host__ __device__ inline int hfg_acc_rows(int G) { return hfg_nstat(G) > 32 ? hfg_nstat(G) : 32; }

namespace hfgk {

__device__ __forceinline__ double pow2_rescale_factor(double m) {
    /* 2^-floor(log2 m): an exact scaling, so rescaled products stay exact up to the matmul roundings */
    const int hi = __double2hiint(m);
    const int e = (hi >> 20) & 0x7ff;
    return __hiloint2double((2046 - e) << 20, 0);
}

__device__ __forceinline__ void mat_rescale(double (&P)[16]) {
    /* entries are non-negative, so their order is the order of their high words as integers */
    int hi = __double2hiint(P[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) hi = max(hi, __double2hiint(P[i]));
    const double s = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
#pragma unroll
    for (int i = 0; i < 16; i++) P[i] *= s;
}

/* C = A * B (row-major 4x4); scan arithmetic uses fused multiply-adds: no reference order exists for it */
__device__ __forceinline__ void mat_mul(const double (&A)[16], const double (&B)[16], double (&C)[16]) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double acc = A[r * 4] * B[c];
            acc = fma(A[r * 4 + 1], B[4 + c], acc);
            acc = fma(A[r * 4 + 2], B[8 + c], acc);
            acc = fma(A[r * 4 + 3], B[12 + c], acc);
            C[r * 4 + c] = acc;
        }
    }
}

__device__ __forceinline__ void mat_identity(double (&P)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) P[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

__device__ __forceinline__ void vec_normalize(double (&v)[4]) {
    const double s = 1.0 / (((v[0] + v[1]) + v[2]) + v[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] *= s;
}

/* v <- v * A */
__device__ __forceinline__ void vec_mat(double (&v)[4], const double (&A)[16]) {
    double o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) o[c] = fma(v[3], A[12 + c], fma(v[2], A[8 + c], fma(v[1], A[4 + c], v[0] * A[c])));
#pragma unroll
    for (int c = 0; c < 4; c++) v[c] = o[c];
}

/* u <- A * u */
__device__ __forceinline__ void mat_vec(const double (&A)[16], double (&u)[4]) {
    double o[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
        o[r] = fma(A[r * 4 + 3], u[3], fma(A[r * 4 + 2], u[2], fma(A[r * 4 + 1], u[1], A[r * 4] * u[0])));
#pragma unroll
    for (int r = 0; r < 4; r++) u[r] = o[r];
}

__device__ __forceinline__ void mat_shfl_up(const double (&P)[16], double (&Q)[16], int off) {
#pragma unroll
    for (int i = 0; i < 16; i++) Q[i] = __shfl_up_sync(0xffffffffu, P[i], off);
}
__device__ __forceinline__ void mat_shfl_down(const double (&P)[16], double (&Q)[16], int off) {
#pragma unroll
    for (int i = 0; i < 16; i++) Q[i] = __shfl_down_sync(0xffffffffu, P[i], off);
}

/* A <- A * B, row by row in place (keeps the live set at two matrices + one row) */
__device__ __forceinline__ void mat_mul_inplace_left(double (&A)[16], const double (&B)[16]) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
        double t[4];
#pragma unroll
        for (int c = 0; c < 4; c++)
            t[c] = fma(A[r * 4 + 3], B[12 + c], fma(A[r * 4 + 2], B[8 + c], fma(A[r * 4 + 1], B[4 + c], A[r * 4] * B[c])));
#pragma unroll
        for (int c = 0; c < 4; c++) A[r * 4 + c] = t[c];
    }
}

/* inclusive prefix products over the first N lanes of a warp: P_l <- P_0 * ... * P_l */
template <int N>
__device__ __forceinline__ void warp_scan_prefix(double (&P)[16], int lane) {
#pragma unroll 1
    for (int off = 1; off < N; off <<= 1) {
        double Q[16];
        mat_shfl_up(P, Q, off);
        if (lane >= off) {
            mat_mul_inplace_left(Q, P); /* Q <- Q * P : earlier lanes on the left */
            mat_rescale(Q);
#pragma unroll
            for (int i = 0; i < 16; i++) P[i] = Q[i];
        }
    }
}

/* inclusive suffix products over the first N lanes: P_l <- P_l * ... * P_{N-1} */
template <int N>
__device__ __forceinline__ void warp_scan_suffix(double (&P)[16], int lane) {
#pragma unroll 1
    for (int off = 1; off < N; off <<= 1) {
        double Q[16];
        mat_shfl_down(P, Q, off);
        if (lane + off < N) {
            mat_mul_inplace_left(P, Q);
            mat_rescale(P);
        }
    }
}

/* ---- emissions ------------------------------------------------------------------------------------------ */

/* exp(q) for q <= 0 (every exponential on this path has a non-positive argument: -0.5*d^2/var, -lambda*x).
 * q is clamped at -708 (e^-708 ~ 1e-308 is far below the 1e-40 floor applied to every pdf), which removes the
 * denormal/overflow handling of the generic routine; 2^n by exponent arithmetic; the degree-11 polynomial is the
 * Chebyshev interpolant of e^r on |r| <= ln2/2 (max relative error 1.6e-17 before rounding), evaluated with Estrin's
 * scheme for instruction-level parallelism.  Measured <= 1 ulp from glibc on the test grid (tests/test_gpu_parity.py). */
__device__ __forceinline__ double exp_nonpos(double q) {
    q = fmax(q, -708.0);
    const double t = fma(q, 1.4426950408889634, 6755399441055744.0); /* q*log2(e) + 1.5*2^52: integer in the low word */
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double r = fma(nf, -6.9314718036912382e-01, q);
    r = fma(nf, -1.9082149292705877e-10, r);
    const double r2 = r * r;
    const double p01 = 1.0 + r; /* c0 + c1*r with c0 = c1 = 1 */
    const double p23 = fma(0.1666666666666668, r, 0.5000000000000019);
    const double p45 = fma(0.008333333333319589, r, 0.04166666666648795);
    const double p67 = fma(0.00019841269890076403, r, 0.0013888888952352863);
    const double p89 = fma(2.755724088722987e-06, r, 2.4801485441561313e-05);
    const double pab = fma(2.5110049204818658e-08, r, 2.763265472252779e-07);
    const double r4 = r2 * r2;
    const double q0 = fma(p23, r2, p01);
    const double q1 = fma(p67, r2, p45);
    const double q2 = fma(pab, r2, p89);
    const double r8 = r4 * r4;
    const double h = fma(q1, r4, q0);
    const double p = fma(q2, r8, h);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

struct Win {
    double x, px, beta;
    double rb, sq; /* beta0/beta and its square root: rescale the tabulated 1/(var*beta0) and w/sqrt(var*beta0*2*PI) */
    int region, mask;
    bool edge, start, second, region_change, chunk_end;
};

__device__ __forceinline__ Win decode_word(uint32_t w, double beta0) {
    Win o;
    o.x = (double) HFG_OBS_X(w);
    o.px = (double) HFG_OBS_PX(w);
    o.region = (int) HFG_OBS_REGION(w);
    o.mask = (w & HFG_OBS_REGION_CHANGE) ? 8 : (int) HFG_OBS_MASK(w); /* 8 = the constant 1/(N+1) table */
    o.edge = (w & HFG_OBS_EDGE) != 0;
    o.start = (w & HFG_OBS_CHUNK_START) != 0;
    o.second = (w & HFG_OBS_SECOND) != 0;
    o.region_change = (w & HFG_OBS_REGION_CHANGE) != 0;
    o.chunk_end = (w & HFG_OBS_CHUNK_END) != 0;
    o.beta = beta0;
    o.rb = 1.0;
    o.sq = 1.0;
    return o;
}

/* TruncExponential_getProb (hmm_utils.c:941-947) */
__device__ __forceinline__ double trunc_exp_prob(const double *rt, const Win &w) {
    const double trunc = rt[RT_TEXP + 1];
    if (trunc < w.x) return 0.0;
    double lam, norm;
    if (!w.edge) {
        lam = rt[RT_TEXP + 2];
        norm = rt[RT_TEXP + 3];
    } else {
        lam = rt[RT_TEXP] / w.beta;
        const double b = w.beta * trunc;
        norm = 1 - exp_nonpos(-lam * b);
    }
    return lam * exp_nonpos(-lam * w.x) / norm;
}

/* one mixture component of Gaussian_getComponentProbs (hmm_utils.c:768-793): mean=((1-a)*mu + a*px)*beta,
 * var*=beta, w/sqrt(var*2*PI)*exp(-0.5*(x-mean)^2/var), floored at 1e-40; PI is the reference's 3.14159 */
__device__ __forceinline__ double gauss_comp(const double *rt, int G, int g, double a, const Win &w, int *nan) {
    const double *ga = rt + RT_GAUSS;
    double mean = (1 - a) * ga[g] + a * w.px;
    mean *= w.beta;
    /* contig-end windows (beta != beta0) reuse the tabulated constants through beta0/beta: 1/(var*beta) =
     * (1/(var*beta0))*(beta0/beta) and w/sqrt(var*beta*2*PI) = (w/sqrt(var*beta0*2*PI))*sqrt(beta0/beta); both factors
     * are exactly 1 for ordinary windows */
    const double inv = ga[4 * G + g] * w.rb;
    const double coef = ga[5 * G + g] * w.sq;
    const double d = w.x - mean;
    double p = coef * exp_nonpos((-0.5 * (d * d)) * inv);
    if (p != p) *nan = 1;
    if (p < 1e-40) p = 1e-40;
    return p;
}

/* all components of a Gaussian state under dependency factor a: returns their sum (Gaussian_getProb,
 * hmm_utils.c:753-758) and, when pc != NULL, the component pdfs.  Deliberately NOT inlined: one copy of the
 * exponential / edge-window code for the whole kernel keeps the instruction footprint inside the I-cache. */
__device__ __noinline__ double gauss_class(const double *rt, int G, int g0, int n, double a, double x, double px,
                                           double beta, double rb, double sq, double *pc, int *nan) {
    Win w;
    w.x = x;
    w.px = px;
    w.beta = beta;
    w.rb = rb;
    w.sq = sq;
    double tot = 0.0;
    for (int c = 0; c < n; c++) {
        const double p = gauss_comp(rt, G, g0 + c, a, w, nan);
        if (pc) pc[c] = p;
        tot += p;
    }
    return tot;
}

/* Gaussian_updateEstimator (hmm_utils.c:812-839) for ALL (state, alpha) classes of one multi-component state at one
 * window.  The reference adds, per preState and component c,  w = count * p_c / sum_c p_c  with count = f*t*e*b/term and
 * e = sum_c p_c: the class emission cancels, w = H * p_c with H = (f*t) * b / term pooled over the preStates of the class.
 * So no division and no second pass: for each component the (<= 4) classes are evaluated side by side (independent
 * exponential chains), their (w*x_adj, w, w*z*z) are summed in registers and the thread's statistics column is updated
 * once per component (three rows per component, stride ld). */
__device__ __noinline__ void gauss_state_stats(const double *rt, int G, int g0, int n, int nd, double H0, double H1,
                                               double H2, double H3, double a0, double a1, double a2, double a3, double x,
                                               double px, double beta, double rb, double sq, double *colg, int ld,
                                               int *nan) {
    const double H[4] = {H0, H1, H2, H3}, a[4] = {a0, a1, a2, a3};
    double xa[4], oma[4], base[4];
#pragma unroll
    for (int d = 0; d < 4; d++) {
        oma[d] = 1.0 - a[d];
        xa[d] = (x - a[d] * px) / oma[d]; /* x_adjusted, hmm_utils.c:818 */
        base[d] = a[d] * px;
    }
    const double *ga = rt + RT_GAUSS;
#pragma unroll 1
    for (int c = 0; c < n; c++) {
        const int g = g0 + c;
        const double mu = ga[g], inv = ga[4 * G + g] * rb, coef = ga[5 * G + g] * sq;
        double s1 = 0.0, s0 = 0.0, s2 = 0.0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            if (d >= nd) break;
            double mean = oma[d] * mu + base[d]; /* (1-a)*mu + a*px */
            mean *= beta;
            const double dd = x - mean;
            double p = coef * exp_nonpos((-0.5 * (dd * dd)) * inv);
            if (p != p) *nan = 1;
            if (p < 1e-40) p = 1e-40;
            const double wgt = H[d] * p;
            const double z = (xa[d] - mu) * oma[d];
            s1 += wgt * xa[d];
            s0 += wgt;
            s2 += wgt * z * z;
        }
        double *cg_ = colg + (size_t) 3 * c * ld;
        cg_[0] += s1;
        cg_[ld] += s0;
        cg_[2 * ld] += s2;
    }
}

/* emission of state s under dependency factor a (EmissionDist_getProb, hmm_utils.c:1409-1417) */
__device__ __forceinline__ double emission(const EstepArgs &A, const double *rt, int s, double a, const Win &w,
                                           int *nan) {
    if (!A.is_gauss[s]) return trunc_exp_prob(rt, w);
    return gauss_class(rt, A.G, A.gbase[s], A.ncomp[s], a, w.x, w.px, w.beta, w.rb, w.sq, NULL, nan);
}

/* transition probability into window w: 1/(N+1) at a region change (hmm.c:398-400), else the masked row */
__device__ __forceinline__ double trans_prob(const double *rt, const Win &w, int pre, int s) {
    return rt[RT_TC + w.mask * 16 + pre * 4 + s]; /* w.mask == 8 selects the constant table at a region change */
}

}  // namespace hfgk

/* ---------------------------------------------------------------------------------------------------------- */

namespace hfgk {

/* one 4x4 transfer matrix of the key table: four 256-bit loads (LDG.E.256).  The table is written in phase T and only
 * read after the grid barrier that follows it, so the ordinary (L1-cached) path is safe; hot keys stay in L1. */
__device__ __forceinline__ void load_mat(const double *p, double (&M)[16]) {
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(M[4 * q]), "=d"(M[4 * q + 1]), "=d"(M[4 * q + 2]), "=d"(M[4 * q + 3])
            : "l"(p + 4 * q));
}

/* ---- warp-cooperative gather of the 32 transfer matrices of one window step -------------------------------------
 * A thread-private gather (four 256-bit loads per lane) touches 32 different 128-byte lines per instruction: 128 L1
 * wavefronts per warp and step, and the L1 wavefront queue was the kernel's top unit (profiles/).  Here four adjacent
 * lanes read the four 32-byte quarters of ONE matrix, so an instruction touches 8 lines and the four instructions of a
 * step 32 lines in all; the quarters go through a padded shared-memory tile (row stride 18 doubles: conflict-free for
 * the 128-bit stores and loads) from which every lane then reads its own matrix. */
#define HFG_STAGE_LD 18
__device__ __forceinline__ void coop_fetch(const double *tab, uint32_t key, int lane, double (&raw)[16]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t ks = __shfl_sync(0xffffffffu, key, q * 8 + (lane >> 2));
        const double *p = tab + (size_t) ks * 16 + (lane & 3) * 4;
        asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(raw[4 * q]), "=d"(raw[4 * q + 1]), "=d"(raw[4 * q + 2]), "=d"(raw[4 * q + 3])
            : "l"(p));
    }
}
__device__ __forceinline__ void coop_deliver(double *stage, int lane, const double (&raw)[16], double (&M)[16]) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
        double2 *d = reinterpret_cast<double2 *>(stage + (q * 8 + (lane >> 2)) * HFG_STAGE_LD + (lane & 3) * 4);
        d[0] = make_double2(raw[4 * q], raw[4 * q + 1]);
        d[1] = make_double2(raw[4 * q + 2], raw[4 * q + 3]);
    }
    __syncwarp();
    const double2 *sp = reinterpret_cast<const double2 *>(stage + lane * HFG_STAGE_LD);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const double2 v = sp[i];
        M[2 * i] = v.x;
        M[2 * i + 1] = v.y;
    }
    __syncwarp();
}

/* 32-byte rows of the per-window records: 256-bit stores, and 256-bit L2 loads in the statistics phase (the records
 * are written by other SMs before the preceding grid barrier and read once) */
__device__ __forceinline__ void store_vec4(double *p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
/* `after` is a value produced behind the grid barrier that precedes the reads (hfgk::fence_token): the load is an ordinary
 * (non-volatile) asm, so the compiler may batch several of them, but none can move above the barrier */
__device__ __forceinline__ void load_vec4_cg(const double *p, unsigned after, double (&v)[4]) {
    asm("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p + after));
}
__device__ __forceinline__ unsigned fence_token() {
    unsigned t;
    asm volatile("mov.u32 %0, 0;" : "=r"(t)::"memory");
    return t;
}

}  // namespace hfgk

template <int THREADS, bool NB = false>
__global__ void __launch_bounds__(THREADS, 1) hfg_estep_kernel(const EstepArgs A) {
    constexpr int WARPS = THREADS / 32;
    using namespace hfgk;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];

    /* device-resident EM: the previous launch raised the stop flag (converged, or a fatal condition): nothing to do.
     * The flag is read by every thread of every block before any barrier, so the whole grid leaves together. */
    if (A.em_mode == 1 && __ldcg(&A.em_state[0]) != 0) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j = blockIdx.x * THREADS + tid; /* segment owned by this thread */
    const int cap = A.capacity, D = A.n_classes, G = A.G, R = A.n_regions;
    const int rt_stride = RT_STRIDE2(G, A.n_tasks);
    const int NSTAT = hfg_nstat(G);
    const int LD = THREADS + 1;
    const int n_threads = gridDim.x * THREADS;

    /* shared memory carve-up */
    double *rtab = smem;                                   /* [R][rt_stride] */
    double *warp_tot = rtab + (size_t) R * rt_stride;      /* [WARPS][16] warp products */
    double *warp_pre = warp_tot + WARPS * 16;          /* [WARPS][16] exclusive prefix over warps */
    double *warp_suf = warp_pre + WARPS * 16;          /* [WARPS][16] exclusive suffix over warps */
    double *blk_vec = warp_suf + WARPS * 16;           /* [8] entering forward / backward message of the block */
    double *acc = blk_vec + 8;                             /* [max(NSTAT,32)][LD]: emission staging (phase T), scan stash
                                                              (phase B), per-thread statistics (phases S, D) */
    __shared__ int s_reset;

    /* ---- prologue: derived per-region tables (redundantly per block; O(R*K) work) ---- */
    if (tid == 0) s_reset = 0;
    for (int idx = tid; idx < R * 32; idx += THREADS) {
        /* one (region, mask, pre) row of the conditional transition table */
        const int r = idx >> 5, mask = (idx >> 2) & 7, pre = idx & 3;
        const hfg_region_params &p = A.params[r];
        bool valid[5] = {true, (mask & 1) == 0, true, (mask & 2) == 0, (mask & 4) != 0};
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 5; k++)
            if (valid[k]) tot += p.trans[pre][k];
#pragma unroll
        for (int s = 0; s < 4; s++)
            rtab[(size_t) r * rt_stride + RT_TC + mask * 16 + pre * 4 + s] = valid[s] ? p.trans[pre][s] / tot : 0.0;
    }
    /* (the three table loops start on different warps so that, for few regions, they run side by side) */
    for (int idx = (tid + THREADS - 64) % THREADS; idx < R * (12 + G); idx += THREADS) {
        const int r = idx / (12 + G), q = idx % (12 + G);
        const hfg_region_params &p = A.params[r];
        double *rt = rtab + (size_t) r * rt_stride;
        if (q < 4) {
            rt[RT_START + q] = p.trans[HFG_NS][q];
#pragma unroll
            for (int i = 0; i < 4; i++) rt[RT_UNI + q * 4 + i] = 1.0 / (HFG_NS + 1);
        }
        else if (q < 8) rt[RT_TERM + q - 4] = p.trans[q - 4][HFG_NS];
        else if (q == 8) rt[RT_TEXP] = p.lambda;
        else if (q == 9) rt[RT_TEXP + 1] = p.trunc_point;
        else if (q == 10) rt[RT_TEXP + 2] = p.lambda / A.beta0;
        else if (q == 11) {
            const double lam = p.lambda / A.beta0;
            const double b = A.beta0 * p.trunc_point;
            rt[RT_TEXP + 3] = 1 - exp_nonpos(-lam * b);
        } else {
            const int g = q - 12;
            int s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (A.is_gauss[k] && g >= A.gbase[k] && g < A.gbase[k] + A.ncomp[k]) s = k;
            const int c = g - A.gbase[s];
            double *ga = rt + RT_GAUSS;
            const double vb = p.var[s][c] * A.beta0;
            ga[g] = p.mean[s][c];
            ga[G + g] = p.var[s][c];
            ga[2 * G + g] = p.weight[s][c];
            ga[3 * G + g] = vb;
            ga[4 * G + g] = 1.0 / vb;
            ga[5 * G + g] = p.weight[s][c] / sqrt(vb * 2 * HFG_PI);
        }
    }
    for (int idx = (tid + THREADS - 128) % THREADS; idx < R * A.n_tasks; idx += THREADS) {
        const int r = idx / A.n_tasks, t = idx % A.n_tasks;
        const hfg_region_params &p = A.params[r];
        const int d = A.task_class[t], c = A.task_comp[t], s = A.class_state[d];
        const double a = A.class_alpha[d];
        const double vb = p.var[s][c] * A.beta0;
        double *tk = rtab + (size_t) r * rt_stride + RT_TASK(G) + 4 * t;
        tk[0] = (1 - a) * p.mean[s][c];
        tk[1] = a;
        tk[2] = 1.0 / vb;
        tk[3] = p.weight[s][c] / sqrt(vb * 2 * HFG_PI);
    }
    __syncthreads();

    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 0] = clock64();
    int nan_flag = 0, uf_flag = 0;

    /* =========================== phase T: emission classes and transfer matrix of every key ================== */
    /* keys (and below, statistics tiles) are dealt to warps round-robin over the blocks, 32 consecutive ones per warp */
    const int deal0 = (warp * gridDim.x + blockIdx.x) * 32 + lane;
    for (int p = deal0; p < A.n_keys; p += n_threads) {
        Win w = decode_word(A.kdesc[p], A.beta0);
        if (w.edge) {
            w.beta = A.kbeta[3 * (size_t) p];
            w.rb = A.kbeta[3 * (size_t) p + 1];
            w.sq = A.kbeta[3 * (size_t) p + 2];
        }
        const double *rt = rtab + (size_t) w.region * rt_stride;
        double *es = acc + tid; /* es[d * LD]: this thread's emission row of the current key */
        /* every distinct (state, alpha) class once (the reference evaluates all 16 (pre,state) pairs of every WINDOW,
         * three times per iteration, plus once more inside the estimator update) */
        if constexpr (NB) {
            /* the emission depends on (region, state, x) alone (NegativeBinomial_getProb, hmm_utils.c:479-516): a look-up
             * in the host-built table; the host passes alpha = 0, so slot s is the class of every (pre, s) */
            const double *tb = A.nb_table + (size_t) w.region * 4 * HFG_NB_XSTRIDE + (int) w.x;
#pragma unroll
            for (int s = 0; s < 4; s++) es[s * LD] = tb[s * HFG_NB_XSTRIDE];
        } else if (w.start) {
            /* chunk starts (alpha = 0, preX = 0 for every state): generic path */
            for (int d = 0; d < D; d++) {
                if (!((A.slots_start >> d) & 1u)) continue;
                es[d * LD] = emission(A, rt, A.class_state[d], A.class_alpha[d], w, &nan_flag);
            }
        } else {
            /* per-component constants tabulated for the interior beta; four evaluations in flight */
            for (int d = 0; d < D; d++) es[d * LD] = 0.0;
            if (A.texp_slot >= 0) es[A.texp_slot * LD] = trunc_exp_prob(rt, w);
            const double *tk = rt + RT_TASK(G);
            const int NT = A.n_tasks;
            for (int t0 = 0; t0 < NT; t0 += 4) {
                double pv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int t = min(t0 + u, NT - 1);
                    const double4 c4 = *reinterpret_cast<const double4 *>(tk + 4 * t);
                    double mean = c4.x + c4.y * w.px; /* (1-a)*mu + a*px */
                    mean *= w.beta;
                    const double dd = w.x - mean;
                    double pp = (c4.w * w.sq) * exp_nonpos((-0.5 * (dd * dd)) * (c4.z * w.rb));
                    if (pp != pp) nan_flag = 1;
                    pv[u] = pp < 1e-40 ? 1e-40 : pp;
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (t0 + u < NT) es[A.task_class[t0 + u] * LD] += pv[u]; /* component order, as Gaussian_getProb */
            }
        }
        double M[16];
        if (w.start) {
            /* EM_fillFirstColumnForward (hmm.c:333-364): preX = 0, alpha = 0, start probabilities, no mask:
             * a rank-1 transfer matrix, every row = the unnormalised first column */
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const double f0 = es[s * LD] * rt[RT_START + s];
#pragma unroll
                for (int pre = 0; pre < 4; pre++) M[pre * 4 + s] = f0;
            }
        } else {
#pragma unroll
            for (int pre = 0; pre < 4; pre++)
#pragma unroll
                for (int s = 0; s < 4; s++) M[pre * 4 + s] = trans_prob(rt, w, pre, s) * es[A.cls[pre][s] * LD];
        }
        double2 *dst = reinterpret_cast<double2 *>(A.tabM + (size_t) p * 16);
#pragma unroll
        for (int q = 0; q < 8; q++) dst[q] = make_double2(M[2 * q], M[2 * q + 1]);
    }
    grid.sync(); /* the key table is complete */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 1] = clock64();

    const int len = A.seg_len[j];
    const int seg_first = A.seg_start[j];
    const uint32_t *wk = A.wkeyT + j; /* wk[k * cap]: key word of the k-th window of this segment */
    const int kmax = __reduce_max_sync(0xffffffffu, len); /* longest segment of this warp */
    double *stage = acc + (size_t) warp * 32 * HFG_STAGE_LD; /* this warp's gather tile (aliases the staging area) */

    /* =========================== phase A: segment transfer product ============================================ */
    {
        double P[16];
        mat_identity(P);
        bool has_start = false;
        {
            /* the key word is fetched two windows ahead, the matrices one window ahead of their use; lanes whose segment
             * is shorter than the longest of the warp keep taking part in the cooperative gather (key 0) */
            uint32_t w0 = len > 0 ? __ldg(wk) : 0u;
            uint32_t w1 = len > 1 ? __ldg(wk + cap) : 0u;
            double Mc[16], raw[16];
            if (kmax > 0) {
                coop_fetch(A.tabM, HFG_KEY_ID(w0), lane, raw);
                coop_deliver(stage, lane, raw, Mc);
            }
#pragma unroll 1
            for (int k = 0; k < kmax; k++) {
                const uint32_t w2 = k + 2 < len ? __ldg(wk + (size_t) (k + 2) * cap) : 0u;
                if (k + 1 < kmax) coop_fetch(A.tabM, HFG_KEY_ID(w1), lane, raw);
                if (k < len) {
                    if (w0 & HFG_KEY_CHUNK_START) has_start = true;
                    mat_mul_inplace_left(P, Mc);
                    if ((k & 3) == 3) mat_rescale(P); /* a window shrinks the product by < 1e-50: every 4th step is ample */
                }
                if (k + 1 < kmax) coop_deliver(stage, lane, raw, Mc);
                w0 = w1;
                w1 = w2;
            }
        }
        mat_rescale(P);
        if (has_start) s_reset = 1; /* benign race: every writer stores 1 */
        __syncthreads(); /* the gather tiles of all warps are done: the area becomes the scan stash */

        /* ======================= phase B: scans ============================================================== */
        double S[16], Q[16];
#pragma unroll
        for (int i = 0; i < 16; i++) S[i] = P[i];
        warp_scan_prefix<32>(S, lane);
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 16; i++) warp_tot[warp * 16 + i] = S[i];
        }
        mat_shfl_up(S, Q, 1);
        if (lane == 0) mat_identity(Q);
#pragma unroll
        for (int i = 0; i < 16; i++) acc[(size_t) i * LD + tid] = Q[i]; /* exclusive prefix inside the warp */
        warp_scan_suffix<32>(P, lane);
        mat_shfl_down(P, Q, 1);
        if (lane == 31) mat_identity(Q);
#pragma unroll
        for (int i = 0; i < 16; i++) acc[(size_t) (16 + i) * LD + tid] = Q[i]; /* exclusive suffix inside the warp */
    }
    __syncthreads();
    /* second level over the WARPS warp products of this block: warp 0 builds the prefixes (and the block total),
     * warp 1 the suffixes, concurrently */
    if (warp == 0) {
        double Sp[16], Q[16];
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) Sp[i] = warp_tot[lane * 16 + i];
        } else {
            mat_identity(Sp);
        }
        warp_scan_prefix<WARPS>(Sp, lane);
        if (lane == WARPS - 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) A.block_tot[(size_t) blockIdx.x * 16 + i] = Sp[i];
            A.block_reset[blockIdx.x] = s_reset;
        }
        mat_shfl_up(Sp, Q, 1);
        if (lane == 0) mat_identity(Q);
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) warp_pre[lane * 16 + i] = Q[i];
        }
    } else if (warp == 1) {
        double Ss[16], Q[16];
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) Ss[i] = warp_tot[lane * 16 + i];
        } else {
            mat_identity(Ss);
        }
        warp_scan_suffix<WARPS>(Ss, lane);
        mat_shfl_down(Ss, Q, 1);
        if (lane >= WARPS - 1) mat_identity(Q);
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) warp_suf[lane * 16 + i] = Q[i];
        }
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 2] = clock64();
    grid.sync(); /* orders the block totals written above (the barrier fences) */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 3] = clock64();

    /* messages entering this block: the products of the blocks back to (and including) the nearest block that contains a
     * chunk start -- its product is rank-1, so nothing beyond it matters -- and, for the backward message, forward to the
     * nearest such block.  Warp 0 builds the forward message, warp 1 the backward one.  A walk of one or two blocks (the
     * usual case: about one chunk start per block) is done by one lane; longer walks (few chunks spread over many blocks, as
     * on a small multi-GPU shard) are an ordered warp-parallel matrix product, O(log) instead of O(blocks). */
    if (warp == 0) {
        const int b = blockIdx.x;
        int b0 = 0; /* first block whose product is applied */
        for (int base = b - 1; base >= 0; base -= 32) {
            const int q = base - lane;
            const unsigned m = __ballot_sync(0xffffffffu, q >= 0 && __ldcg(&A.block_reset[q]) != 0);
            if (m) {
                b0 = base - (__ffs(m) - 1);
                break;
            }
        }
        double v[4] = {0.25, 0.25, 0.25, 0.25};
        if (b - b0 <= 2) {
            if (lane == 0) {
                for (int q = b0; q < b; q++) {
                    double T[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) q * 16 + i]);
                    vec_mat(v, T);
                    vec_normalize(v);
                }
            }
        } else {
            for (int g0 = b0; g0 < b; g0 += 32) {
                double T[16];
                const int q = g0 + lane;
                if (q < b) {
#pragma unroll
                    for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) q * 16 + i]);
                } else {
                    mat_identity(T);
                }
                if (b - g0 <= 8) warp_scan_prefix<8>(T, lane); else warp_scan_prefix<32>(T, lane);
                const int last = b - g0 <= 8 ? 7 : 31;
#pragma unroll
                for (int i = 0; i < 16; i++) T[i] = __shfl_sync(0xffffffffu, T[i], last);
                vec_mat(v, T);
                vec_normalize(v);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) blk_vec[i] = v[i];
        }
    } else if (warp == 1) {
        const int b = blockIdx.x, nb = gridDim.x;
        int b1 = nb - 1; /* last block whose product is applied */
        for (int base = b + 1; base < nb; base += 32) {
            const int q = base + lane;
            const unsigned m = __ballot_sync(0xffffffffu, q < nb && __ldcg(&A.block_reset[q]) != 0);
            if (m) {
                b1 = base + (__ffs(m) - 1);
                break;
            }
        }
        double u[4] = {1.0, 1.0, 1.0, 1.0};
        if (b1 - b <= 2) {
            if (lane == 0) {
                for (int q = b1; q > b; q--) {
                    double T[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) q * 16 + i]);
                    mat_vec(T, u);
                    vec_normalize(u);
                }
            }
        } else {
            /* groups of 32 blocks, the farthest group first: u <- (T_g0 * ... * T_g0+31) * u */
            const int n_groups = (b1 - b + 31) / 32;
            for (int g = n_groups - 1; g >= 0; g--) {
                const int g0 = b + 1 + 32 * g, cnt = min(32, b1 - g0 + 1);
                double T[16];
                if (lane < cnt) {
#pragma unroll
                    for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) (g0 + lane) * 16 + i]);
                } else {
                    mat_identity(T);
                }
                if (cnt <= 8) warp_scan_prefix<8>(T, lane); else warp_scan_prefix<32>(T, lane);
                const int last = cnt <= 8 ? 7 : 31;
#pragma unroll
                for (int i = 0; i < 16; i++) T[i] = __shfl_sync(0xffffffffu, T[i], last);
                mat_vec(T, u);
                vec_normalize(u);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; i++) blk_vec[4 + i] = u[i];
        }
    }
    __syncthreads();

    double v_in[4], u_in[4];
    {
        double T[16];
#pragma unroll
        for (int i = 0; i < 4; i++) v_in[i] = blk_vec[i];
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = warp_pre[warp * 16 + i];
        vec_mat(v_in, T);
        vec_normalize(v_in);
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = acc[(size_t) i * LD + tid];
        vec_mat(v_in, T);
        vec_normalize(v_in);
#pragma unroll
        for (int i = 0; i < 4; i++) u_in[i] = blk_vec[4 + i];
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = warp_suf[warp * 16 + i];
        mat_vec(T, u_in);
        vec_normalize(u_in);
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = acc[(size_t) (16 + i) * LD + tid];
        mat_vec(T, u_in);
        vec_normalize(u_in);
    }
    __syncthreads(); /* the scan stash has been read: the area holds the gather tiles again */

    /* =========================== phase C1: forward inside the segment ======================================== */
    double loglik = 0.0;
    double f[4] = {v_in[0], v_in[1], v_in[2], v_in[3]};
    {
        /* sum_i log(c_i) = log(prod c_i): the product is carried as mantissa x 2^exponent (exact rescaling), one log
         * per segment instead of one per window */
        double cprod = 1.0;
        int cexp = 0;
        uint32_t w0 = len > 0 ? __ldg(wk) : 0u;
        uint32_t w1 = len > 1 ? __ldg(wk + cap) : 0u;
        double Mc[16], raw[16];
        if (kmax > 0) {
            coop_fetch(A.tabM, HFG_KEY_ID(w0), lane, raw);
            coop_deliver(stage, lane, raw, Mc);
        }
#pragma unroll 1
        for (int k = 0; k < kmax; k++) {
            const uint32_t w2 = k + 2 < len ? __ldg(wk + (size_t) (k + 2) * cap) : 0u;
            if (k + 1 < kmax) coop_fetch(A.tabM, HFG_KEY_ID(w1), lane, raw);
            if (k < len) {
                const bool start = (w0 & HFG_KEY_CHUNK_START) != 0;
                double fn[4];
                if (start) {
                    /* EM_fillFirstColumnForward: f[0][s] = e * start probability = any row of the rank-1 matrix */
#pragma unroll
                    for (int s = 0; s < 4; s++) fn[s] = Mc[s];
                } else {
                    /* f[i][s] = sum_pre f[i-1][pre] * (tProb * eProb), preState ascending (hmm.c:386-408) */
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        double a = 0.0;
#pragma unroll
                        for (int pre = 0; pre < 4; pre++) a += f[pre] * Mc[pre * 4 + s];
                        fn[s] = a;
                    }
                }
                const double c = ((fn[0] + fn[1]) + fn[2]) + fn[3];
                if (!start && c < 1e-50) uf_flag = 1; /* "scale is very low" (hmm.c:412-415) */
                /* f^ = f / c as one correctly rounded reciprocal and four products (<= 1 ulp from the four divisions of
                 * hmm.c:417-419) */
                const double rc = __drcp_rn(c);
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    f[s] = fn[s] * rc;
                    A.scrFT[((size_t) k * 4 + s) * cap + j] = f[s]; /* segment-transposed: read back by this thread in C2 */
                }
                cprod *= c;
                {
                    const int hi = __double2hiint(cprod);
                    const int e = ((hi >> 20) & 0x7ff) - 1023;
                    cexp += e;
                    cprod = __hiloint2double(hi - (e << 20), __double2loint(cprod));
                }
            }
            if (k + 1 < kmax) coop_deliver(stage, lane, raw, Mc);
            w0 = w1;
            w1 = w2;
        }
        loglik = log(cprod) + (double) cexp * 0.6931471805599453;
        A.seg_loglik[j] = loglik;
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 4] = clock64(); /* thread 0's own C1 end (no barrier here) */

    /* =========================== phase C2: backward + decode ================================================= */
    if (!A.forward_only) {
        /* f[] holds f^ of the segment's last window.  b is a direction: the statistics need it only up to the per-window
         * normalisation  sum_{pre,s} f^_{i-1}[pre] M_i[pre][s] b_i[s] = 1, the decode only up to a positive factor.
         * All lanes of the warp walk k = kmax-1 .. 0 together (cooperative gather); a lane works while k < len and its
         * chunk start has not been passed. */
        double bh[4] = {1.0, 1.0, 1.0, 1.0};
        bool done = len == 0;
        /* decode of one window: EM_getPosterior / EM_getMostProbableState (hmm.c:671-692), first maximum
         * (common.c:292-303); a common positive factor of b does not change the order */
        auto decode = [&](const double (&fw)[4], const double (&bw)[4], int gi) {
            double g[4];
#pragma unroll
            for (int s = 0; s < 4; s++) g[s] = fw[s] * bw[s];
            if (A.posteriors) {
                const double tot = ((g[0] + g[1]) + g[2]) + g[3];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    g[s] /= tot;
                    A.posteriors[(size_t) gi * 4 + s] = g[s];
                }
            }
            int best = 0;
#pragma unroll
            for (int s = 1; s < 4; s++)
                if (g[best] < g[s]) best = s;
            A.labels[gi] = (int8_t) best;
        };
        if (len > 0) {
            const uint32_t wl = __ldg(wk + (size_t) (len - 1) * cap);
            if (wl & HFG_KEY_CHUNK_END) {
                /* EM_fillLastColumnBackward (hmm.c:452-467): b = terminationProb / scale */
                const double *rt = rtab + (size_t) HFG_OBS_REGION(__ldg(&A.kdesc[HFG_KEY_ID(wl)])) * rt_stride;
#pragma unroll
                for (int s = 0; s < 4; s++) bh[s] = rt[RT_TERM + s] * HFG_INV_TERM;
            } else {
#pragma unroll
                for (int s = 0; s < 4; s++) bh[s] = u_in[s];
            }
            decode(f, bh, seg_first + len - 1); /* the segment's last window; every other window is decoded at the end of
                                                   the step that produces its b */
        }
        uint32_t w0 = kmax - 1 < len && kmax > 0 ? __ldg(wk + (size_t) (kmax - 1) * cap) : 0u;
        uint32_t w1 = kmax - 2 < len && kmax > 1 ? __ldg(wk + (size_t) (kmax - 2) * cap) : 0u;
        double Mc[16], raw[16];
        if (kmax > 0) {
            coop_fetch(A.tabM, HFG_KEY_ID(w0), lane, raw);
            coop_deliver(stage, lane, raw, Mc);
        }
        /* f^ of window k-1 is fetched one step ahead of its use as well (it feeds the normalisation, the head of the
         * step's dependent chain) */
        double fq[4] = {0.0, 0.0, 0.0, 0.0};
        if (kmax >= 2 && kmax - 2 < len) {
#pragma unroll
            for (int s = 0; s < 4; s++) fq[s] = A.scrFT[((size_t) (kmax - 2) * 4 + s) * cap + j];
        }
#pragma unroll 1
        for (int k = kmax - 1; k >= 0; k--) {
            const uint32_t w2 = (k >= 2 && k - 2 < len) ? __ldg(wk + (size_t) (k - 2) * cap) : 0u;
            if (k >= 1) coop_fetch(A.tabM, HFG_KEY_ID(w1), lane, raw);
            double fq_next[4] = {0.0, 0.0, 0.0, 0.0}; /* f^ of window k-2, for the next step */
            if (k >= 2 && k - 2 < len) {
#pragma unroll
                for (int s = 0; s < 4; s++) fq_next[s] = A.scrFT[((size_t) (k - 2) * 4 + s) * cap + j];
            }
            if (k < len && !done) {
                const int gi = seg_first + k;
                if (w0 & HFG_KEY_CHUNK_START) {
                    done = true; /* first window of a chunk: nothing to the left */
                } else {
                    /* f^ of the previous window (last window of the previous segment == the entering message) */
                    double fp[4];
                    if (k > 0) {
#pragma unroll
                        for (int s = 0; s < 4; s++) fp[s] = fq[s];
                    } else {
#pragma unroll
                        for (int s = 0; s < 4; s++) fp[s] = v_in[s];
                    }
                    /* b[i-1][pre] = sum_s tProb*eProb*b[i][s] (hmm.c:493-520), then the normalisation */
                    double bn[4];
#pragma unroll
                    for (int pre = 0; pre < 4; pre++)
                        bn[pre] = ((Mc[pre * 4] * bh[0] + Mc[pre * 4 + 1] * bh[1]) + Mc[pre * 4 + 2] * bh[2]) +
                                  Mc[pre * 4 + 3] * bh[3];
                    const double dot = ((fp[0] * bn[0] + fp[1] * bn[1]) + fp[2] * bn[2]) + fp[3] * bn[3];
                    const double r = 1.0 / dot;
                    double bs[4];
#pragma unroll
                    for (int s = 0; s < 4; s++) bs[s] = bh[s] * r;
                    /* the window's record for the statistics: (f^ of the previous window, normalised b) */
                    store_vec4(A.scrXB + (size_t) gi * 8, fp);
                    store_vec4(A.scrXB + (size_t) gi * 8 + 4, bs);
#pragma unroll
                    for (int s = 0; s < 4; s++) bh[s] = bn[s] * r;
                    if (k > 0) decode(fp, bh, gi - 1);
                }
            }
            if (k >= 1) coop_deliver(stage, lane, raw, Mc);
#pragma unroll
            for (int s = 0; s < 4; s++) fq[s] = fq_next[s];
            w0 = w1;
            w1 = w2;
        }
    }
    __syncthreads(); /* gather tiles done: the area becomes the statistics columns */
    if (A.labels_host != NULL && !A.forward_only && warp >= WARPS - 2) {
        /* The labels of this block's windows (one contiguous range: segments are in genome order) are final.  The last two
         * warps -- the statistics tiles fill the block from thread 0 up -- stream them into the caller's page-locked buffer
         * over PCIe with 16-byte stores while the other warps run phase S. */
        const int s0 = blockIdx.x * THREADS;
        if (s0 < A.n_seg) {
            const long long w_begin = A.seg_start[s0];
            const long long w_end = s0 + THREADS < A.n_seg ? (long long) A.seg_start[s0 + THREADS] : (long long) A.n_windows;
            const int t = (warp - (WARPS - 2)) * 32 + lane; /* 0..63 */
            const long long a16 = (w_begin + 15) & ~15LL, b16 = w_end & ~15LL;
            if (a16 < b16) {
                for (long long o = a16 + 16LL * t; o < b16; o += 16 * 64)
                    *reinterpret_cast<int4 *>(A.labels_host + o) = __ldcg(reinterpret_cast<const int4 *>(A.labels + o));
                for (long long o = w_begin + t; o < a16; o += 64) A.labels_host[o] = __ldcg(A.labels + o);
                for (long long o = b16 + t; o < w_end; o += 64) A.labels_host[o] = __ldcg(A.labels + o);
            } else {
                for (long long o = w_begin + t; o < w_end; o += 64) A.labels_host[o] = __ldcg(A.labels + o);
            }
        }
    }
    if (uf_flag) atomicOr(A.err_flags, 1);
    grid.sync(); /* f^ and b of every window are in place */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 5] = clock64();
    const unsigned tok = fence_token();

    /* =========================== phases S and D, region by region ============================================ */
    /* per-thread statistics live in column tid of acc: rows 0..15 transition counts, 16..17 truncated exponential,
     * 18+3g.. (meanNum, den, varNum) of Gaussian component g, last row the log-likelihood */
    double *col = acc + tid;
    /* every block takes a contiguous, equal share of the tiles (they are sorted by region): a block meets one or two
     * regions, not all R */
    const long long n_tiles_all = A.forward_only ? 0 : __ldg(&A.region_tile_begin[HFG_MAX_REGIONS]);
    const int blk_t0 = (int) (n_tiles_all * blockIdx.x / gridDim.x), blk_t1 = (int) (n_tiles_all * (blockIdx.x + 1) / gridDim.x);
    for (int r = 0; r < R; r++) {
        const int t_begin = max(blk_t0, __ldg(&A.region_tile_begin[r])), t_end = min(blk_t1, __ldg(&A.region_tile_begin[r + 1]));
        if (r > 0 && t_begin >= t_end) { /* (block-uniform) none of this region's tiles here; region 0 carries the log-likelihood */
            for (int q = tid; q < NSTAT; q += THREADS) A.partials[((size_t) r * NSTAT + q) * gridDim.x + blockIdx.x] = 0.0;
            continue;
        }
        for (int q = 0; q < NSTAT; q++) col[(size_t) q * LD] = 0.0;
        if (r == 0) col[(size_t) (NSTAT - 1) * LD] = loglik;
        for (int t = t_begin + tid; t < t_end; t += THREADS) {
            const int p = __ldg(&A.tile_key[t]), lb = __ldg(&A.tile_begin[t]), ln = __ldg(&A.tile_cnt[t]);
            /* S[pre][s] = sum over the tile's windows of f^_{i-1}[pre] * b_i[s]; two windows' records in flight */
            double S[16];
#pragma unroll
            for (int i = 0; i < 16; i++) S[i] = 0.0;
            for (int i0 = 0; i0 < ln; i0 += 2) {
                double fp[2][4], bb[2][4];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int gi = __ldg(&A.klist[lb + min(i0 + u, ln - 1)]);
                    load_vec4_cg(A.scrXB + (size_t) gi * 8, tok, fp[u]);
                    load_vec4_cg(A.scrXB + (size_t) gi * 8 + 4, tok, bb[u]);
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (i0 + u < ln) {
#pragma unroll
                        for (int pre = 0; pre < 4; pre++)
#pragma unroll
                            for (int s = 0; s < 4; s++) S[pre * 4 + s] = fma(fp[u][pre], bb[u][s], S[pre * 4 + s]);
                    }
                }
            }
            Win w = decode_word(__ldg(&A.kdesc[p]), A.beta0);
            if (w.edge) {
                w.beta = A.kbeta[3 * (size_t) p];
                w.rb = A.kbeta[3 * (size_t) p + 1];
                w.sq = A.kbeta[3 * (size_t) p + 2];
            }
            const double *rt = rtab + (size_t) w.region * rt_stride;
            double M[16];
            load_mat(A.tabM + (size_t) p * 16, M);
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double xi[4]; /* pooled pair counts into state s, by preState (count / terminationProb, hmm.c:613-614) */
                double hh[4]; /* the same without the emission factor (multi-component statistics) */
#pragma unroll
                for (int pre = 0; pre < 4; pre++) {
                    xi[pre] = S[pre * 4 + s] * M[pre * 4 + s];
                    hh[pre] = S[pre * 4 + s] * trans_prob(rt, w, pre, s);
                }
#pragma unroll
                for (int pre = 0; pre < 4; pre++) col[(pre * 4 + s) * LD] += xi[pre]; /* hmm_utils.c:2010-2015 */
                if constexpr (NB) {
                    /* hmm.c:615-617: the pair mass goes into the state's histogram over x; one tile = one key = one x */
                    A.nb_tile_col[(size_t) t * 4 + s] = ((xi[0] + xi[1]) + xi[2]) + xi[3];
                } else if (!A.is_gauss[s]) {
                    /* TruncExponential_updateEstimator (hmm_utils.c:1027-1034) */
                    const double sum = ((xi[0] + xi[1]) + xi[2]) + xi[3];
                    col[16 * LD] += sum * w.x;
                    col[17 * LD] += sum;
                } else {
                    /* Gaussian_updateEstimator (hmm_utils.c:812-839), grouped by (state, alpha) class: preStates that
                     * share alpha share x_adjusted, z and the component responsibilities */
                    const int n = A.ncomp[s], g0 = A.gbase[s];
                    if (n == 1) {
                        /* single component: the responsibility is 1, so (w*x_adj, w, w*z*z) with w = the pair count;
                         * summed over the four preStates, then one update of the thread's column */
                        const double mu = rt[RT_GAUSS + g0];
                        double s1 = 0.0, s0 = 0.0, s2 = 0.0;
#pragma unroll
                        for (int pre = 0; pre < 4; pre++) {
                            const double a = A.alpha[pre][s];
                            const double x_adj = (w.x - a * w.px) * A.inv_one_minus_alpha[pre][s];
                            const double z = (x_adj - mu) * (1.0 - a);
                            s1 += xi[pre] * x_adj;
                            s0 += xi[pre];
                            s2 += xi[pre] * z * z;
                        }
                        double *cg_ = col + (size_t) (18 + 3 * g0) * LD;
                        cg_[0] += s1;
                        cg_[LD] += s0;
                        cg_[2 * LD] += s2;
                    } else {
                        /* pool H = (f*t)*b/term over the preStates of each (state, alpha) class */
                        double Hc[4] = {0.0, 0.0, 0.0, 0.0}, ac[4] = {0.0, 0.0, 0.0, 0.0};
                        int nd = 0;
#pragma unroll
                        for (int pre = 0; pre < 4; pre++) {
                            if (A.first_pre_of_class[pre][s] != pre) continue;
                            double hsum = hh[pre];
#pragma unroll
                            for (int p2 = pre + 1; p2 < 4; p2++)
                                if (A.first_pre_of_class[p2][s] == pre) hsum += hh[p2];
#pragma unroll
                            for (int q = 0; q < 4; q++)
                                if (q == nd) {
                                    Hc[q] = hsum;
                                    ac[q] = A.alpha[pre][s];
                                }
                            nd++;
                        }
                        gauss_state_stats(rt, G, g0, n, nd, Hc[0], Hc[1], Hc[2], Hc[3], ac[0], ac[1], ac[2], ac[3], w.x,
                                          w.px, w.beta, w.rb, w.sq, col + (size_t) (18 + 3 * g0) * LD, LD, &nan_flag);
                    }
                }
            }
        }
        __syncthreads();
        /* deterministic block reduction: one warp per statistic row; each lane adds its THREADS/32 entries in index
         * order, then a fixed xor-shuffle tree -- the same association on every run */
        for (int q = warp; q < NSTAT; q += WARPS) {
            const double *cl = acc + (size_t) q * LD;
            double sum = 0.0;
#pragma unroll
            for (int t = lane; t < THREADS; t += 32) sum += cl[t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (lane == 0) A.partials[((size_t) r * NSTAT + q) * gridDim.x + blockIdx.x] = sum; /* [R][NSTAT][grid] */
        }
        __syncthreads();
    }
    if (nan_flag) atomicOr(A.err_flags, 2);

    /* =========================== phase D: grid reduction ===================================================== */
    {
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 6] = clock64();
        grid.sync();
        if (tid == 0) {
            A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 7] = clock64();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 8] = (long long) smid;
        }
        if (blockIdx.x == 0) {
            const int SD = (int) (sizeof(hfg_region_stats) / sizeof(double));
            const int nb = gridDim.x;
            /* four totals per warp at a time: the lanes add the blocks' partials with stride 32 (all loads of a round are
             * independent and in flight together), then a fixed xor-shuffle tree -- the same association on every run */
            {
                constexpr int QU = 4;
                const int NQ = R * NSTAT;
                for (int q0 = warp * QU; q0 < NQ; q0 += WARPS * QU) {
                    double sum[QU];
#pragma unroll
                    for (int u = 0; u < QU; u++) sum[u] = 0.0;
                    /* up to 5 x QU loads per lane issued before the first add (grids of up to 160 blocks in one go) */
                    for (int b0 = 0; b0 < nb; b0 += 160) {
                        double v[5][QU];
#pragma unroll
                        for (int i = 0; i < 5; i++) {
                            const int b = b0 + 32 * i + lane;
#pragma unroll
                            for (int u = 0; u < QU; u++)
                                v[i][u] = (b < nb && q0 + u < NQ) ? __ldcg(&A.partials[(size_t) (q0 + u) * nb + b]) : 0.0;
                        }
#pragma unroll
                        for (int i = 0; i < 5; i++)
#pragma unroll
                            for (int u = 0; u < QU; u++) sum[u] += v[i][u];
                    }
#pragma unroll
                    for (int u = 0; u < QU; u++) {
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) sum[u] += __shfl_xor_sync(0xffffffffu, sum[u], off);
                        if (lane == 0 && q0 + u < NQ) acc[q0 + u] = sum[u]; /* reuse shared memory: [R][NSTAT] totals */
                    }
                }
            }
            __syncthreads();
            if (tid == 0) A.phase_clock[11] = clock64(); /* block 0: grid totals in shared memory */
            /* the hfg_region_stats layout (include/hfg.h), one output element per thread: no zero-fill, no read-back */
            for (int q = tid; q < R * SD; q += THREADS) {
                const int r = q / SD, i = q % SD;
                const double *tot = acc + (size_t) r * NSTAT;
                const int MC = HFG_MAX_COMPS, BL = HFG_NS * HFG_MAX_COMPS;
                double v = 0.0;
                if (i < 18) {
                    v = tot[i]; /* trans_count[pre][s], lambda_num, lambda_den */
                } else {
                    const int kind = (i - 18) / BL, at = (i - 18) % BL, s = at / MC, cc = at % MC;
                    if (A.is_gauss[s] && cc < A.ncomp[s]) {
                        const int g = A.gbase[s] + cc;
                        if (kind == 0) v = tot[18 + 3 * g];          /* mean_num */
                        else if (kind == 2) v = tot[18 + 3 * g + 2]; /* var_num */
                        else if (kind < 5) v = tot[18 + 3 * g + 1];  /* mean_den = var_den = weight_num (same addends) */
                        else {
                            /* weight_den[s][c'] = sum over the components of s
                             * (ParameterEstimator_incrementDenominatorForAllComps) */
                            for (int c2 = 0; c2 < A.ncomp[s]; c2++) v += tot[18 + 3 * (A.gbase[s] + c2) + 1];
                        }
                    }
                }
                A.out[q] = v;
            }
            if (tid == 0) A.out[(size_t) R * SD] = acc[NSTAT - 1]; /* log-likelihood (kept in region 0's row) */
            if (tid == 0) A.out[(size_t) R * SD + 1] = (double) __ldcg(A.err_flags);

            /* ---- fused collective: sum the block over all ranks through peer memory ------------------------------
             * Each rank stores its vector into slot [rank] of every peer's mailbox (plain coalesced stores over
             * NVLink), publishes an arrival counter with system scope, waits for the N counters of its own mailbox, and
             * adds the N vectors in rank order -- the same association on every rank, so all ranks end with the same
             * bits and run the same host M-step.  Two epochs alternate so that a fast rank can never overwrite a slot
             * a slow rank is still reading. */
            if (A.n_ranks > 1) {
                __syncthreads();
                const int n = A.out_doubles, N = A.n_ranks;
                const unsigned long long e = *A.epoch + 1;
                const size_t box = (size_t) (e & 1) * HFG_MAX_PEERS * n;
                const size_t cnt_base = (size_t) 2 * HFG_MAX_PEERS * n; /* counters live behind the slots (as u64) */
                for (int p = 0; p < N; p++) {
                    double *dst = A.peer_box[p] + box + (size_t) A.rank * n;
                    for (int q = tid; q < n; q += THREADS) dst[q] = A.out[q];
                }
                __threadfence_system();
                __syncthreads();
                if (tid < N) {
                    volatile unsigned long long *c =
                        (volatile unsigned long long *) (A.peer_box[tid] + cnt_base) + (e & 1) * HFG_MAX_PEERS + A.rank;
                    *c = e;
                    __threadfence_system();
                    /* wait for sender `tid` in the local mailbox (bounded: a lost peer becomes an error, not a hang) */
                    volatile unsigned long long *mine =
                        (volatile unsigned long long *) (A.peer_box[A.rank] + cnt_base) + (e & 1) * HFG_MAX_PEERS + tid;
                    const long long t0 = clock64();
                    while (*mine < e) {
                        if (clock64() - t0 > 4000000000LL) { /* ~2 s */
                            atomicOr(A.err_flags, 4);
                            break;
                        }
                    }
                }
                __threadfence_system();
                __syncthreads();
                const double *in = A.peer_box[A.rank] + box;
                for (int q = tid; q < n; q += THREADS) {
                    double sum = 0.0;
                    if (q == n - 1) { /* error flags: bitwise OR over ranks */
                        int f = 0;
                        for (int p = 0; p < N; p++) f |= (int) __ldcv(in + (size_t) p * n + q);
                        sum = (double) f;
                    } else {
                        for (int p = 0; p < N; p++) sum += __ldcv(in + (size_t) p * n + q);
                    }
                    A.out[q] = sum;
                }
                __syncthreads();
                if (tid == 0) {
                    if (__ldcg(A.err_flags) & 4) A.out[n - 1] = (double) ((int) A.out[n - 1] | 4);
                    *A.epoch = e;
                }
            }

            /* ---- device-resident EM: M-step of every region (one thread each), bookkeeping ------------------------ */
            if (tid == 0) A.phase_clock[9] = clock64(); /* block 0: totals reduced (and exchanged) */
            if (A.em_mode) {
                __syncthreads();
                const int flags = (int) A.out[(size_t) R * SD + 1];
                int settled = 1;
                if (A.em_mode == 1 && flags == 0) {
                    /* Three warps per region -- Gaussian parameters, rate fit (warp-parallel), transition rows: disjoint
                     * parameters -- on a shared-memory copy of the region's parameters and statistics (the M-step is a
                     * chain of dependent loads, divisions and stores).  Every lane of a warp computes and stores the
                     * same values. */
                    constexpr int PD = (int) (sizeof(hfg_region_params) / sizeof(double));
                    constexpr int GROUPS = WARPS / 3;
                    const int grp = warp / 3, role = warp % 3;
                    for (int r0 = 0; r0 < R; r0 += GROUPS) {
                        const int r = r0 + grp;
                        const bool mine = grp < GROUPS && r < R;
                        double *mp = acc + (size_t) grp * (PD + SD);
                        double *gp = reinterpret_cast<double *>(&A.em_params[mine ? r : 0]);
                        if (mine) {
                            for (int i = role * 32 + lane; i < PD; i += 96) mp[i] = gp[i];
                            for (int i = role * 32 + lane; i < SD; i += 96) mp[PD + i] = A.out[(size_t) r * SD + i];
                        }
                        __syncthreads();
                        if (mine) {
                            hfg_region_params *p = reinterpret_cast<hfg_region_params *>(mp);
                            const hfg_region_stats *st = reinterpret_cast<const hfg_region_stats *>(mp + PD);
                            if (role == 0) settled &= hfg_mstep_rate(A.model_type, p, st, A.em_tol);
                            else if (role == 1) settled &= hfg_mstep_gauss(A.model_type, A.ncomp, p, st, A.em_tol);
                            else settled &= hfg_mstep_trans(p, st, A.em_tol);
                        }
                        __syncthreads();
                        if (mine) {
                            /* the truncation point follows the NEW Hap mean, after the fit used the old one */
                            if (role == 0 && lane == 0 && A.model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN)
                                mp[1] = reinterpret_cast<hfg_region_params *>(mp)->mean[HFG_STATE_HAP][0] * TRUNC_POINT_FRACTION;
                            __syncwarp();
                        }
                        __syncthreads();
                        if (mine)
                            for (int i = role * 32 + lane; i < PD; i += 96) gp[i] = mp[i];
                        __syncthreads();
                    }
                }
                const int all_settled = __syncthreads_and(settled);
                if (tid == 0) {
                    const int k = A.em_state[1];
                    if (k < A.em_max_logliks) A.em_logliks[k] = A.out[(size_t) R * SD];
                    A.em_state[1] = k + 1;
                    if (flags) {
                        A.em_state[2] |= flags;
                        A.em_state[0] = 1;
                    } else if (A.em_mode == 1 && all_settled) {
                        A.em_state[3] = 1;
                        A.em_state[0] = 1;
                    }
                    A.phase_clock[10] = clock64(); /* block 0: M-step done */
                }
            }
            /* results straight into the caller-visible pinned block, error flags cleared for the next launch */
            __syncthreads();
            if (A.out_host)
                for (int q = tid; q < A.out_doubles; q += THREADS) A.out_host[q] = A.out[q];
            if (tid == 0) *A.err_flags = 0;
        }
    }
}
