/*
 * hfg_nb_dev.cuh -- the negative-binomial model on the device.
 *
 * Device-resident EM loop (hfg_em_*, hfg_run_em): the formulas of the host half (hfg_nb.c) run in fp64 on the device so that
 * successive iterations are back-to-back kernels: the pmf (NegativeBinomial_getComponentProbs, hmm_utils.c:494-515) per
 * observation key in the key-matrix phase; after the statistics phase the histogram of the pair mass over the coverage value
 * (hmm.c:615-617, count_data.c:56-64), folded from the per-tile column sums by the WHOLE grid, one (region, bin) per CTA at a
 * time (grid_fold_histogram); and in the last CTA's tail NegativeBinomial_updateEstimator per non-empty bin
 * (hmm_utils.c:536-563, psi(r + x) by the recurrence of :392-406: tail_estimators) and the M-step (hfg_nb_mstep_inl.h, the
 * host's code).  Differences from the host path: CUDA's lgamma / exp / log instead of glibc's, psi(r + x) - psi(r) as a
 * running sum of reciprocals without the long-double digamma(r) the host adds and subtracts again, bins summed by a fixed
 * tree instead of in ascending order: rounding only (tests/test_gpu_nb.py compares the two loops).
 * The blocking calls (hfg_em_iteration) use the grid-wide fold as well and read the histogram back; their emission table
 * (libm) and estimator update (long-double digamma) stay on the host: the bits of the reference.
 */
#pragma once

#define HFG_HD static __device__
#include "hfg_nb_mstep_inl.h"
#undef HFG_HD

#define HFG_NB_DEV_BINS 250  /* HFG_NB_BINS: x = 250 is folded into bin 249 (count_data.c:56-64) */
#define HFG_NB_DEV_X 251     /* HFG_NB_TABLE_X */

#define HFG_NB_TAIL_DOUBLES(NP) ((4 + 1 + (NP)) * 256 + 7 * (NP)) /* histogram, lgamma(x + 1), pmf of NP components, their constants */

namespace hfgnb {

struct Comp {
    double r, w, lgamma_r, r_log_theta, log_1m_theta, bt;
};
__device__ __forceinline__ Comp comp_setup(const hfg_region_params &p, int s, int k) {
    Comp c;
    const double theta = p.mean[s][k], lt = log(theta);
    c.r = -1 * p.var[s][k] / lt; /* NegativeBinomial_getR, hmm_utils.c:455-458 */
    c.w = p.weight[s][k];
    c.lgamma_r = lgamma(c.r);
    c.r_log_theta = c.r * lt;
    c.log_1m_theta = log(1 - theta);
    c.bt = -1 * theta / (1 - theta) - 1 / lt; /* hmm_utils.c:545 */
    return c;
}
/* weighted pmf of one component at x, floored at 1e-40 */
__device__ __forceinline__ double comp_prob(const Comp &c, int x, double lg_x1, int *nan) {
    double v = c.w * exp(lgamma(c.r + x) - c.lgamma_r - lg_x1 + c.r_log_theta + (double) x * c.log_1m_theta);
    if (v != v) *nan = 1;
    return v < 1e-40 ? 1e-40 : v;
}
/* emission of (region parameters p, state s) at coverage x: the sum over the components in component order */
__device__ __forceinline__ double state_prob(const hfg_region_params &p, int s, int n, int x, double lg_x1, int *nan) {
    double tot = 0.0;
#pragma unroll 1
    for (int k = 0; k < n; k++) {
        const Comp c = comp_setup(p, s, k);
        tot += comp_prob(c, x, lg_x1, nan);
    }
    return tot;
}

/* Histogram of the pair mass over the coverage value (hmm.c:615-617, count_data.c:56-64), folded by the WHOLE grid after the
 * statistics phase (the caller has passed a grid barrier: every tile's column sums are in tile_col): bin (region, x) = the sum
 * over the tiles whose key has coverage x, by state.  The (region, bin) pairs are dealt round-robin over the CTAs; a CTA adds a
 * bin's tiles with one 32-byte load per thread and tile (all loads in flight), then a fixed tree: lanes, then warps in
 * order -- the same bits on every run.  `red` = [WARPS][4] doubles of shared memory. */
template <int THREADS>
__device__ __forceinline__ void grid_fold_histogram(int R, const double *tile_col, const int32_t *bin_begin, const int32_t *bin_tiles,
                                                    double *hist_g, double *red) {
    constexpr int WARPS = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int pair = blockIdx.x; pair < R * HFG_NB_DEV_BINS; pair += gridDim.x) {
        const int b0 = __ldg(bin_begin + pair), b1 = __ldg(bin_begin + pair + 1);
        const int reg = pair / HFG_NB_DEV_BINS, x = pair % HFG_NB_DEV_BINS;
        if (b0 == b1) { /* (CTA-uniform) an empty bin */
            if (tid < 4) hist_g[((size_t) reg * 4 + tid) * 256 + x] = 0.0;
            continue;
        }
        double h[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i0 = b0 + tid; i0 < b1; i0 += 4 * THREADS) {
            int tl[4];
#pragma unroll
            for (int u = 0; u < 4; u++) tl[u] = i0 + u * THREADS < b1 ? __ldg(bin_tiles + i0 + u * THREADS) : -1;
            double v[4][4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (tl[u] >= 0) {
                    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                                 : "=d"(v[u][0]), "=d"(v[u][1]), "=d"(v[u][2]), "=d"(v[u][3])
                                 : "l"(tile_col + (size_t) tl[u] * 4)
                                 : "memory");
                } else {
                    v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int k = 0; k < 4; k++) h[k] += v[u][k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) h[k] += __shfl_xor_sync(0xffffffffu, h[k], off);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) red[warp * 4 + k] = h[k];
        }
        __syncthreads();
        if (tid < 4) {
            double sum = 0.0;
#pragma unroll 4
            for (int wv = 0; wv < WARPS; wv++) sum += red[wv * 4 + tid];
            hist_g[((size_t) reg * 4 + tid) * 256 + x] = sum;
        }
        __syncthreads();
    }
}

/* Tail of the device-resident loop, run by the NW worker threads of the last CTA: histogram -> estimator sums, written into
 * the hfg_region_stats block `out` ([R][SD] doubles).  `scr` = shared memory, HFG_NB_TAIL_DOUBLES doubles.  `sync` = the
 * workers' barrier.  Returns a NaN flag (thread-local; OR it over the threads).
 * Two parallel steps per region instead of a loop over states and components: (1) the weighted pmf of every (component, bin)
 * with mass in its state's histogram, dealt over all threads; (2) one WARP per (state, component): eight consecutive bins
 * per lane, psi(r + x) - psi(r) as a running sum of reciprocals (lane-local, then a shuffle scan of the lanes' totals), the
 * four estimator sums in bin order per lane and by a fixed shuffle tree over the lanes: no cross-warp reduction. */
template <typename Sync>
__device__ __forceinline__ int tail_estimators(const hfg_region_params *params, int R, const int32_t *ncomp, const double *hist_g,
                                               const double *lg_x1_g, double *out, double *scr, int tid, int NW, Sync sync,
                                               long long *clk) {
    const int SD = (int) (sizeof(hfg_region_stats) / sizeof(double));
    const int lane = tid & 31, warp = tid >> 5, NWARP = NW / 32;
    int nan = 0;
    const int pb[HFG_NS + 1] = {0, ncomp[0], ncomp[0] + ncomp[1], ncomp[0] + ncomp[1] + ncomp[2], ncomp[0] + ncomp[1] + ncomp[2] + ncomp[3]};
    const int NP = pb[HFG_NS]; /* (state, component) pairs, state-major */
    double *hist = scr;                /* [4][256] of the current region */
    double *lgx = hist + 4 * 256;      /* [256] lgamma(x + 1) */
    double *probs = lgx + 256;         /* [NP][256] weighted pmf of every component */
    Comp *cs = reinterpret_cast<Comp *>(probs + (size_t) NP * 256); /* [NP] constants of every component of the region */
    double *wn = reinterpret_cast<double *>(cs + NP);               /* [NP] weight numerators, for the states' denominators */
    if (tid < 256) lgx[tid] = tid < HFG_NB_DEV_X ? __ldg(lg_x1_g + tid) : 0.0;
    auto state_of = [&](int q) { return q < pb[1] ? 0 : (q < pb[2] ? 1 : (q < pb[3] ? 2 : 3)); };
    for (int reg = 0; reg < R; reg++) {
        const hfg_region_params &p = params[reg];
        hfg_region_stats *st = reinterpret_cast<hfg_region_stats *>(out + (size_t) reg * SD);
        if (tid < NP) {
            const int sq = state_of(tid);
            cs[tid] = comp_setup(p, sq, tid - pb[sq]);
        }
        /* histogram of this region: folded by the whole grid before the tail (grid_fold_histogram) */
        for (int i = tid; i < 4 * 256; i += NW) hist[i] = (i & 255) < HFG_NB_DEV_BINS ? __ldcg(hist_g + (size_t) reg * 4 * 256 + i) : 0.0;
        sync();
        if (clk && tid == 0 && reg == 0) clk[12] = clock64();
        /* (1) the pmf wherever the state's histogram has mass (NegativeBinomial_updateEstimator skips the empty bins) */
        for (int i = tid; i < NP * 256; i += NW) {
            const int q = i >> 8, x = i & 255, sq = state_of(q);
            if (x < HFG_NB_DEV_BINS && 0 < hist[sq * 256 + x]) probs[i] = comp_prob(cs[q], x, lgx[x], &nan);
        }
        sync();
        if (clk && tid == 0 && reg == 0) clk[13] = clock64();
        /* (2) one warp per (state, component) */
        for (int q = warp; q < NP; q += NWARP) {
            const int sq = state_of(q), c = q - pb[sq], nc = ncomp[sq];
            const Comp cc = cs[q];
            /* running sum of 1 / (r + j), j < x: the lanes before this one (shuffle scan of the lanes' totals), then the lane's
             * own bins as it goes.  The loops are NOT unrolled: this code runs once per launch on cold instruction caches, where
             * every distinct instruction costs a fetch -- short loops beat straight-line code here. */
            double run = 0.0;
#pragma unroll 1
            for (int u = 0; u < 8; u++) run += 1.0 / (cc.r + (8 * lane + u));
            double before = run; /* inclusive scan of the lanes' totals, made exclusive below */
#pragma unroll 1
            for (int off = 1; off < 32; off <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, before, off);
                if (lane >= off) before += t;
            }
            double psi = before - run; /* psi(r + x) - psi(r) at the lane's first bin */
            double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int u = 0; u < 8; u++) {
                const int x = 8 * lane + u;
                const double mass = x < HFG_NB_DEV_BINS ? hist[sq * 256 + x] : 0.0;
                if (0 < mass) {
                    double total = 0.0;
#pragma unroll 1
                    for (int c2 = 0; c2 < nc; c2++) total += probs[(pb[sq] + c2) * 256 + x];
                    const double w = mass * probs[q * 256 + x] / total;
                    const double delta = cc.r * psi;
                    v[0] += w * delta;                           /* var_num (lambda estimator) */
                    v[1] += w;                                   /* var_den, weight_num */
                    v[2] += w * delta * cc.bt;                   /* mean_num (theta estimator) */
                    v[3] += w * delta * cc.bt + w * (x - delta); /* mean_den */
                }
                psi += 1.0 / (cc.r + x);
            }
#pragma unroll 1
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
            }
            if (lane == 0) {
                wn[q] = v[1];
                st->var_num[sq][c] = v[0];
                st->var_den[sq][c] = v[1];
                st->weight_num[sq][c] = v[1];
                st->mean_num[sq][c] = v[2];
                st->mean_den[sq][c] = v[3];
            }
        }
        sync();
        if (clk && tid == 0 && reg == 0) clk[14] = clock64();
        /* every component's weight estimator has the mass of all components of the state as its denominator */
        if (tid < HFG_NS) {
            double den = 0.0;
            for (int c = 0; c < ncomp[tid]; c++) den += wn[pb[tid] + c];
            for (int c = 0; c < ncomp[tid]; c++) st->weight_den[tid][c] = den;
        }
        sync();
        if (clk && tid == 0 && reg == 0) clk[15] = clock64();
    }
    return nan;
}

/* EmissionDistSeries_estimateParameters for MODEL_NEGATIVE_BINOMIAL (hfg_nb_mstep_region_inl's emission half) for one WARP:
 * one lane per component does the divisions, every lane adds the pooled sums in the serial routine's order (state, then
 * component, ascending), so the parameters come out with its bits.  `scr` = 128 doubles of warp-private shared memory. */
__device__ __forceinline__ int mstep_emis_warp(const int32_t *n_comps, hfg_region_params *p, const hfg_region_stats *st, double tol,
                                               double *scr, int lane) {
    int ok = 1, G = 0, gs[2] = {0, 0}, gc[2] = {0, 0};
    for (int s = 0; s < HFG_NS; s++) {
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g >= G && g < G + n_comps[s]) {
                gs[u] = s;
                gc[u] = g - G;
            }
        }
        G += n_comps[s];
    }
    for (int type = 0; type < 2; type++) { /* theta, then lambda: one pooled ("bound") estimate each */
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g < G) {
                const double f = type == 0 ? 1.0 : nb_lambda_coef(gs[u], gc[u]);
                scr[g] = (type == 0 ? st->mean_num[gs[u]][gc[u]] : st->var_num[gs[u]][gc[u]]) / f;
                scr[64 + g] = type == 0 ? st->mean_den[gs[u]][gc[u]] : st->var_den[gs[u]][gc[u]];
            }
        }
        __syncwarp();
        double num = 0.0, den = 0.0;
        for (int g = 0; g < G; g++) {
            num += scr[g];
            den += scr[64 + g];
        }
        __syncwarp();
        const double pooled = den == 0 ? 0.0 : num / den;
        if (!(NB_MIN_COUNT < den)) continue;
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g < G) {
                double *dst = type == 0 ? &p->mean[gs[u]][gc[u]] : &p->var[gs[u]][gc[u]];
                const double v = pooled * (type == 0 ? 1.0 : nb_lambda_coef(gs[u], gc[u]));
                ok &= nb_settled(*dst, v, tol);
                *dst = v;
            }
        }
    }
    for (int u = 0; u < 2; u++) {
        const int g = lane + 32 * u;
        if (g < G) {
            const double den = st->weight_den[gs[u]][gc[u]];
            if (NB_MIN_COUNT < den) {
                const double v = st->weight_num[gs[u]][gc[u]] / den;
                ok &= nb_settled(p->weight[gs[u]][gc[u]], v, tol);
                p->weight[gs[u]][gc[u]] = v;
            }
        }
    }
    __syncwarp();
    return __all_sync(0xffffffffu, ok);
}

}  // namespace hfgnb
