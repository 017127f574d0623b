/*
 * hfg_estep_v3.cuh -- the E-step kernel of libhfg, third generation (sm_100a, fp64, no tensor cores).
 *
 * Same job, data layout and phases as hfg_estep.cuh (EM_runForward + EM_runBackward + EM_updateEstimators + the label loop
 * of submodules/hmm/hmm.c:423-434, 535-545, 638-650, 715-737 for ALL chunks of one EM_runOneIterationForList call,
 * hmm.c:739-780, one persistent cooperative launch).  What two rounds of measurement (profiles/) say bounds this kernel
 * and what this version does about each:
 *   * instruction issue.  One thread per segment makes every instruction serve 32 windows; the four-lanes-per-segment
 *     kernel (hfg_estep_quad.cuh) serves 8 and executed 2.4x the warp instructions for the same work.  Thread per segment.
 *   * the trip of every 128-byte transfer matrix into registers, three times per window.  The first kernel gathered them
 *     from L1/L2 through LDG -> STS -> LDS (three trips through the L1 data pipe, L2 latency on the 35 % of first touches
 *     a phase has).  Here the HOT keys' matrices -- the ~1 500 that cover 96 % of the windows -- are copied into shared
 *     memory once per launch and every lane reads its matrix straight from there: 8 LDS.128 with XOR-swizzled 16-byte chunks,
 *     29-cycle latency, no tags, no misses.  Cold keys (4 % of the windows) are read from the global table behind an L1
 *     prefetch issued two windows ahead.
 *   * shared memory.  The per-thread statistics columns (127 KB) are gone: the statistics phase works on contiguous
 *     records with four lanes per tile and reduces by fixed shuffle trees into one row per warp (hfg_estep_quad.cuh's
 *     phase S); the scan stash lives in global scratch (coalesced); labels are staged in shared memory and leave as
 *     16-byte vectors.  That is what makes room for the hot table.
 *   * dependent chains.  The sweeps carry f and b unnormalised with exact power-of-two rescaling every fourth window:
 *     reciprocal and sum leave the loop-carried chain; posteriors, labels and pair statistics are ratios, and
 *     sum_i log c_i telescopes to log(sum of the last f) - (sum of rescaling exponents) ln 2.
 *   * the tail.  The grid reduction runs on the last CTA to arrive (atomic ticket), no fourth grid barrier.
 * Arithmetic inside a segment keeps the reference's operation order (sum over preState ascending, hmm.c:386-408; sum over
 * state ascending, hmm.c:493-520); results differ from hfg_estep.cuh only by rounding.
 */
#pragma once

#include "hfg_estep_quad.cuh"

#define HFG3_THREADS 512 /* default CTA: 16 warps, <= 128 registers per thread; one persistent CTA per SM */
#define HFG3_HOT (1u << 28) /* key word: bits 0..27 hold the slot of the key's matrix in the shared-memory table, not the key id */
#define HFG3_ID(w) ((w) & 0x0fffffffu)

namespace hfg3 {

using namespace hfgq;

/* One 4x4 transfer matrix, from the shared-memory table (hot) or the global table (cold), read as eight 16-byte chunks in
 * a LANE-ROTATED order: the c-th load of lane l fetches chunk c ^ (l & 7).  The eight lanes of a quarter-warp therefore hit
 * eight different 16-byte bank groups whatever rows they read (rows are 128-byte aligned): no bank conflicts, where a
 * row-keyed swizzle costs 2.6x (tools/mio_bench.cu: 27 against 69 cycles per warp and matrix).  The price: the registers hold
 * the matrix XOR-permuted by a lane constant,
 *     M'[a][b] = M[a ^ kp][b ^ kb],   kp = (l & 7) >> 1,  kb = 2 * (l & 1)
 * (chunk index = 2 * row + column pair).  The sweeps keep their vectors in the matching permuted order, so that the
 * products need no fix-up; what remains is one XOR-permutation of a 4-vector per window (xorperm4: 8 selects). */
__device__ __forceinline__ void load_M(uint32_t w, const double *hot, const double *tabM, int li, double (&M)[16]) {
    /* ONE code path for hot and cold keys: a generic pointer into shared or global memory, generic 128-bit loads.  (As
     * if / else the compiler duplicates the rest of the loop body into both branches, and a warp with hot and cold lanes --
     * three warp steps in four -- then runs the arithmetic twice.) */
    const double2 *p = reinterpret_cast<const double2 *>(((w & HFG3_HOT) ? hot : tabM) + (size_t) HFG3_ID(w) * 16);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double2 v;
        asm("ld.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + (c ^ li)));
        M[2 * c] = v.x;
        M[2 * c + 1] = v.y;
    }
}
/* c ? a : b as a select INSTRUCTION.  The conditions are lane constants: written as ?:, the compiler unswitches the sweep
 * loops on them (four copies of every loop, each warp running all four at quarter occupancy). */
__device__ __forceinline__ double selp(bool c, double a, double b) {
    double r;
    asm("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n selp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int) c));
    return r;
}
/* y[a] = x[a ^ d] for a lane constant d in 0..3 (d0 = d & 1, d1 = d & 2 as predicates) */
__device__ __forceinline__ void xorperm4(double (&x)[4], bool d0, bool d1) {
    const double t0 = selp(d0, x[1], x[0]), t1 = selp(d0, x[0], x[1]), t2 = selp(d0, x[3], x[2]), t3 = selp(d0, x[2], x[3]);
    x[0] = selp(d1, t2, t0);
    x[1] = selp(d1, t3, t1);
    x[2] = selp(d1, t0, t2);
    x[3] = selp(d1, t1, t3);
}
/* cold keys: pull the line into L1 ahead of its use (no register, no scoreboard).  Callers pass a key word that was loaded an
 * iteration earlier, so that the prefetch does not wait for the load */
__device__ __forceinline__ void prefetch_M(uint32_t w, const double *tabM) {
    if (w != 0u && !(w & HFG3_HOT)) asm volatile("prefetch.global.L1 [%0];" ::"l"(tabM + (size_t) HFG3_ID(w) * 16));
}

/* key words and record positions are read once per phase: from L2, leaving the small L1 to the prefetched cold matrices */
__device__ __forceinline__ uint32_t ldkw(const uint32_t *p) {
    uint32_t v;
    v = __ldg(p); // asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_row_cg(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

}  // namespace hfg3

template <int THREADS, bool NB = false>
__global__ void __launch_bounds__(THREADS, 1) hfg_estep_v3_kernel(const __grid_constant__ EstepArgs A) {
    constexpr int WARPS = THREADS / 32, QUADS = THREADS / 4;
    using namespace hfg3;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];

    /* device-resident EM: the previous launch raised the stop flag: nothing to do (read by every thread before any barrier) */
    if (A.em_mode == 1 && __ldcg(&A.em_state[0]) != 0) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = tid & 3, quad = lane >> 2;
    const int j = blockIdx.x * THREADS + tid; /* segment owned by this thread */
    const int cap = A.capacity, G = A.G, R = A.n_regions;
    const int rt_stride = QRT_STRIDE(G);
    const int NSTAT = hfg_nstat(G);

    /* shared memory carve-up (doubles) */
    double *rtab = smem;                                              /* [R][rt_stride] */
    double *warp_tot = rtab + (((size_t) R * rt_stride + 1) & ~(size_t) 1); /* [WARPS][16] warp products */
    double *warp_pre = warp_tot + WARPS * 16;                         /* [WARPS][16] exclusive prefix over warps */
    double *warp_suf = warp_pre + WARPS * 16;                         /* [WARPS][16] exclusive suffix over warps */
    double *blk_vec = warp_suf + WARPS * 16;                          /* [8] forward / backward message entering the CTA */
    double *wstat = blk_vec + 8;                                      /* [WARPS][NSTAT] statistics per warp; tail: [R][NSTAT] totals */
    int8_t *lab_s = reinterpret_cast<int8_t *>(wstat + (((size_t) (WARPS > R ? WARPS : R) * NSTAT + 1) & ~(size_t) 1));
    double *hot = reinterpret_cast<double *>(lab_s + A.lab_bytes);    /* [n_hot][16] matrices of the hot keys (lab_bytes % 16 == 0) */
    __shared__ int s_reset, s_last;
    __shared__ int s_stride[HFG_MAX_REGIONS]; /* statistics phase: the stride that spreads a region's tiles over the CTAs */
    __shared__ __align__(8) unsigned long long s_mbar;
    /* lane constants of the rotated matrix reads (load_M) */
    const int li = lane & 7;
    const bool kp0 = (li & 2) != 0, kp1 = (li & 4) != 0, kb1 = (li & 1) != 0; /* kp = li >> 1 (rows), kb = 2 * (li & 1) (columns) */
    const bool d0 = kp0, d1 = kp1 != kb1;                                     /* d = kp ^ kb */
    const int kp = li >> 1, kb = (li & 1) << 1;

    /* ---- prologue: derived per-region tables (redundantly per CTA; O(R*K) work) ---- */
    if (tid == 0) s_reset = 0;
    if (tid < R) {
        /* ~0.618 n, nudged until coprime to n */
        const int n = A.region_tile_begin[tid + 1] - A.region_tile_begin[tid];
        int st = (int) (0.6180339887 * n);
        if (st < 1) st = 1;
        for (;; st++) {
            int a = st, b = n;
            while (b) {
                const int c = a % b;
                a = b;
                b = c;
            }
            if (a == 1 || n <= 1) break;
        }
        s_stride[tid] = st;
    }
    for (int idx = tid; idx < R * 32; idx += THREADS) {
        /* one (region, mask, pre) row of the conditional transition table (Transition_getProbConditional,
         * hmm_utils.c:2278-2292) */
        const int r = idx >> 5, mask = (idx >> 2) & 7, pre = idx & 3;
        const hfg_region_params &p = A.params_inline ? A.inl_params : A.params[r];
        bool valid[5] = {true, (mask & 1) == 0, true, (mask & 2) == 0, (mask & 4) != 0};
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 5; k++)
            if (valid[k]) tot += p.trans[pre][k];
#pragma unroll
        for (int s = 0; s < 4; s++)
            rtab[(size_t) r * rt_stride + RT_TC + mask * 16 + pre * 4 + s] = valid[s] ? p.trans[pre][s] / tot : 0.0;
    }
    for (int idx = (tid + THREADS - 256) % THREADS; idx < R * (12 + G); idx += THREADS) {
        const int r = idx / (12 + G), qq = idx % (12 + G);
        const hfg_region_params &p = A.params_inline ? A.inl_params : A.params[r];
        double *rt = rtab + (size_t) r * rt_stride;
        if (qq < 4) {
            rt[RT_START + qq] = p.trans[HFG_NS][qq];
#pragma unroll
            for (int i = 0; i < 4; i++) rt[RT_UNI + qq * 4 + i] = 1.0 / (HFG_NS + 1);
        } else if (qq < 8) rt[RT_TERM + qq - 4] = p.trans[qq - 4][HFG_NS];
        else if (qq == 8) rt[RT_TEXP] = p.lambda;
        else if (qq == 9) rt[RT_TEXP + 1] = p.trunc_point;
        else if (qq == 10) rt[RT_TEXP + 2] = p.lambda / A.beta0;
        else if (qq == 11) {
            const double lam = p.lambda / A.beta0;
            const double b = A.beta0 * p.trunc_point;
            rt[RT_TEXP + 3] = 1 - exp_nonpos(-lam * b);
        } else {
            const int g = qq - 12;
            int s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (A.is_gauss[k] && g >= A.gbase[k] && g < A.gbase[k] + A.ncomp[k]) s = k;
            const int c = g - A.gbase[s];
            double *ga = rt + RT_GAUSS;
            if constexpr (NB) {
                /* negative binomial: theta, log(1 - theta), weight, r, lgamma(r), r log(theta) (hfg_nb_dev.cuh) */
                const hfgnb::Comp cc = hfgnb::comp_setup(p, s, c);
                ga[g] = p.mean[s][c];
                ga[G + g] = cc.log_1m_theta;
                ga[2 * G + g] = cc.w;
                ga[3 * G + g] = cc.r;
                ga[4 * G + g] = cc.lgamma_r;
                ga[5 * G + g] = cc.r_log_theta;
                continue;
            }
            const double vb = p.var[s][c] * A.beta0;
            ga[g] = p.mean[s][c];
            ga[G + g] = p.var[s][c];
            ga[2 * G + g] = p.weight[s][c];
            ga[3 * G + g] = vb;
            ga[4 * G + g] = 1.0 / vb;
            ga[5 * G + g] = p.weight[s][c] / sqrt(vb * 2 * HFG_PI);
        }
    }
    __syncthreads();

    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 0] = clock64();
    int nan_flag = 0, uf_flag = 0;

    /* =========================== phase T: transfer matrix of every key ======================================= */
    /* four lanes per key, keys dealt round-robin over the CTAs; lane pre evaluates the emission of every state under its own
     * alpha[pre][s] (what the reference does per window and (pre, s) pair, hmm.c:386-408) and writes row pre of the table */
    for (int p = (tid >> 2) * gridDim.x + blockIdx.x; p < A.n_keys; p += gridDim.x * QUADS) {
        Win w = decode_word(__ldg(&A.kdesc[p]), A.beta0);
        if (w.edge) {
            w.beta = A.kbeta[3 * (size_t) p];
            w.rb = A.kbeta[3 * (size_t) p + 1];
            w.sq = A.kbeta[3 * (size_t) p + 2];
        }
        const double *rt = rtab + (size_t) w.region * rt_stride;
        double row[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
            double e;
            if constexpr (NB) {
                /* the emission depends on (region, state, x) alone (NegativeBinomial_getProb, hmm_utils.c:479-516) */
                if (A.em_mode) { /* device-resident loop: the pmf from the component constants of the prologue (hfg_nb_dev.cuh) */
                    const double *ga = rt + RT_GAUSS;
                    const double lgx = __ldg(A.nb_lgx1 + (int) w.x);
                    e = 0.0;
                    for (int c = 0; c < A.ncomp[s]; c++) {
                        const int g = A.gbase[s] + c;
                        hfgnb::Comp cc;
                        cc.r = ga[3 * G + g];
                        cc.w = ga[2 * G + g];
                        cc.lgamma_r = ga[4 * G + g];
                        cc.r_log_theta = ga[5 * G + g];
                        cc.log_1m_theta = ga[G + g];
                        cc.bt = 0.0;
                        e += hfgnb::comp_prob(cc, (int) w.x, lgx, &nan_flag);
                    }
                } else
                    e = A.nb_table[((size_t) w.region * 4 + s) * HFG_NB_XSTRIDE + (int) w.x];
            } else if (!A.is_gauss[s]) {
                e = trunc_exp_prob(rt, w);
            } else {
                /* chunk starts: alpha = 0 and preX = 0 for every state (EM_fillFirstColumnForward, hmm.c:333-364; the packed
                 * word carries px = 0 there) */
                const double a = w.start ? 0.0 : A.alpha[q][s];
                e = gauss_state_prob(rt, G, A.gbase[s], A.ncomp[s], a, w, &nan_flag);
            }
            /* a chunk start is the rank-1 matrix whose rows are the unnormalised first column e * startProb */
            row[s] = w.start ? e * rt[RT_START + s] : trans_prob(rt, w, q, s) * e;
        }
        double2 *dst = reinterpret_cast<double2 *>(A.tabM + (size_t) p * 16 + q * 4);
        dst[0] = make_double2(row[0], row[1]);
        dst[1] = make_double2(row[2], row[3]);
    }
    asm volatile("fence.proxy.async;" ::: "memory"); /* the table is read through the async proxy (bulk copies) below */
    grid.sync(); /* the key table is complete */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 1] = clock64();

    /* the hot keys' matrices -> shared memory: bulk asynchronous copies (one per region: its hot keys are consecutive in the
     * table), issued by one thread, completion on an mbarrier every thread waits on */
    if (A.n_hot > 0) {
        if (tid == 0) {
            const uint32_t mb = (uint32_t) __cvta_generic_to_shared(&s_mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            uint32_t total = 0;
            for (int r = 0; r < R; r++) total += (uint32_t) __ldg(&A.hot_range[3 * r + 1]) * 128u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(total) : "memory");
            for (int r = 0; r < R; r++) {
                const int kbeg = __ldg(&A.hot_range[3 * r]), cnt = __ldg(&A.hot_range[3 * r + 1]), base = __ldg(&A.hot_range[3 * r + 2]);
                /* pieces of <= 32 KB */
                for (int o = 0; o < cnt; o += 256) {
                    const uint32_t bytes = (uint32_t) min(256, cnt - o) * 128u;
                    const uint32_t dst = (uint32_t) __cvta_generic_to_shared(hot + (size_t) (base + o) * 16);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "l"(A.tabM + (size_t) (kbeg + o) * 16), "r"(bytes), "r"(mb)
                                 : "memory");
                }
            }
        }
        __syncthreads(); /* the barrier object is initialised before anyone polls it */
        {
            const uint32_t mb = (uint32_t) __cvta_generic_to_shared(&s_mbar);
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done)
                             : "r"(mb), "r"(0u)
                             : "memory");
            }
        }
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 7] = clock64();

    const int len = A.seg_len[j];
    const int seg_first = A.seg_start[j];
    const uint32_t *wk = A.wkeyT + j; /* wk[k * cap]: key word of the k-th window of this segment */

    /* =========================== phase A: segment transfer product ============================================ */
    {
        /* P is carried with its COLUMNS permuted by kp (P'[i][a] = P[i][a ^ kp]), which is what multiplying by the permuted
         * M' from the right needs; the product comes out with columns permuted by kb and is brought back by one XOR-permutation
         * of every row by d = kp ^ kb (ALU selects, which this phase has to spare) */
        double P[16];
        mat_identity(P);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double row[4] = {P[4 * r], P[4 * r + 1], P[4 * r + 2], P[4 * r + 3]};
            xorperm4(row, kp0, kp1);
#pragma unroll
            for (int c = 0; c < 4; c++) P[4 * r + c] = row[c];
        }
        bool has_start = false;
        {
            uint32_t w0 = len > 0 ? ldkw(wk) : 0u;
            uint32_t w1 = len > 1 ? ldkw(wk + cap) : 0u;
            uint32_t w2 = len > 2 ? ldkw(wk + 2 * (size_t) cap) : 0u;
            prefetch_M(w0, A.tabM);
            prefetch_M(w1, A.tabM);
            double M[16], Mn[16];
            load_M(w0, hot, A.tabM, li, M);
#pragma unroll 1
            for (int k = 0; k < len; k++) {
                const uint32_t w3 = k + 3 < len ? ldkw(wk + (size_t) (k + 3) * cap) : 0u;
                prefetch_M(w2, A.tabM);
                load_M(w1, hot, A.tabM, li, Mn); /* the next window's matrix is in flight while this one is used */
                if (w0 & HFG_KEY_CHUNK_START) has_start = true;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    double t[4];
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        t[c] = fma(P[r * 4 + 3], M[12 + c], fma(P[r * 4 + 2], M[8 + c], fma(P[r * 4 + 1], M[4 + c], P[r * 4] * M[c])));
                    xorperm4(t, d0, d1);
#pragma unroll
                    for (int c = 0; c < 4; c++) P[r * 4 + c] = t[c];
                }
                if ((k & 3) == 3) mat_rescale(P); /* a window shrinks the product by < 1e-50: every 4th step is ample */
#pragma unroll
                for (int i = 0; i < 16; i++) M[i] = Mn[i];
                w0 = w1;
                w1 = w2;
                w2 = w3;
            }
            if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 8] = clock64();
        }
#pragma unroll
        for (int r = 0; r < 4; r++) { /* back to the canonical column order for the scans */
            double row[4] = {P[4 * r], P[4 * r + 1], P[4 * r + 2], P[4 * r + 3]};
            xorperm4(row, kp0, kp1);
#pragma unroll
            for (int c = 0; c < 4; c++) P[4 * r + c] = row[c];
        }
        mat_rescale(P);
        if (has_start) s_reset = 1; /* benign race: every writer stores 1 */

        /* ======================= phase B: scans ============================================================== */
        /* level 1: the 32 segments of a warp by shuffles; the exclusive prefix / suffix of every thread is parked in global
         * scratch (coalesced: [i][thread]) until the messages entering the CTA are known */
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
        double *stash = A.scan_stash + j;
#pragma unroll
        for (int i = 0; i < 16; i++) st_cg(stash + (size_t) i * cap, Q[i]); /* exclusive prefix inside the warp */
        warp_scan_suffix<32>(P, lane);
        mat_shfl_down(P, Q, 1);
        if (lane == 31) mat_identity(Q);
#pragma unroll
        for (int i = 0; i < 16; i++) st_cg(stash + (size_t) (16 + i) * cap, Q[i]); /* exclusive suffix inside the warp */
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 9] = clock64();
    }
    __syncthreads();
    /* level 2 over the WARPS warp products of this CTA: warp 0 builds the prefixes (and the CTA total), warp 1 the suffixes */
    if (warp == 0) {
        double Sp[16], Q[16];
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) Sp[i] = warp_tot[lane * 16 + i];
        } else {
            mat_identity(Sp);
        }
        warp_scan_prefix<32>(Sp, lane);
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
        warp_scan_suffix<32>(Ss, lane);
        mat_shfl_down(Ss, Q, 1);
        if (lane >= WARPS - 1) mat_identity(Q);
        if (lane < WARPS) {
#pragma unroll
            for (int i = 0; i < 16; i++) warp_suf[lane * 16 + i] = Q[i];
        }
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 2] = clock64();
    grid.sync(); /* orders the block totals written above */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 3] = clock64();

    /* level 3, messages entering this CTA: the products of the CTAs back to (and including) the nearest one that contains a
     * chunk start -- its product is rank-1, so nothing beyond it matters -- and, for the backward message, forward to the
     * nearest such CTA.  Warp 0 walks backward in the genome for the forward message, warp 1 forward for the backward one;
     * every lane carries the whole 4-vector, the block totals are fetched HFGQ_WALK_STAGE at a time (one per lane, all
     * loads in flight) and handed round by shuffles. */
    if (warp < 2) {
        const int b = blockIdx.x, nb = gridDim.x;
        if (warp == 0) {
            int b0 = 0; /* first CTA whose product is applied */
            for (int base = b - 1; base >= 0; base -= 32) {
                const int qq = base - lane;
                const unsigned m = __ballot_sync(FULL, qq >= 0 && __ldcg(&A.block_reset[qq]) != 0);
                if (m) {
                    b0 = base - (__ffs(m) - 1);
                    break;
                }
            }
            double v[4] = {0.25, 0.25, 0.25, 0.25};
            for (int g0 = b0; g0 < b; g0 += HFGQ_WALK_STAGE) {
                const int cnt = min(HFGQ_WALK_STAGE, b - g0);
                double T[16];
                const int src = g0 + min(lane, cnt - 1);
#pragma unroll
                for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) src * 16 + i]);
                for (int s = 0; s < cnt; s++) {
                    double o[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) o[c] = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) o[c] = fma(v[r], __shfl_sync(FULL, T[r * 4 + c], s), o[c]);
#pragma unroll
                    for (int c = 0; c < 4; c++) v[c] = o[c];
                    vec_rescale(v);
                }
            }
            if (lane < 4) blk_vec[lane] = sel4(v, lane);
        } else {
            int b1 = nb - 1; /* last CTA whose product is applied */
            for (int base = b + 1; base < nb; base += 32) {
                const int qq = base + lane;
                const unsigned m = __ballot_sync(FULL, qq < nb && __ldcg(&A.block_reset[qq]) != 0);
                if (m) {
                    b1 = base + (__ffs(m) - 1);
                    break;
                }
            }
            double u[4] = {1.0, 1.0, 1.0, 1.0};
            /* the farthest CTA first: u <- T_g * u for g = b1 .. b+1 */
            for (int g1 = b1; g1 > b; g1 -= HFGQ_WALK_STAGE) {
                const int cnt = min(HFGQ_WALK_STAGE, g1 - b);
                double T[16];
                const int src = g1 - min(lane, cnt - 1);
#pragma unroll
                for (int i = 0; i < 16; i++) T[i] = __ldcg(&A.block_tot[(size_t) src * 16 + i]);
                for (int s = 0; s < cnt; s++) {
                    double o[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        o[r] = 0.0;
#pragma unroll
                        for (int c = 0; c < 4; c++) o[r] = fma(__shfl_sync(FULL, T[r * 4 + c], s), u[c], o[r]);
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) u[r] = o[r];
                    vec_rescale(u);
                }
            }
            if (lane < 4) blk_vec[4 + lane] = sel4(u, lane);
        }
    }
    __syncthreads();

    /* messages entering this segment */
    double v_in[4], u_in[4];
    {
        double T[16];
        const double *stash = A.scan_stash + j;
#pragma unroll
        for (int i = 0; i < 4; i++) v_in[i] = blk_vec[i];
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = warp_pre[warp * 16 + i];
        vec_mat(v_in, T);
        vec_rescale(v_in);
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = ld_cg(stash + (size_t) i * cap);
        vec_mat(v_in, T);
        vec_normalize(v_in); /* f^ of the window before the segment: sums to one, as the reference's scaled forward does */
#pragma unroll
        for (int i = 0; i < 4; i++) u_in[i] = blk_vec[4 + i];
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = warp_suf[warp * 16 + i];
        mat_vec(T, u_in);
        vec_rescale(u_in);
#pragma unroll
        for (int i = 0; i < 16; i++) T[i] = ld_cg(stash + (size_t) (16 + i) * cap);
        mat_vec(T, u_in);
        vec_normalize(u_in);
    }

    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 10] = clock64();
    /* =========================== phase C1: forward inside the segment ======================================== */
    /* The vector is carried UNNORMALISED, f_k = (f_{k-1} M_k) 2^{e_k} with e_k != 0 every fourth window: the scales of
     * hmm.c:410-419 are c_k = sum(f_{k-1} M_k) / sum(f_{k-1}), so that sum_k log c_k = log(sum f_last) - ln 2 sum_k e_k
     * (the entering vector sums to one; a chunk start restarts the product with c = sum of its first column, which is the
     * same formula). */
    /* Permuted order (load_M): the vector is kept as f'[a] = f[a ^ kp]; f' * M' gives the new vector permuted by kb, one
     * xorperm4 by d = kp ^ kb restores the order f' needs.  The sum over preState therefore runs in a lane-dependent (fixed)
     * order. */
    double f[4] = {v_in[0], v_in[1], v_in[2], v_in[3]};
    xorperm4(f, kp0, kp1);
    {
        int esum = 0;
        double fsum = 1.0; /* sum of the current vector */
        uint32_t w0 = len > 0 ? ldkw(wk) : 0u;
        uint32_t w1 = len > 1 ? ldkw(wk + cap) : 0u;
        uint32_t w2 = len > 2 ? ldkw(wk + 2 * (size_t) cap) : 0u;
        prefetch_M(w0, A.tabM);
        prefetch_M(w1, A.tabM);
        double *ft = A.scrFT + j; /* segment-transposed [k][s][j], canonical state order: slot a is component a ^ kp */
        const size_t fo0 = (size_t) (0 ^ kp) * cap, fo1 = (size_t) (1 ^ kp) * cap, fo2 = (size_t) (2 ^ kp) * cap, fo3 = (size_t) (3 ^ kp) * cap;
        double M[16], Mn[16];
        load_M(w0, hot, A.tabM, li, M);
#pragma unroll 1
        for (int k = 0; k < len; k++) {
            const uint32_t w3 = k + 3 < len ? ldkw(wk + (size_t) (k + 3) * cap) : 0u;
            prefetch_M(w2, A.tabM);
            load_M(w1, hot, A.tabM, li, Mn);
            const bool start = (w0 & HFG_KEY_CHUNK_START) != 0;
            double fn[4];
            if (start) {
                /* EM_fillFirstColumnForward: f[0][s] = e * start probability = any row of the rank-1 matrix */
#pragma unroll
                for (int s = 0; s < 4; s++) fn[s] = M[s];
            } else {
                /* f[i][s] = sum_pre f[i-1][pre] * (tProb * eProb) (hmm.c:386-408) */
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    double a = 0.0;
#pragma unroll
                    for (int pre = 0; pre < 4; pre++) a += f[pre] * M[pre * 4 + s];
                    fn[s] = a;
                }
            }
            xorperm4(fn, d0, d1); /* from column order kb to the order of f' */
            const double c = ((fn[0] + fn[1]) + fn[2]) + fn[3];
            if (!start && c < 1e-50 * fsum) uf_flag = 1; /* "scale is very low": c_k < 1e-50 (hmm.c:412-415) */
            double sc = 1.0;
            if ((k & 3) == 3 || k == len - 1) {
                const int e = max_exp4(fn);
                sc = pow2_of(2046 - e); /* largest entry into [1,2) */
                esum += 1023 - e;
            }
#pragma unroll
            for (int s = 0; s < 4; s++) f[s] = fn[s] * sc;
            {
                double *fk = ft + (size_t) k * 4 * cap;
                st_cg(fk + fo0, f[0]);
                st_cg(fk + fo1, f[1]);
                st_cg(fk + fo2, f[2]);
                st_cg(fk + fo3, f[3]);
            }
            fsum = c * sc;
#pragma unroll
            for (int i = 0; i < 16; i++) M[i] = Mn[i];
            w0 = w1;
            w1 = w2;
            w2 = w3;
        }
        /* sum_k log c_k of this segment */
        A.seg_loglik[j] = len > 0 ? log(fsum) - (double) esum * 0.6931471805599453 : 0.0;
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 4] = clock64(); /* thread 0's own C1 end (no barrier here) */

    /* =========================== phase C2: backward + decode ================================================= */
    const bool lab_smem = A.lab_bytes > 0;
    const int cta_w0 = A.seg_start[min(blockIdx.x * THREADS, max(A.n_seg - 1, 0))]; /* first window of this CTA */
    const int lab_shift = cta_w0 & 15; /* shared-memory offset == global offset (mod 16): the copy-out moves aligned vectors */
    if (!A.forward_only) {
        /* f[] holds the forward vector of the segment's last window.  b is a direction, carried unnormalised with exact
         * rescaling: the decode needs it up to a positive factor, the statistics records are normalised one by one
         * (sum_{pre,s} f_{i-1}[pre] M_i[pre][s] b_i[s] = 1, the invariant of the reference's scaling, hmm.c:452-467,613-614). */
        /* Permuted order: b is kept as b'[b] = b[b ^ kb] (what M' * b' needs), the forward vectors are read in the same order
         * (fp'[b] = fp[b ^ kb]); M' b' comes out permuted by kp and is brought to order kb by one xorperm4. */
        double bh[4] = {1.0, 1.0, 1.0, 1.0};
        /* decode of one window: EM_getPosterior / EM_getMostProbableState (hmm.c:671-692), first maximum
         * (common.c:292-303); a common positive factor does not change the order */
        auto decode = [&](const double (&fw)[4], const double (&bw)[4], int gi) {
            /* (both in order kb: slots (0,1) and (2,3) trade places when kb != 0) */
            const double h0 = fw[0] * bw[0], h1 = fw[1] * bw[1], h2 = fw[2] * bw[2], h3 = fw[3] * bw[3];
            double g0 = selp(kb1, h2, h0), g1 = selp(kb1, h3, h1), g2 = selp(kb1, h0, h2), g3 = selp(kb1, h1, h3);
            if (A.posteriors) {
                const double tot = ((g0 + g1) + g2) + g3;
                g0 /= tot;
                g1 /= tot;
                g2 /= tot;
                g3 /= tot;
                st_row_cg(A.posteriors + (size_t) gi * 4, g0, g1, g2, g3);
            }
            int best = 0;
            double m = g0;
            if (m < g1) { m = g1; best = 1; }
            if (m < g2) { m = g2; best = 2; }
            if (m < g3) { m = g3; best = 3; }
            if (lab_smem) lab_s[gi - cta_w0 + lab_shift] = (int8_t) best;
            else A.labels[gi] = (int8_t) best;
        };
        bool done = len == 0;
        if (len > 0) {
            const uint32_t wl = ldkw(wk + (size_t) (len - 1) * cap);
            if (wl & HFG_KEY_CHUNK_END) {
                /* EM_fillLastColumnBackward (hmm.c:452-467): b = terminationProb / scale */
                const int key = (wl & HFG3_HOT) ? __ldg(&A.hot_key[HFG3_ID(wl)]) : (int) HFG3_ID(wl);
                const double *rt = rtab + (size_t) HFG_OBS_REGION(__ldg(&A.kdesc[key])) * rt_stride;
#pragma unroll
                for (int s = 0; s < 4; s++) bh[s] = rt[RT_TERM + s] * HFG_INV_TERM;
            } else {
#pragma unroll
                for (int s = 0; s < 4; s++) bh[s] = u_in[s];
            }
            xorperm4(bh, false, kb1);
            xorperm4(f, d0, d1); /* the last forward vector: from order kp to order kb */
            decode(f, bh, seg_first + len - 1); /* the segment's last window; every other window is decoded at the end of
                                                   the step that produces its b */
        }
        const uint32_t *wp = A.wposT + j; /* position of the window's record in key-list order, or HFGQ_NOPOS */
        uint32_t w0 = len > 0 ? ldkw(wk + (size_t) (len - 1) * cap) : 0u;
        uint32_t w1 = len > 1 ? ldkw(wk + (size_t) (len - 2) * cap) : 0u;
        uint32_t w2 = len > 2 ? ldkw(wk + (size_t) (len - 3) * cap) : 0u;
        prefetch_M(w0, A.tabM);
        prefetch_M(w1, A.tabM);
        const double *ft = A.scrFT + j;
        const size_t bo0 = (size_t) (0 ^ kb) * cap, bo1 = (size_t) (1 ^ kb) * cap, bo2 = (size_t) (2 ^ kb) * cap, bo3 = (size_t) (3 ^ kb) * cap;
        double vp[4] = {v_in[0], v_in[1], v_in[2], v_in[3]};
        xorperm4(vp, false, kb1);
        /* forward vector of window k - 1 (the entering message for k == 0), in order kb */
        auto load_fp = [&](int k, double (&fp)[4]) {
            if (k > 0) {
                const double *fk = ft + (size_t) (k - 1) * 4 * cap;
                fp[0] = ld_cg(fk + bo0);
                fp[1] = ld_cg(fk + bo1);
                fp[2] = ld_cg(fk + bo2);
                fp[3] = ld_cg(fk + bo3);
            } else {
#pragma unroll
                for (int s = 0; s < 4; s++) fp[s] = vp[s];
            }
        };
        double M[16], Mn[16], fp[4], fpn[4];
        uint32_t pos = HFGQ_NOPOS, posn = HFGQ_NOPOS;
        if (len > 0) {
            load_M(w0, hot, A.tabM, li, M);
            load_fp(len - 1, fp);
            pos = ldkw(wp + (size_t) (len - 1) * cap);
        }
#pragma unroll 1
        for (int k = len - 1; k >= 0 && !done; k--) {
            const uint32_t w3 = k >= 3 ? ldkw(wk + (size_t) (k - 3) * cap) : 0u;
            prefetch_M(w2, A.tabM);
            /* everything step k - 1 reads from memory is requested now */
            load_M(w1, hot, A.tabM, li, Mn);
            if (k > 0) {
                load_fp(k - 1, fpn);
                posn = ldkw(wp + (size_t) (k - 1) * cap);
            }
            if (w0 & HFG_KEY_CHUNK_START) {
                done = true; /* first window of a chunk: nothing to the left */
            } else {
                const int gi = seg_first + k;
                /* b[i-1][pre] = sum_s tProb*eProb*b[i][s] (hmm.c:493-520) */
                double bn[4];
#pragma unroll
                for (int pre = 0; pre < 4; pre++)
                    bn[pre] = ((M[pre * 4] * bh[0] + M[pre * 4 + 1] * bh[1]) + M[pre * 4 + 2] * bh[2]) + M[pre * 4 + 3] * bh[3];
                xorperm4(bn, d0, d1); /* from row order kp to order kb */
                if (pos != HFGQ_NOPOS) {
                    /* the window's record for the statistics: (f_{i-1}, b_i / (f_{i-1} . M_i b_i)), interleaved f0 b0 f1 b1 | f2 b2 f3 b3
                     * in canonical state order: slots (0,1) are states (kb, kb + 1) */
                    const double dot = ((fp[0] * bn[0] + fp[1] * bn[1]) + fp[2] * bn[2]) + fp[3] * bn[3];
                    const double r = 1.0 / dot;
                    double *rec = A.scrXB + (size_t) pos * 8;
                    st_row_cg(rec + 2 * kb, fp[0], bh[0] * r, fp[1], bh[1] * r);
                    st_row_cg(rec + 2 * (kb ^ 2), fp[2], bh[2] * r, fp[3], bh[3] * r);
                }
                double sc = 1.0;
                if ((k & 3) == 0) sc = pow2_of(2046 - max_exp4(bn));
#pragma unroll
                for (int s = 0; s < 4; s++) bh[s] = bn[s] * sc;
                if (k > 0) decode(fp, bh, gi - 1);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) M[i] = Mn[i];
#pragma unroll
            for (int s = 0; s < 4; s++) fp[s] = fpn[s];
            pos = posn;
            w0 = w1;
            w1 = w2;
            w2 = w3;
        }
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 11] = clock64();
    }
    __syncthreads(); /* this CTA's labels are final */
    if (!A.forward_only && lab_smem) {
        /* labels of this CTA's windows (one contiguous range: segments are in genome order): shared memory -> global memory
         * and, for the blocking calls, -> the caller's page-locked buffer over PCIe, as aligned 16-byte vectors (posted
         * writes: they drain while the statistics phase runs) */
        const int s0 = blockIdx.x * THREADS;
        if (s0 < A.n_seg) {
            const long long w_begin = cta_w0;
            const long long w_end = s0 + THREADS < A.n_seg ? (long long) A.seg_start[s0 + THREADS] : (long long) A.n_windows;
            const long long a16 = (w_begin + 15) & ~15LL, b16 = w_end & ~15LL, base16 = w_begin & ~15LL;
            if (a16 < b16) {
                for (long long o = a16 + 16LL * tid; o < b16; o += 16LL * THREADS) {
                    const int4 v = *reinterpret_cast<const int4 *>(lab_s + (o - base16));
                    *reinterpret_cast<int4 *>(A.labels + o) = v;
                    if (A.labels_host) *reinterpret_cast<int4 *>(A.labels_host + o) = v;
                }
            }
            const long long head_end = a16 < w_end ? a16 : w_end, tail_begin = b16 > head_end ? b16 : head_end;
            for (long long o = w_begin + tid; o < head_end; o += THREADS) {
                const int8_t v = lab_s[o - base16];
                A.labels[o] = v;
                if (A.labels_host) A.labels_host[o] = v;
            }
            for (long long o = tail_begin + tid; o < w_end; o += THREADS) {
                const int8_t v = lab_s[o - base16];
                A.labels[o] = v;
                if (A.labels_host) A.labels_host[o] = v;
            }
        }
    } else if (A.labels_host != NULL && !A.forward_only && warp >= WARPS - 2) {
        const int s0 = blockIdx.x * THREADS;
        if (s0 < A.n_seg) {
            const long long w_begin = cta_w0;
            const long long w_end = s0 + THREADS < A.n_seg ? (long long) A.seg_start[s0 + THREADS] : (long long) A.n_windows;
            const int t = (warp - (WARPS - 2)) * 32 + lane; /* 0..63 */
            for (long long o = w_begin + t; o < w_end; o += 64) A.labels_host[o] = __ldcg(A.labels + o);
        }
    }
    if (uf_flag) atomicOr(A.err_flags, 1);
    grid.sync(); /* the records of every window are in place */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 5] = clock64();

    /* =========================== phases S and D, region by region ============================================ */
    /* statistic rows: 0..15 transition counts, 16..17 truncated exponential, 18+3g.. (meanNum, den, varNum) of Gaussian
     * component g, last row the log-likelihood (kept in region 0).  Four lanes per tile.  Every CTA takes a contiguous, equal
     * share of the tile POSITIONS (regions are contiguous: a CTA meets one or two regions, not all R; with the tile length the
     * layout picks, at most one tile per quad).  Inside a region the tiles are ordered hottest key first -- full tiles first,
     * one-window tiles last -- so position u of a region maps to tile (u * stride) mod n with a stride coprime to n: every CTA
     * gets the same mix of long and short tiles. */
    const long long n_tiles_all = A.forward_only ? 0 : __ldg(&A.region_tile_begin[HFG_MAX_REGIONS]);
    const int blk_t0 = (int) (n_tiles_all * blockIdx.x / gridDim.x), blk_t1 = (int) (n_tiles_all * (blockIdx.x + 1) / gridDim.x);
    for (int r = 0; r < R; r++) {
        const int rb = A.forward_only ? 0 : __ldg(&A.region_tile_begin[r]), re = A.forward_only ? 0 : __ldg(&A.region_tile_begin[r + 1]);
        const int t_begin = max(blk_t0, rb), t_end = min(blk_t1, re);
        const long long rn = re - rb, rstride = s_stride[r];
        if (r > 0 && t_begin >= t_end) { /* (CTA-uniform) none of this region's tiles here; region 0 carries the log-likelihood */
            for (int qq = tid; qq < NSTAT; qq += THREADS) A.partials[((size_t) r * NSTAT + qq) * gridDim.x + blockIdx.x] = 0.0;
            continue;
        }
        double *ws = wstat + (size_t) warp * NSTAT; /* this warp's row: one writer per entry */
        for (int qq = lane; qq < NSTAT; qq += 32) ws[qq] = 0.0;
        __syncwarp();
        if (r == 0) {
            const double ll = warp_sum(ld_cg(A.seg_loglik + j)); /* written by this thread in C1 */
            if (lane == 0) ws[NSTAT - 1] = ll;
        }
#pragma unroll 1
        for (int u0 = t_begin; u0 < t_end; u0 += QUADS) { /* (CTA-uniform trip count) */
            const int u = u0 + warp * 8 + quad;
            const bool act = u < t_end;
            const int t = act ? rb + (int) (((long long) (u - rb) * rstride) % rn) : 0;
            const int p = act ? __ldg(&A.tile_key[t]) : 0, lb = act ? __ldg(&A.tile_begin[t]) : 0;
            const int ln = act ? __ldg(&A.tile_cnt[t]) : 0;
            /* The tile's records are consecutive in key-list order.  Lane q takes windows q, q+4, ... of the tile and sums
             * their outer products f (x) b (the four lanes of a quad read 256 contiguous bytes per step); the quad then
             * folds the four sums so that lane pre = q ends with row pre of S[pre][s] = sum_i f_{i-1}[pre] b_i[s]. */
            double S[4];
            {
                double O[16];
#pragma unroll
                for (int i = 0; i < 16; i++) O[i] = 0.0;
                /* RECS records of this lane in flight at a time (the loads are L2 round trips) */
                constexpr int RECS = 4;
                for (int i0 = q; i0 < ln; i0 += 4 * RECS) {
                    double lo[RECS][4], hi[RECS][4];
#pragma unroll
                    for (int u = 0; u < RECS; u++) {
                        const int i = min(i0 + 4 * u, ln - 1); /* (a repeat of the last record, masked below) */
                        const double *rec = A.scrXB + (size_t) (lb + i) * 8; /* f0 b0 f1 b1 | f2 b2 f3 b3 */
                        ld_row_cg(rec, lo[u]);
                        ld_row_cg(rec + 4, hi[u]);
                    }
#pragma unroll
                    for (int u = 0; u < RECS; u++) {
                        if (i0 + 4 * u < ln) {
                            const double fw[4] = {lo[u][0], lo[u][2], hi[u][0], hi[u][2]}, bw[4] = {lo[u][1], lo[u][3], hi[u][1], hi[u][3]};
#pragma unroll
                            for (int a = 0; a < 4; a++)
#pragma unroll
                                for (int b = 0; b < 4; b++) O[a * 4 + b] = fma(fw[a], bw[b], O[a * 4 + b]);
                        }
                    }
                }
                __syncwarp();
                /* fold over the quad: first the lane pairs (q, q^2) split rows {0,1} / {2,3}, then (q, q^1) split the two rows */
                const bool up = (q & 2) != 0;
                double H[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const double keep = up ? O[8 + i] : O[i], send = up ? O[i] : O[8 + i];
                    H[i] = keep + __shfl_xor_sync(FULL, send, 2);
                }
                const bool odd = (q & 1) != 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double keep = odd ? H[4 + i] : H[i], send = odd ? H[i] : H[4 + i];
                    S[i] = keep + __shfl_xor_sync(FULL, send, 1);
                }
            }
            /* The warps leave the record loop at very different times (short and long tiles, L2 queueing).  Measured: warps that
             * run the estimator code below while others are still elsewhere take 4-5x longer (the code of this phase does not fit
             * the instruction caches when warps are spread over it); in lockstep the whole CTA is through in the time the
             * stragglers needed alone.  Hence a CTA barrier between the memory-bound and the compute-bound half. */
            __syncthreads();
            if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 12] = clock64();
            Win w = decode_word(__ldg(&A.kdesc[p]), A.beta0);
            if (w.edge) {
                w.beta = A.kbeta[3 * (size_t) p];
                w.rb = A.kbeta[3 * (size_t) p + 1];
                w.sq = A.kbeta[3 * (size_t) p + 2];
            }
            const double *rt = rtab + (size_t) w.region * rt_stride;
            double Mr[4];
            ld_row(A.tabM + (size_t) p * 16 + q * 4, Mr);
            /* pooled pair counts preState q -> state s (count / terminationProb, hmm.c:613-614) */
            double xi[4];
#pragma unroll
            for (int s = 0; s < 4; s++) xi[s] = S[s] * Mr[s];
            {
                /* transition counts (hmm_utils.c:2010-2015): summed over the 8 quads; the lanes with bits (4, 3) = (a, b) end
                 * with count [q][2a + b] */
                const double tc = warp_fold<4, 3>(xi, lane);
                if ((lane & 4) == 0) ws[q * 4 + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)] += tc;
            }
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if constexpr (NB) {
                    /* hmm.c:615-617: the pair mass goes into the state's histogram over x; one tile = one key = one x */
                    double col = xi[s];
                    col += __shfl_xor_sync(FULL, col, 1);
                    col += __shfl_xor_sync(FULL, col, 2);
                    if (act && q == 0) A.nb_tile_col[(size_t) t * 4 + s] = col;
                } else if (!A.is_gauss[s]) {
                    /* TruncExponential_updateEstimator (hmm_utils.c:1027-1034) */
                    const double v[2] = {xi[s] * w.x, xi[s]};
                    const double tot = warp_fold<2, 5>(v, lane);
                    if ((lane & 15) == 0) ws[16 + (lane >> 4)] += tot;
                } else {
                    /* Gaussian_updateEstimator (hmm_utils.c:812-839) for preState q: x_adjusted and z follow alpha[q][s];
                     * responsibilities w = count * p_c / sum_c p_c with count = f*t*e*b/term and e = sum_c p_c, so the
                     * emission cancels: w = H * p_c with H = (f*t) * b / term */
                    const int n = A.ncomp[s], g0 = A.gbase[s];
                    const double a = A.alpha[q][s], oma = 1.0 - a;
                    const double x_adj = (w.x - a * w.px) / oma; /* hmm_utils.c:818 */
                    const double *ga = rt + RT_GAUSS;
                    /* lanes 0 / 8 / 16 end with (sum w x_adj, sum w, sum w z^2) */
                    if (n == 1) {
                        /* single component: the responsibility is 1 */
                        const double z = (x_adj - ga[g0]) * oma;
                        const double v[3] = {xi[s] * x_adj, xi[s], xi[s] * z * z};
                        const double tot = warp_fold<3, 5>(v, lane);
                        if ((lane & 7) == 0 && lane < 24) ws[18 + 3 * g0 + (lane >> 4) * 2 + ((lane >> 3) & 1)] += tot;
                    } else {
                        const double H = S[s] * trans_prob(rt, w, q, s);
#pragma unroll 1
                        for (int c = 0; c < n; c++) {
                            const int g = g0 + c;
                            const double mu = ga[g];
                            const double pc = gauss_comp(rt, G, g, a, w, &nan_flag);
                            const double wgt = H * pc;
                            const double z = (x_adj - mu) * oma;
                            const double v[3] = {wgt * x_adj, wgt, wgt * z * z};
                            const double tot = warp_fold<3, 5>(v, lane);
                            if ((lane & 7) == 0 && lane < 24) ws[18 + 3 * g + (lane >> 4) * 2 + ((lane >> 3) & 1)] += tot;
                        }
                    }
                }
            }
        }
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 13] = clock64();
        __syncthreads();
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 14] = clock64();
        /* deterministic CTA reduction: the warps' rows in warp order */
        for (int qq = tid; qq < NSTAT; qq += THREADS) {
            double sum = 0.0;
#pragma unroll 4
            for (int wv = 0; wv < WARPS; wv++) sum += wstat[(size_t) wv * NSTAT + qq];
            A.partials[((size_t) r * NSTAT + qq) * gridDim.x + blockIdx.x] = sum; /* [R][NSTAT][grid] */
        }
        if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 15] = clock64();
        __syncthreads();
    }
    if (nan_flag) atomicOr(A.err_flags, 2);

    if constexpr (NB) {
        /* the histogram the estimators are fed from (hmm.c:615-617) is folded by the whole grid, one (region, coverage bin) at
         * a time per CTA (hfg_nb_dev.cuh): the last CTA's tail reads it (device-resident loop) or the host does (blocking calls) */
        if (A.em_mode != 2 && !A.forward_only) { /* (the blocking calls read the folded histogram back: 8 KB per region) */
            grid.sync(); /* every tile's column sums are in place */
            hfgnb::grid_fold_histogram<THREADS>(R, A.nb_tile_col, A.nb_bin_begin, A.nb_bin_tiles, A.nb_hist, warp_tot);
        }
    }
    hfg_estep_tail<THREADS, NB>(A, wstat, warp_tot, &s_last);
}
