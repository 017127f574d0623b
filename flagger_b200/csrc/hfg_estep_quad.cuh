/*
 * hfg_estep_quad.cuh -- the E-step kernel of libhfg, second generation (sm_100a, fp64, no tensor cores).
 *
 * Same job as hfg_estep.cuh (EM_runForward + EM_runBackward + EM_updateEstimators + the label loop of
 * submodules/hmm/hmm.c:423-434, 535-545, 638-650, 715-737 for ALL chunks of one EM_runOneIterationForList call,
 * hmm.c:739-780, in one persistent cooperative launch), same data layout (observation keys, segments, statistics
 * tiles), same phases -- but a different mapping of the 4-state algebra onto the machine:
 *
 *   FOUR LANES PER SEGMENT ("quad").  Lane q of a quad owns row q (or column q) of every 4x4 object of its segment:
 *     phase T   lane pre evaluates the four emissions under ITS alpha[pre][.] and writes row pre of the key's matrix;
 *     phase A   lane q carries row q of the segment product  P <- P * M  (M read whole, the four lanes read the same
 *               128-byte line: one L1 line per quad and step, no shared-memory staging);
 *     phase B   8 quads of a warp: Kogge-Stone over shuffles; the warps of a CTA: Kogge-Stone through shared memory,
 *               rows spread over four lanes; CTAs: block totals in global memory + one grid barrier;
 *     phase C1  lane s reads COLUMN s of M -- row s of the TRANSPOSED table the key phase writes next to the table, one
 *               256-bit load -- and forms f[s] = sum_pre f[pre] M[pre][s] in the reference's order; the four sums are
 *               exchanged by shuffles;
 *     phase C2  lane pre reads ROW pre of M (one 256-bit load) and forms b[pre] = sum_s M[pre][s] b[s];
 *     phase S   a tile's records are contiguous (they are written in key-list order): every lane takes whole windows and
 *               sums their outer products f (x) b, the quad folds the four sums so that lane pre holds row pre, then runs
 *               the estimator updates of ITS preState; sums over lanes / quads by fixed shuffle trees (deterministic).
 *   The unit that bounds this kernel is the L1 data pipe: one wavefront per distinct 128-byte line an instruction touches,
 *   and one per shuffle (ncu, profiles/).  Hence: every matrix is fetched by ONE instruction per phase (8 quads = 8 lines);
 *   per-window scratch is segment-transposed (the 8 quads of a warp step touch 2 lines, not 8); labels are staged in shared
 *   memory and leave as 16-byte vectors; the sweeps carry f and b UNNORMALISED (exact power-of-two rescaling every fourth
 *   window), which takes reciprocal and sum out of the loop-carried chain -- posteriors, labels and the pair statistics are
 *   ratios, and sum_i log c_i telescopes to log(sum of the last f) - (sum of the rescaling exponents) ln 2.
 *   What this buys over one thread per segment (hfg_estep.cuh): ~8 live doubles per lane instead of 48+, so 32 warps per
 *   SM instead of 16 and no spills; every matrix is read once per phase straight from L1 (the old kernel moved each one
 *   through LDG -> STS -> LDS, and the L1 data pipe was its top unit); the per-thread statistics columns (127 KB of
 *   shared memory) become per-warp rows (a few KB), which leaves ~190 KB of L1 for the key table; the grid reduction is
 *   done by the last CTA to arrive (atomic ticket) instead of CTA 0 behind a fourth grid barrier.
 *
 * Arithmetic: the per-window recurrences keep the reference's operation order exactly as hfg_estep.cuh does; scan
 * arithmetic (no reference order exists) uses fma.  Results differ from hfg_estep.cuh only by rounding.
 */
#pragma once

#include "hfg_estep.cuh"
#include "hfg_nb_dev.cuh"

#define HFGQ_THREADS 1024 /* 256 segments per CTA, one persistent CTA per SM */
#define QRT_STRIDE(G) (RT_GAUSS + 6 * (G)) /* per-region table in shared memory: hfg_estep.cuh's layout without the task table */
#define HFGQ_WALK_STAGE 32 /* block totals staged per round of a long inter-CTA walk */
#define HFGQ_LAB_SMAX 64   /* longest segment whose labels are staged in shared memory (else: byte stores to global memory) */
#define HFGQ_NOPOS 0xffffffffu /* wposT entry of a window that is in no statistics list */

namespace hfgq {

using namespace hfgk;

#define FULL 0xffffffffu

__device__ __forceinline__ double sel4(const double (&a)[4], int q) {
    return q == 0 ? a[0] : (q == 1 ? a[1] : (q == 2 ? a[2] : a[3]));
}

/* largest high word over the 16 entries of a matrix whose rows live in the four lanes of a quad */
__device__ __forceinline__ int quad_max_hi(const double (&r)[4]) {
    int hi = max(max(__double2hiint(r[0]), __double2hiint(r[1])), max(__double2hiint(r[2]), __double2hiint(r[3])));
    hi = max(hi, __shfl_xor_sync(FULL, hi, 1));
    hi = max(hi, __shfl_xor_sync(FULL, hi, 2));
    return hi;
}
/* exact (power-of-two) rescaling of a quad-distributed matrix: largest entry into [1,2) */
__device__ __forceinline__ void quad_rescale(double (&r)[4]) {
    const int hi = quad_max_hi(r);
    const double s = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] *= s;
}
/* the same for a 4-vector held whole by one lane */
__device__ __forceinline__ void vec_rescale(double (&v)[4]) {
    const int hi = max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
    const double s = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] *= s;
}

/* out = a * B for a row vector a and a matrix B whose row r lives in lane b0 + r (as Brow) */
__device__ __forceinline__ void row_times_quadmat(const double (&a)[4], const double (&Brow)[4], int b0, double (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const double b = __shfl_sync(FULL, Brow[c], b0 + r);
            out[c] = r == 0 ? a[0] * b : fma(a[r], b, out[c]);
        }
    }
}

/* out = a * B, B row-major in (shared) memory */
__device__ __forceinline__ void row_times_mat(const double (&a)[4], const double *B, double (&out)[4]) {
#pragma unroll
    for (int c = 0; c < 4; c++)
        out[c] = fma(a[3], B[12 + c], fma(a[2], B[8 + c], fma(a[1], B[4 + c], a[0] * B[c])));
}
/* out = B * u */
__device__ __forceinline__ void mat_times_col(const double *B, const double (&u)[4], double (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++)
        out[r] = fma(B[r * 4 + 3], u[3], fma(B[r * 4 + 2], u[2], fma(B[r * 4 + 1], u[1], B[r * 4] * u[0])));
}

__device__ __forceinline__ void ld_row(const double *p, double (&v)[4]) {
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
/* L2 loads of per-launch scratch written by other threads (ordered by __syncwarp / the grid barrier before them) */
__device__ __forceinline__ void ld_row_cg(const double *p, double (&v)[4]) {
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ double ld_cg(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_cg(double *p, double v) { asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

__device__ __forceinline__ void st_cg2(double *p, double a, double b) {
    asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
/* exponent (biased) of the largest of four non-negative doubles */
__device__ __forceinline__ int max_exp4(const double (&v)[4]) {
    const int hi = max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
    return (hi >> 20) & 0x7ff;
}
__device__ __forceinline__ double pow2_of(int biased) { return __hiloint2double(biased << 20, 0); } /* 2^(biased-1023) */

/* Sums of N (<= 4) per-lane values over the 32 lanes of a warp by recursive halving: the first rounds split the values
 * between the two halves (each lane sends what the other half keeps), the rest are butterflies.  Returns the total of
 * value 2 * bit4(lane) + bit3(lane) (N > 2), bit4(lane) (N == 2): i.e. lanes 0 / 8 / 16 / 24 end with values 0 / 1 / 2 / 3
 * (N == 2: lanes 0 / 16 with 0 / 1).  A fixed tree: the same bits on every run.  6 (N > 2) or 5 shuffles of doubles
 * instead of 5 N.  bits = 5: over all lanes; bits = 3: over the 8 quads only (lanes with equal lane & 3). */
template <int N, int BITS>
__device__ __forceinline__ double warp_fold(const double (&v)[N], int lane) {
    static_assert(N >= 2 && N <= 4, "2..4 values");
    double x;
    int first = 16;
    if (N > 2) {
        const bool hi = (lane & 16) != 0;
        const double v3 = N > 3 ? v[N > 3 ? 3 : 0] : 0.0;
        double k0 = hi ? v[2] : v[0], k1 = hi ? v3 : v[1];
        const double s0 = hi ? v[0] : v[2], s1 = hi ? v[1] : v3;
        k0 += __shfl_xor_sync(FULL, s0, 16);
        k1 += __shfl_xor_sync(FULL, s1, 16);
        const bool hi2 = (lane & 8) != 0;
        x = (hi2 ? k1 : k0) + __shfl_xor_sync(FULL, hi2 ? k0 : k1, 8);
        first = 4;
    } else {
        const bool hi = (lane & 16) != 0;
        x = (hi ? v[1] : v[0]) + __shfl_xor_sync(FULL, hi ? v[0] : v[1], 16);
        first = 8;
    }
    const int last = BITS == 5 ? 1 : 4;
#pragma unroll
    for (int off = first; off >= last; off >>= 1) x += __shfl_xor_sync(FULL, x, off);
    return x;
}

/* sum over the 32 lanes / over the 8 quads (same q) of a warp: fixed xor trees, identical on every lane */
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}
__device__ __forceinline__ double quads_sum(double v) {
#pragma unroll
    for (int off = 16; off >= 4; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

/* emission of Gaussian state s (components g0 .. g0+n-1) under dependency factor a: the sum of the component pdfs in
 * component order (Gaussian_getProb, hmm_utils.c:753-758) */
__device__ __forceinline__ double gauss_state_prob(const double *rt, int G, int g0, int n, double a, const Win &w, int *nan) {
    double tot = 0.0;
    for (int c = 0; c < n; c++) tot += gauss_comp(rt, G, g0 + c, a, w, nan);
    return tot;
}

}  // namespace hfgq

/* ---- the rate fit of the kernel tail -------------------------------------------------------------------------------------
 * The tail runs once per launch, on one CTA, on cold instruction caches: an instruction executed for the first time costs a
 * fetch from L2 (measured: ~15 cycles per instruction, 11 000 cycles for the first objective evaluation against ~1 000 warm).
 * Hence (i) one copy of each piece of code (__noinline__ pieces shared by all call sites) and (ii) SCOUT warps: while the
 * other warps sum the CTAs' partials, four otherwise idle warps run the four pieces on dummy numbers, so that the fit proper
 * finds its instructions in the SM's cache (hfg_estep_tail). */
__device__ __noinline__ double hfg_objective_dev(double rate, double trunc, double sum_x, double sum_w) {
    return hfg_trunc_exp_objective(rate, trunc, sum_x, sum_w);
}
__device__ __forceinline__ double *hfg_align16(double *p) { return reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t) 15); }
#define HFG_TAIL_TOP 576 /* doubles at the top of the tail's work area: 4 x 64 scout scratch, 256 scout snapshots, flags */
struct HfgFitState {
    double lo, x1, x2, hi, span, y1, y2;
};
/* x*: Newton on g1(x) = w/x - w T q - s (the objective's derivative), q = 1/expm1(xT), steps kept inside the bracket; then
 * the Taylor coefficients of the objective around x*: co[0] = x*, co[1..4] = derivatives 1..4 over 1!, 2!, 3!, 4!.
 * Inaccuracy here only costs predictions. */
__device__ __noinline__ void hfg_fit_newton_dev(double x, double lo, double hi, double trunc, double sum_x, double sum_w, double *co) {
    bool last = false;
#pragma unroll 1
    for (int it = 0; it < 8; it++) {
        const double q = 1.0 / expm1(x * trunc);
        const double ix = 1.0 / x, q1 = q * (1.0 + q);
        const double g1 = sum_w * ix - sum_w * trunc * q - sum_x;
        const double g2 = -sum_w * ix * ix + sum_w * trunc * trunc * q1;
        if (last || it == 7) { /* the coefficients at the last iterate */
            const double T2 = trunc * trunc, ix2 = ix * ix;
            co[0] = x;
            co[1] = g1;
            co[2] = 0.5 * g2;
            co[3] = (2.0 * sum_w * ix2 * ix - sum_w * T2 * trunc * q1 * (1.0 + 2.0 * q)) * (1.0 / 6.0);
            co[4] = (-6.0 * sum_w * ix2 * ix2 + sum_w * T2 * T2 * q1 * (1.0 + 6.0 * q1)) * (1.0 / 24.0);
            break;
        }
        double xn = x - g1 * (1.0 / g2);
        if (!(xn > lo)) xn = 0.5 * (x + lo);
        if (!(xn < hi)) xn = 0.5 * (x + hi);
        last = fabs(xn - x) <= 1e-7 * x; /* quadratic convergence: the next iterate is good to ~1e-14 */
        x = xn;
    }
}
/* PREDICTED round: L <= 32 steps of the serial search by one warp.  The position of x* (and, when x* lies between the two
 * interior points, the Taylor polynomial) says how each comparison will come out, so the walk needs no objective values:
 * every lane walks the same predicted path, lane 0 leaves each step's point, bracket and bookkeeping in `snap` ([32][8]
 * doubles of shared memory); lane t then evaluates the point of step t, and one ballot checks every predicted comparison
 * against the evaluated values.  Returns m: steps 0 .. m-1 are exactly the serial routine's steps (state updated to after
 * step m-1); m < L = the prediction of step m was wrong. */
__device__ __noinline__ int hfg_fit_pred_round_dev(HfgFitState *st, const double *co, double trunc, double sum_x, double sum_w, int lane,
                                                   int L, double *snap, long long *clk) {
    const double inv_phi = (sqrt(5.0) - 1.0) / 2.0, inv_phi2 = (3.0 - sqrt(5.0)) / 2.0;
    const unsigned FULLW = 0xffffffffu;
    const double xs = co[0], c1 = co[1], c2 = co[2], c3 = co[3], c4 = co[4];
    const double y1 = st->y1, y2 = st->y2;
    double slo = st->lo, sx1 = st->x1, sx2 = st->x2, shi = st->hi, sspan = st->span;
    const unsigned snap_s = (unsigned) __cvta_generic_to_shared(snap);
    /* The walk is ONE dependent chain on one warp: its time is its instruction count (a dependent instruction issues every
     * ~4 cycles).  So the loop carries only what the next step needs -- the bracket, the two interior points and the
     * polynomial's value at them (one new evaluation per step, whether the step needs it or not), every update a select, no
     * branch -- and leaves behind the step's point (one store) and its predicted outcome (one bit).  Who held which value at
     * which step is replayed from the bits afterwards, by all lanes at once.  The arithmetic that produces the points is the
     * serial routine's (lo + inv_phi2 * span, lo + inv_phi * span; no contraction). */
    auto poly = [&](double x) {
        const double d = x - xs;
        return d * fma(d, fma(d, fma(d, c4, c3), c2), c1);
    };
    double q1 = poly(sx1), q2 = poly(sx2);
    unsigned bits = 0u;
#pragma unroll 1
    for (int t = 0; t < L; t++) {
        /* both points on one side of the maximum: the nearer one is higher; x* between them: the polynomial decides */
        const bool left = sx2 <= xs, right = sx1 >= xs;
        const bool b = left ? false : (right ? true : q1 > q2);
        bits |= (unsigned) b << t;
        sspan = inv_phi * sspan;
        const double nlo = b ? slo : sx1, nhi = b ? sx2 : shi;
        const double xn = nlo + (b ? inv_phi2 : inv_phi) * sspan;
        const double qn = poly(xn);
        const double nx1 = b ? xn : sx2, nx2 = b ? sx1 : xn;
        const double nq1 = b ? qn : q2, nq2 = b ? q1 : qn;
        slo = nlo; shi = nhi; sx1 = nx1; sx2 = nx2; q1 = nq1; q2 = nq2;
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(snap_s + 8u * (unsigned) t), "d"(xn) : "memory"); /* (every lane, same value) */
    }
    __syncwarp();
    if (clk && lane == 0) clk[14] = clock64();
    double xmine;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(xmine) : "r"(snap_s + 8u * (unsigned) (lane < L ? lane : 0)));
    const double ynew = lane < L ? hfg_objective_dev(xmine, trunc, sum_x, sum_w) : 0.0;
    if (clk && lane == 0) clk[13] = clock64();
    /* who holds the value at x1 / x2 when step `lane` compares them (0, 1: the round's y1, y2; 2 + t: the point of step t), and
     * after it */
    int ca = 0, cb = 1;
#pragma unroll 1
    for (int t = 0; t < lane; t++) {
        const bool bt = (bits >> t) & 1u;
        const int n1 = bt ? 2 + t : cb, n2 = bt ? ca : 2 + t;
        ca = n1;
        cb = n2;
    }
    const int pb = (int) ((bits >> lane) & 1u);
    const int a1 = pb ? 2 + lane : cb, a2 = pb ? ca : 2 + lane;
    /* the evaluated values behind every comparison, then the first step whose prediction was wrong */
    const double ya_s = __shfl_sync(FULLW, ynew, ca < 2 ? 0 : ca - 2), yb_s = __shfl_sync(FULLW, ynew, cb < 2 ? 0 : cb - 2);
    const double ya = ca == 0 ? y1 : ca == 1 ? y2 : ya_s, yb = cb == 0 ? y1 : cb == 1 ? y2 : yb_s;
    const unsigned wrong = __ballot_sync(FULLW, lane < L && (int) (ya > yb) != pb);
    const int m = wrong ? __ffs(wrong) - 1 : L;
    const double n1_s = __shfl_sync(FULLW, ynew, a1 < 2 ? 0 : a1 - 2), n2_s = __shfl_sync(FULLW, ynew, a2 < 2 ? 0 : a2 - 2);
    const double n1 = a1 == 0 ? y1 : a1 == 1 ? y2 : n1_s, n2 = a2 == 0 ? y1 : a2 == 1 ? y2 : n2_s;
    if (m > 0) {
        if (m < L) {
            /* (rare) a wrong prediction at step m: the bracket after step m - 1, by replaying the positions of the first m
             * steps (their predicted outcomes were the true ones) */
            slo = st->lo; sx1 = st->x1; sx2 = st->x2; shi = st->hi; sspan = st->span;
#pragma unroll 1
            for (int t = 0; t < m; t++) {
                const bool b = (bits >> t) & 1u;
                sspan = inv_phi * sspan;
                const double nlo = b ? slo : sx1, nhi = b ? sx2 : shi;
                const double xn = nlo + (b ? inv_phi2 : inv_phi) * sspan;
                const double nx1 = b ? xn : sx2, nx2 = b ? sx1 : xn;
                slo = nlo; shi = nhi; sx1 = nx1; sx2 = nx2;
            }
        }
        st->lo = slo;
        st->x1 = sx1;
        st->x2 = sx2;
        st->hi = shi;
        st->span = sspan;
        st->y1 = __shfl_sync(FULLW, n1, m - 1);
        st->y2 = __shfl_sync(FULLW, n2, m - 1);
    }
    __syncwarp();
    return m;
}

/* The golden-section rate fit of hfg_mstep_inl.h (hfg_fit_rate / hfg_fit_rate_warp) for a GROUP of `gsize` consecutive
 * threads (a power of two >= 32), with the bits of the serial routine.  The point evaluated at a step depends only on the
 * outcomes of the comparisons so far, so a step's objective value can be computed before the comparisons that lead to it:
 *   - TREE round: the 2 + 4 + ... + 2^D candidate points of the next D steps (D = 7 for 256 threads) one per thread, then
 *     every thread walks the true path through the group's buffer `fb` (gsize doubles);
 *   - PREDICTED round (hfg_fit_pred_round_dev): the objective is concave with its maximum at x*, which a few Newton steps
 *     inside the bracket give; a 4th-order Taylor polynomial around x* predicts the comparisons of up to 32 steps, which
 *     are then all evaluated at once and checked.  A wrong prediction (two values closer than the polynomial's error) ends
 *     the round there and the next round is a TREE round.  Every warp of the group does the same work on the same values.
 * ~34 dependent objective evaluations become 1 TREE round + 1 PREDICTED round in the common case.  The group synchronises
 * on its own named barrier `bar`: all gsize threads of the group must call it, other groups and warps are not involved.
 * `st8` = 24 doubles of group scratch; st8[7] receives tree rounds | predicted rounds << 8 | serial steps << 16; `snap` = 256 doubles for the predicted rounds.  `x_hint`: where Newton starts
 * (the rate of the previous EM iteration).
 * Scouting (see above): start_mode 1 skips the first two evaluations; max_rounds bounds the rounds. */
__device__ __forceinline__ void group_sync(int bar, int n) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __noinline__ double hfg_fit_rate_cta(bool active, int gtid, int gsize, int bar, double trunc, double sum_x, double sum_w,
                                                double *st8, double *fb, bool predict, long long *clk, int start_mode, int max_rounds, double x_hint, double *snap) {
    const double inv_phi = (sqrt(5.0) - 1.0) / 2.0, inv_phi2 = (3.0 - sqrt(5.0)) / 2.0;
    const int lane = gtid & 31;
    int D = 0;
    while ((4 << D) - 2 <= gsize) D++; /* 2^(D+1) - 2 nodes fit the group */
    HfgFitState s;
    s.lo = 0.0;
    s.hi = trunc;
    s.span = s.hi - s.lo;
    double result = (s.hi + s.lo) / 2.0;
    if (active && !(s.span > GOLDEN_TOL)) active = false; /* hmm_utils.c: the interval is already short */
    int total = 0;
    s.x1 = s.lo + inv_phi2 * s.span;
    s.x2 = s.lo + inv_phi * s.span;
    s.y1 = 1.0;
    s.y2 = 2.0;
    if (active) {
        const int steps = (int) ceil(log(GOLDEN_TOL / s.span) / log(inv_phi));
        total = steps - 1;
        if (gtid < 2 && start_mode == 0) st8[gtid] = hfg_objective_dev(gtid == 0 ? s.x1 : s.x2, trunc, sum_x, sum_w);
        if (predict && start_mode == 0 && gsize >= 64 && gtid >= 32 && gtid < 64) {
            /* x* and the polynomial on the group's second warp, beside the first two evaluations */
            double c[5];
            const double x0 = x_hint > s.lo && x_hint < s.hi ? x_hint : 0.5 * (s.lo + s.hi);
            hfg_fit_newton_dev(x0, s.lo, s.hi, trunc, sum_x, sum_w, c);
            if (lane == 0) {
                st8[2] = c[0]; st8[3] = c[1]; st8[4] = c[2]; st8[5] = c[3]; st8[6] = c[4];
            }
        }
    }
    group_sync(bar, gsize);
    if (active && start_mode == 0) {
        s.y1 = st8[0];
        s.y2 = st8[1];
    }
    if (clk && gtid == 0) clk[12] = clock64();
    /* node -> (step t of the round, outcomes b_0..b_t of its comparisons, b_0 in the top bit) */
    int my_t = 0;
    while ((4 << my_t) - 2 <= gtid) my_t++;
    const int my_bits = gtid - ((2 << my_t) - 2);
    int k = 0, n_tree = 0, n_pred = 0;
    bool tree = !predict, have_star = false; /* predicted rounds first: x* from the previous rate */
    double co[5];
    if (active && predict && start_mode == 0 && gsize >= 64) {
        have_star = true;
        co[0] = st8[2]; co[1] = st8[3]; co[2] = st8[4]; co[3] = st8[5]; co[4] = st8[6];
    }
    int n_serial = 0;
    /* (active, total, k and tree are the same in every thread of the group: uniform loop) */
    while (active && k < total && n_tree + n_pred + n_serial < max_rounds) {
        if (tree || !predict) {
            const int d = total - k < D ? total - k : D;
            if (my_t < d) {
                /* positions along this node's assumed path (the arithmetic of the serial loop, same order) */
                double slo = s.lo, sx1 = s.x1, sx2 = s.x2, sspan = s.span, xnew = s.x1;
                for (int t = 0; t <= my_t; t++) {
                    const int b = (my_bits >> (my_t - t)) & 1;
                    sspan = inv_phi * sspan;
                    if (b) {
                        sx2 = sx1;
                        sx1 = slo + inv_phi2 * sspan;
                        xnew = sx1;
                    } else {
                        slo = sx1;
                        sx1 = sx2;
                        sx2 = slo + inv_phi * sspan;
                        xnew = sx2;
                    }
                }
                fb[gtid] = hfg_objective_dev(xnew, trunc, sum_x, sum_w);
            }
            group_sync(bar, gsize);
            /* the true path */
            int path = 0;
            for (int t = 0; t < d; t++) {
                const int b = s.y1 > s.y2;
                path = (path << 1) | b;
                const double y = fb[((2 << t) - 2) + path];
                s.span = inv_phi * s.span;
                if (b) {
                    s.hi = s.x2; s.x2 = s.x1; s.y2 = s.y1;
                    s.x1 = s.lo + inv_phi2 * s.span;
                    s.y1 = y;
                } else {
                    s.lo = s.x1; s.x1 = s.x2; s.y1 = s.y2;
                    s.x2 = s.lo + inv_phi * s.span;
                    s.y2 = y;
                }
            }
            k += d;
            n_tree++;
            tree = false;
            group_sync(bar, gsize); /* the buffer is read before the next round writes it */
            continue;
        }
        if (have_star && n_pred > 0 && total - k <= 2) {
            /* one or two steps left over by a full predicted round: the serial routine's steps as they are */
            const int b = s.y1 > s.y2;
            s.span = inv_phi * s.span;
            if (b) {
                s.hi = s.x2; s.x2 = s.x1; s.y2 = s.y1;
                s.x1 = s.lo + inv_phi2 * s.span;
                s.y1 = hfg_objective_dev(s.x1, trunc, sum_x, sum_w);
            } else {
                s.lo = s.x1; s.x1 = s.x2; s.y1 = s.y2;
                s.x2 = s.lo + inv_phi * s.span;
                s.y2 = hfg_objective_dev(s.x2, trunc, sum_x, sum_w);
            }
            k++;
            n_serial++;
            continue;
        }
        /* warp 0 of the group runs the round, the others pick the result up from the group's scratch (two buffers in turn) */
        const int L = total - k < 32 ? total - k : 32;
        double *buf = st8 + 8 + 8 * (n_pred & 1);
        if (gtid < 32) {
            if (!have_star) {
                double x0 = n_tree ? (s.y1 > s.y2 ? s.x1 : s.x2) : x_hint;
                if (!(x0 > s.lo && x0 < s.hi)) x0 = 0.5 * (s.lo + s.hi);
                hfg_fit_newton_dev(x0, s.lo, s.hi, trunc, sum_x, sum_w, co);
            }
            const int mm = hfg_fit_pred_round_dev(&s, co, trunc, sum_x, sum_w, lane, L, snap, n_pred == 0 ? clk : nullptr);
            if (lane == 0) {
                buf[0] = s.lo; buf[1] = s.x1; buf[2] = s.x2; buf[3] = s.hi; buf[4] = s.span; buf[5] = s.y1; buf[6] = s.y2;
                buf[7] = (double) mm;
            }
        }
        have_star = true;
        group_sync(bar, gsize);
        s.lo = buf[0]; s.x1 = buf[1]; s.x2 = buf[2]; s.hi = buf[3]; s.span = buf[4]; s.y1 = buf[5]; s.y2 = buf[6];
        const int m = (int) buf[7];
        k += m;
        if (clk && gtid == 0 && n_pred == 0) clk[15] = clock64();
        n_pred++;
        tree = m < L; /* the step the polynomial got wrong is taken by a TREE round */
    }
    if (active) result = s.y1 > s.y2 ? (s.lo + s.x2) / 2.0 : (s.x1 + s.hi) / 2.0;
    if (gtid == 0) st8[7] = (double) (n_tree | (n_pred << 8) | (n_serial << 16));
    return result;
}

/* hfg_mstep_gauss / hfg_mstep_trans (hfg_mstep_inl.h) for one WARP: one lane per mixture component / transition entry.  The
 * pooled sums are added in the serial routine's order (state, then component, ascending) and every division is the serial
 * routine's division, so the parameters come out with the same bits; what runs side by side are the divisions and the
 * convergence tests (a single thread needs ~20 000 cycles for them, a chain of ~100 dependent double-precision divisions).
 * `scr` = 128 doubles of warp-private shared memory.  Return value as the serial routines, identical in every lane. */
__device__ __forceinline__ int hfg_mstep_gauss_warp(int model_type, const int32_t *n_comps, hfg_region_params *p,
                                                    const hfg_region_stats *st, double tol, double *scr, int lane) {
    int ok = 1, G = 0, gs[2] = {0, 0}, gc[2] = {0, 0};
    /* the lane's components: g = lane and lane + 32 of the list (state ascending, component ascending) */
    for (int s = 0; s < HFG_NS; s++) {
        if (!hfg_is_gaussian_state(model_type, s)) continue;
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g >= G && g < G + n_comps[s]) {
                gs[u] = s;
                gc[u] = g - G;
            }
        }
        G += n_comps[s];
    }
    for (int which = 0; which < 2; which++) {
        const double (*num)[HFG_MAX_COMPS] = which == 0 ? st->mean_num : st->var_num;
        const double (*den)[HFG_MAX_COMPS] = which == 0 ? st->mean_den : st->var_den;
        double (*dst)[HFG_MAX_COMPS] = which == 0 ? p->mean : p->var;
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g < G) {
                scr[g] = num[gs[u]][gc[u]] / hfg_binding(gs[u], gc[u]);
                scr[64 + g] = den[gs[u]][gc[u]];
            }
        }
        __syncwarp();
        double pooled_num = 0.0, pooled_den = 0.0;
        for (int g = 0; g < G; g++) {
            pooled_num += scr[g];
            pooled_den += scr[64 + g];
        }
        __syncwarp();
        if (!(MIN_COUNT_FOR_UPDATE < pooled_den)) continue; /* hmm_utils.c:1846 */
        const double unit = pooled_num / pooled_den;
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            if (g < G) {
                const double v = unit * hfg_binding(gs[u], gc[u]);
                ok &= hfg_settled(dst[gs[u]][gc[u]], v, tol, 1.0e-4);
                dst[gs[u]][gc[u]] = v;
            }
        }
    }
    /* mixture weights: unbound, each component from its own estimator */
    for (int u = 0; u < 2; u++) {
        const int g = lane + 32 * u;
        if (g < G) {
            const double d = st->weight_den[gs[u]][gc[u]];
            if (MIN_COUNT_FOR_UPDATE < d) {
                const double v = st->weight_num[gs[u]][gc[u]] / d;
                ok &= hfg_settled(p->weight[gs[u]][gc[u]], v, tol, 1.0e-4);
                p->weight[gs[u]][gc[u]] = v;
            }
        }
    }
    __syncwarp();
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ int hfg_mstep_trans_warp(hfg_region_params *p, const hfg_region_stats *st, double tol, int lane) {
    int ok = 1;
    if (lane < HFG_NS * HFG_NS) {
        const int a = lane >> 2, b = lane & 3;
        double row = 0.0;
        for (int i = 0; i < HFG_NS; i++) row += st->trans_count[a][i] + PSEUDO_COUNT;
        const double v = (st->trans_count[a][b] + PSEUDO_COUNT) / row * (1.0 - HFG_TERM_PROB);
        ok &= hfg_settled(p->trans[a][b], v, tol, 1.0e-6);
        p->trans[a][b] = v;
        if (b == 0) p->trans[a][HFG_NS] = HFG_TERM_PROB;
        if (a == 0) p->trans[HFG_NS][b] = 1.0 / HFG_NS;
        if (lane == 0) p->trans[HFG_NS][HFG_NS] = 0.0;
    }
    __syncwarp();
    return __all_sync(0xffffffffu, ok);
}

/* Phase D, shared by the kernel generations that end in per-CTA partials [R][NSTAT][grid]: the last CTA to arrive (atomic
 * ticket) sums them in a fixed order, writes the hfg_region_stats block, exchanges it with the other ranks (multi-GPU),
 * runs the M-step of the device-resident loop and clears the flags.  Every other CTA returns at once.  `wstat` (>= R * NSTAT
 * doubles) and `work` (the M-step work area, sized by the host) are shared memory. */
template <int THREADS, bool NB = false>
__device__ __forceinline__ void hfg_estep_tail(const EstepArgs &A, double *wstat, double *work, int *s_last) {
    constexpr int WARPS = THREADS / 32;
    using namespace hfgq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = A.n_regions, NSTAT = hfg_nstat(A.G);
    /* =========================== phase D: grid reduction by the last CTA to arrive ========================== */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 6] = clock64();
    __syncthreads();
    if (tid == 0) {
        /* one cumulative fence behind the CTA barrier (the pattern of a cooperative-groups grid barrier) orders every thread's
         * partials -- and, at system scope when the host polls a completion word, the labels this CTA streamed into host
         * memory -- before the ticket */
        if (A.done_flag) __threadfence_system();
        else __threadfence();
        const int ticket = atomicAdd(A.ticket, 1);
        *s_last = ticket == (int) gridDim.x - 1;
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    /* SCOUT warps (see hfg_fit_rate_cta): the last four warps of a large CTA pre-execute the four pieces of the rate fit on dummy
     * numbers while the others -- the NW worker threads, which synchronise on named barrier 14 -- run the tail proper */
    constexpr int SCOUTS = THREADS >= 512 ? 4 : 0;
    const int PER_REGION = (int) ((sizeof(hfg_region_params) + sizeof(hfg_region_stats)) / sizeof(double)) + 282;
    const bool scouting = SCOUTS > 0 && A.em_mode == 1 && A.model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && !(A.dbg & 32) &&
                          A.work_doubles - HFG_TAIL_TOP >= PER_REGION + THREADS + WARPS * 128;
    const int NW = scouting ? THREADS - 32 * SCOUTS : THREADS;
    auto tsync = [&]() { asm volatile("bar.sync 14, %0;" ::"r"(NW) : "memory"); };
    if (tid >= NW) {
        const int sc = warp - NW / 32;
        double *scr = work + A.work_doubles - HFG_TAIL_TOP + sc * 64; /* 24 + 32 doubles of scratch per scout, out of the workers' way */
        double sink = 0.0;
        if (sc == 0) sink = hfg_fit_rate_cta(true, lane, 32, 15, 7.25, 1000.0, 2000.0, scr, scr + 24, false, nullptr, 0, 0, 0.0, nullptr);
        if (sc == 1) sink = hfg_fit_rate_cta(true, lane, 32, 13, 7.25, 1000.0, 2000.0, scr, scr + 24, false, nullptr, 1, 1, 0.0, nullptr);
        if (sc == 2) {
            double co[5];
            hfg_fit_newton_dev(1.7, 1.6, 1.9, 7.25, 1000.0, 2000.0, co);
            sink = co[0] + co[4];
        }
        if (sc == 3) {
            HfgFitState st = {1.6, 1.7146, 1.7854, 1.9, 0.3, -1.0, -1.01};
            const double co[5] = {1.74, 1.0e-3, -300.0, 90.0, -40.0};
            sink = (double) hfg_fit_pred_round_dev(&st, co, 7.25, 1000.0, 2000.0, lane, 26, hfg_align16(work + A.work_doubles - HFG_TAIL_TOP + 256), nullptr) + st.lo;
        }
        if (lane == 0) scr[60] = sink;
        if (lane == 0 && sc == 3) (A.phase_clock + (size_t) gridDim.x * HFG_PC_STRIDE)[9] = clock64(); /* the longest scout's end */
        return;
    }
    {
        /* tail clocks live in the row behind the CTAs' rows: 0 tail start, 1 totals in shared memory, 2 statistics block
         * written, 3 / 4 exchange start / end (multi-GPU), 5 M-step done, 6 end, 7 CTA that ran the tail, 8 M-step copy-in
         * done, 9 last scout done, 10 updates done (thread 0), 11 rate-fit rounds of group 0 (tree | predicted << 8 |
         * serial << 16), 12-15 inside group 0's rate fit: first evaluations done, predicted round's evaluation done, its
         * walk done, round done */
        long long *tail_clock = A.phase_clock + (size_t) gridDim.x * HFG_PC_STRIDE;
        if (tid == 0) {
            *A.ticket = 0; /* for the next launch */
            tail_clock[0] = clock64();
            tail_clock[7] = (long long) blockIdx.x;
        }
        const int SD = (int) (sizeof(hfg_region_stats) / sizeof(double));
        const int nb = gridDim.x;
        double *acc = wstat;
        /* four totals per warp at a time: the lanes add the CTAs' partials with stride 32 (all loads of a round are
         * independent and in flight together), then a fixed xor-shuffle tree -- the same association on every run */
        {
            constexpr int QU = 4; /* (measured: 8 totals per warp and round -- 40 loads in flight -- is SLOWER, 20 300 against 16 300 cycles
                                     for 7 regions: the tail runs cold, longer straight-line code costs more instruction fetches) */
            const int NQ = R * NSTAT;
            for (int q0 = warp * QU; q0 < NQ; q0 += (NW / 32) * QU) {
                double sum[QU];
#pragma unroll
                for (int u = 0; u < QU; u++) sum[u] = 0.0;
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
                    sum[u] = warp_sum(sum[u]);
                    if (lane == 0 && q0 + u < NQ) acc[q0 + u] = sum[u]; /* [R][NSTAT] totals */
                }
            }
        }
        tsync();
        if (tid == 0) tail_clock[1] = clock64();
        /* the hfg_region_stats layout (include/hfg.h), one output element per thread: no zero-fill, no read-back */
        for (int qq = tid; qq < R * SD; qq += NW) {
            const int r = qq / SD, i = qq % SD;
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
                        /* weight_den[s][c'] = sum over the components of s (ParameterEstimator_incrementDenominatorForAllComps) */
                        for (int c2 = 0; c2 < A.ncomp[s]; c2++) v += tot[18 + 3 * (A.gbase[s] + c2) + 1];
                    }
                }
            }
            A.out[qq] = v;
        }
        if (tid == 0) A.out[(size_t) R * SD] = acc[NSTAT - 1]; /* log-likelihood (kept in region 0's row) */
        if (tid == 0) A.out[(size_t) R * SD + 1] = (double) __ldcg(A.err_flags);
        if (tid == 0) tail_clock[2] = clock64();

        /* ---- negative binomial, device-resident loop: histogram -> estimator sums into the statistics block ---- */
        if constexpr (NB) {
            if (A.em_mode == 1) {
                tsync();
                const int nan = hfgnb::tail_estimators(A.params, R, A.ncomp, A.nb_hist, A.nb_lgx1, A.out, work, tid, NW, tsync, tail_clock);
                if (nan) atomicOr(A.err_flags, 2);
                tsync();
                if (tid == 0) A.out[(size_t) R * SD + 1] = (double) __ldcg(A.err_flags);
            }
        }

        /* ---- fused collective: sum the block over all ranks through peer memory ---------------------------------------
         * Every 64-bit word a rank stores into a peer's mailbox carries 32 bits of data and the 32-bit epoch of the exchange,
         * so data and "it has arrived" travel in ONE store over NVLink (NCCL's LL idea): no fence, no separate counter, no
         * second trip.  Sender: thread i puts the two halves of out[i] into slot [epoch & 1][own rank][i] of every peer.
         * Receiver: thread i spins on the two words of [epoch & 1][p][i] for p = 0 .. N-1 (its own value comes straight from
         * out[i]) and adds them in rank order -- the same order on every rank, so every rank's M-step sees the same bits.
         * Two epochs' slots suffice: a rank can start exchange e + 2 only after every peer has sent e + 1, i.e. has finished
         * reading e. */
        if (A.n_ranks > 1) {
            tsync();
            const int n = A.out_doubles, N = A.n_ranks;
            const unsigned long long e = *A.epoch + 1;
            const unsigned long long flag = (e & 0xffffffffull) << 32;
            const size_t ll_base = (size_t) 2 * HFG_MAX_PEERS * n + 3 * HFG_MAX_PEERS; /* doubles before the flagged slots */
            const size_t slot0 = ((size_t) (e & 1) * HFG_MAX_PEERS) * n * 2;            /* words */
            if (tid == 0) tail_clock[3] = clock64();
            for (int qq = tid; qq < n; qq += NW) {
                const unsigned long long v = (unsigned long long) __double_as_longlong(A.out[qq]);
                const unsigned long long w0 = (v & 0xffffffffull) | flag, w1 = (v >> 32) | flag;
                for (int p = 0; p < N; p++) {
                    if (p == A.rank) continue;
                    unsigned long long *dst =
                        reinterpret_cast<unsigned long long *>(A.peer_box[p] + ll_base) + slot0 + ((size_t) A.rank * n + qq) * 2;
                    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
                }
            }
            const unsigned long long *in = reinterpret_cast<const unsigned long long *>(A.peer_box[A.rank] + ll_base) + slot0;
            const long long t0 = clock64();
            bool lost = false;
            for (int qq = tid; qq < n; qq += NW) {
                double sum = 0.0;
                int fl = 0;
                for (int p = 0; p < N; p++) {
                    double v;
                    if (p == A.rank) {
                        v = A.out[qq];
                    } else {
                        const unsigned long long *src = in + ((size_t) p * n + qq) * 2;
                        unsigned long long w0, w1;
                        for (;;) {
                            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
                            if ((w0 & 0xffffffff00000000ull) == flag && (w1 & 0xffffffff00000000ull) == flag) break;
                            if (clock64() - t0 > 4000000000LL) { /* ~2 s: a lost peer becomes an error, not a hang */
                                lost = true;
                                break;
                            }
                        }
                        v = __longlong_as_double((long long) ((w0 & 0xffffffffull) | (w1 << 32)));
                    }
                    if (qq == n - 1) fl |= (int) v; /* error flags: bitwise OR over ranks */
                    else sum += v;
                }
                A.out[qq] = qq == n - 1 ? (double) fl : sum;
            }
            if (lost) atomicOr(A.err_flags, 4);
            tsync();
            if (tid == 0) {
                if (__ldcg(A.err_flags) & 4) A.out[n - 1] = (double) ((int) A.out[n - 1] | 4);
                *A.epoch = e;
                tail_clock[4] = clock64();
            }
        }

        /* ---- device-resident EM: M-step of every region, bookkeeping ------------------------------------------ */
        if (A.em_mode) {
            tsync();
            const int flags = (int) A.out[(size_t) R * SD + 1];
            int settled = 1;
            if (A.em_mode == 1 && flags == 0) {
                /* HMM_estimateParameters on shared-memory copies of the parameters and statistics, a batch of regions at a
                 * time: Gaussian parameters and transition rows one warp each (every lane computes and stores the same
                 * values), then the rate fits of the batch side by side, each on a power-of-two group of threads
                 * (hfg_fit_rate_cta).  The pieces touch disjoint parameters; the rate fit uses the OLD truncation point, which
                 * follows the new Hap mean afterwards. */
                constexpr int PD = (int) (sizeof(hfg_region_params) / sizeof(double));
                const int per_region = PD + SD + 282; /* 24 doubles of group scratch + [32][8] of the predicted rounds, 16-byte aligned */
                constexpr int FIT_THREADS = THREADS / 2 >= 256 ? THREADS / 2 : THREADS - 64; /* the rest runs the other updates */
                int RB = (A.work_doubles - HFG_TAIL_TOP - THREADS - WARPS * 128) / per_region;
                RB = RB < 1 ? 1 : (RB > R ? R : RB);
                if (RB > FIT_THREADS / 32) RB = FIT_THREADS / 32;
                if (RB > 12) RB = 12; /* named barriers 1..12 (13-15: the scouts' and the workers') */
                for (int r0 = 0; r0 < R; r0 += RB) {
                    const int nb = min(RB, R - r0);
                    double *mscr = work + (size_t) nb * per_region; /* [WARPS][128] warp scratch, then the exchange buffers */
                    for (int i = tid; i < nb * PD; i += NW)
                        work[(size_t) (i / PD) * per_region + i % PD] = reinterpret_cast<const double *>(&A.em_params[r0])[i];
                    for (int i = tid; i < nb * SD; i += NW) work[(size_t) (i / SD) * per_region + PD + i % SD] = A.out[(size_t) r0 * SD + i];
                    tsync();
                    if (tid == 0) tail_clock[8] = clock64();
                    /* the rate fits of the batch on the first fit_threads threads (one power-of-two group each, own named
                     * barrier), the Gaussian / transition updates on the remaining warps, side by side */
                    int gsize = 32;
                    while (gsize * 2 * nb <= FIT_THREADS) gsize *= 2;
                    if constexpr (NB) {
                        /* negative binomial: the host's M-step (hfg_nb_mstep_inl.h), emission and transition halves one warp each */
                        for (int task = warp; task < 2 * nb; task += NW / 32) {
                            double *mp = work + (size_t) (task >> 1) * per_region;
                            hfg_region_params *p = reinterpret_cast<hfg_region_params *>(mp);
                            const hfg_region_stats *st = reinterpret_cast<const hfg_region_stats *>(mp + PD);
                            if (task & 1) settled &= hfg_mstep_trans_warp(p, st, A.em_tol, lane);
                            else settled &= hfgnb::mstep_emis_warp(A.ncomp, p, st, A.em_tol, mscr + (size_t) warp * 128, lane);
                        }
                    } else if (tid < nb * gsize) {
                        const int g = tid / gsize, gtid = tid % gsize;
                        double *mp = work + (size_t) g * per_region;
                        hfg_region_params *p = reinterpret_cast<hfg_region_params *>(mp);
                        const hfg_region_stats *st = reinterpret_cast<const hfg_region_stats *>(mp + PD);
                        /* TruncExponential: golden-section fit against the truncation point still in force (hmm_utils.c:1872-1882) */
                        const bool fit = A.model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN && MIN_COUNT_FOR_UPDATE < st->lambda_den;
                        const double trunc = p->trunc_point, sx = st->lambda_num, sw = st->lambda_den, old = p->lambda;
                        const double v = hfg_fit_rate_cta(fit, gtid, gsize, 1 + g, trunc, sx, sw, mp + PD + SD, mscr + WARPS * 128 + (size_t) g * gsize,
                                                          !(A.dbg & 16), g == 0 ? tail_clock : nullptr, 0, 1000, old, hfg_align16(mp + PD + SD + 24));
                        if (tid == 0) tail_clock[11] = (long long) mp[PD + SD + 7];
                        if (fit) {
                            settled &= hfg_settled(old, v, A.em_tol, 1.0e-4);
                            if (gtid == 0) p->lambda = v;
                        }
                    } else if (tid >= FIT_THREADS) {
                        for (int task = warp - FIT_THREADS / 32; task < 2 * nb; task += NW / 32 - FIT_THREADS / 32) {
                            double *mp = work + (size_t) (task >> 1) * per_region;
                            hfg_region_params *p = reinterpret_cast<hfg_region_params *>(mp);
                            const hfg_region_stats *st = reinterpret_cast<const hfg_region_stats *>(mp + PD);
                            if (task & 1) settled &= hfg_mstep_trans_warp(p, st, A.em_tol, lane);
                            else settled &= hfg_mstep_gauss_warp(A.model_type, A.ncomp, p, st, A.em_tol, mscr + (size_t) warp * 128, lane);
                        }
                    }
                    if (tid == 0) tail_clock[10] = clock64();
                    tsync();
                    /* the truncation point follows the NEW Hap mean, after the fit used the old one */
                    if (tid < nb && A.model_type == HFG_MODEL_TRUNC_EXP_GAUSSIAN) {
                        hfg_region_params *p = reinterpret_cast<hfg_region_params *>(work + (size_t) tid * per_region);
                        p->trunc_point = p->mean[HFG_STATE_HAP][0] * TRUNC_POINT_FRACTION;
                    }
                    tsync();
                    for (int i = tid; i < nb * PD; i += NW)
                        reinterpret_cast<double *>(&A.em_params[r0])[i] = work[(size_t) (i / PD) * per_region + i % PD];
                    tsync();
                }
            }
            /* AND over the worker threads (the scouts are gone): through shared memory */
            int *s_settled = reinterpret_cast<int *>(work + A.work_doubles - 1);
            if (tid == 0) *s_settled = 1;
            tsync();
            if (!__all_sync(0xffffffffu, settled) && lane == 0) atomicAnd(s_settled, 0);
            tsync();
            const int all_settled = *s_settled;
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
                tail_clock[5] = clock64();
            }
        }
        /* results straight into the caller-visible pinned block, error flags cleared for the next launch */
        tsync();
        if (A.out_host)
            for (int qq = tid; qq < A.out_doubles; qq += NW) A.out_host[qq] = A.out[qq];
        if (tid == 0) {
            *A.err_flags = 0;
            tail_clock[6] = clock64();
        }
        if (A.done_flag) {
            /* every result of this call -- the labels of all CTAs (ordered by their fences and the ticket), the block above --
             * precedes the word the host polls */
            __threadfence_system();
            tsync();
            if (tid == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.done_flag), "l"(A.done_seq) : "memory");
        }
    }
}

template <int THREADS, bool NB = false>
__global__ void __launch_bounds__(THREADS, 1) hfg_estep_quad_kernel(const EstepArgs A) {
    constexpr int WARPS = THREADS / 32, QUADS = THREADS / 4;
    using namespace hfgq;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];

    /* device-resident EM: the previous launch raised the stop flag: nothing to do (read by every thread before any barrier) */
    if (A.em_mode == 1 && __ldcg(&A.em_state[0]) != 0) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = tid & 3, qbase = lane & ~3, quad = lane >> 2;
    const int j = blockIdx.x * QUADS + (tid >> 2); /* segment owned by this quad */
    const int cap = A.capacity, G = A.G, R = A.n_regions;
    const int rt_stride = QRT_STRIDE(G);
    const int NSTAT = hfg_nstat(G);

    /* shared memory carve-up (doubles) */
    double *rtab = smem;                                            /* [R][rt_stride] */
    double *ks = rtab + (((size_t) R * rt_stride + 1) & ~(size_t) 1); /* [2 directions][2 buffers][WARPS][16] second-level scan */
    double *blk_vec = ks + 4 * WARPS * 16;                          /* [8] forward / backward message entering the CTA */
    double *qstash = blk_vec + 8;                                   /* [QUADS][8] messages entering the segment (v_in, u_in): parked
                                                                       here while the sweeps run, to keep them out of registers */
    double *wstat = qstash + QUADS * 8;                             /* [WARPS][NSTAT] statistics per warp; tail: [R][NSTAT] totals */
    int8_t *lab_s = reinterpret_cast<int8_t *>(wstat + (((size_t) (WARPS > R ? WARPS : R) * NSTAT + 1) & ~(size_t) 1)); /* [QUADS * smax + 16] */
    __shared__ int s_reset, s_last;

    /* ---- prologue: derived per-region tables (redundantly per CTA; O(R*K) work) ---- */
    if (tid == 0) s_reset = 0;
    for (int idx = tid; idx < R * 32; idx += THREADS) {
        /* one (region, mask, pre) row of the conditional transition table (Transition_getProbConditional,
         * hmm_utils.c:2278-2292) */
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
    for (int idx = (tid + THREADS - 256) % THREADS; idx < R * (12 + G); idx += THREADS) {
        const int r = idx / (12 + G), qq = idx % (12 + G);
        const hfg_region_params &p = A.params[r];
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
    /* one quad per key, keys dealt round-robin over the CTAs; lane pre evaluates the emission of every state under its own
     * alpha[pre][s] (what the reference does per window and (pre, s) pair, hmm.c:386-408) and writes row pre of the table
     * and column pre of the transposed table */
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
        double *dt = A.tabMT + (size_t) p * 16 + q;
#pragma unroll
        for (int s = 0; s < 4; s++) dt[s * 4] = row[s];
    }
    grid.sync(); /* the key tables are complete */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 1] = clock64();

    const int len = A.seg_len[j];
    const int seg_first = A.seg_start[j];
    const uint32_t *wk = A.wkeyT + j; /* wk[k * cap]: key word of the k-th window of this segment */
    const int kmax = __reduce_max_sync(FULL, len); /* longest segment of this warp: the step loops are warp-uniform */

    /* =========================== phase A: segment transfer product, row q =================================== */
    /* P <- P * M: lane q holds row q of P and fetches row q of M (one line per quad); the other three rows arrive by
     * shuffles.  Key words are fetched two steps ahead, rows one step ahead. */
    double P[4];
#pragma unroll
    for (int c = 0; c < 4; c++) P[c] = c == q ? 1.0 : 0.0;
    {
        bool has_start = false;
        uint32_t w0 = len > 0 ? __ldg(wk) : 0u;
        uint32_t w1 = len > 1 ? __ldg(wk + cap) : 0u;
        double Mc[4], Mn[4];
        ld_row(A.tabM + (size_t) HFG_KEY_ID(w0) * 16 + q * 4, Mc);
#pragma unroll 1
        for (int k = 0; k < kmax; k++) {
            const uint32_t w2 = k + 2 < len ? __ldg(wk + (size_t) (k + 2) * cap) : 0u;
            ld_row(A.tabM + (size_t) HFG_KEY_ID(w1) * 16 + q * 4, Mn);
            double o[4];
            {
                const double pq = sel4(P, q);
#pragma unroll
                for (int c = 0; c < 4; c++) o[c] = pq * Mc[c];
            }
#pragma unroll
            for (int d = 1; d < 4; d++) {
                const double pd = sel4(P, q ^ d); /* entry q^d of my row multiplies row q^d of M, held by lane q^d */
#pragma unroll
                for (int c = 0; c < 4; c++) o[c] = fma(pd, __shfl_xor_sync(FULL, Mc[c], d), o[c]);
            }
            if (k < len) {
                if (w0 & HFG_KEY_CHUNK_START) has_start = true;
#pragma unroll
                for (int c = 0; c < 4; c++) P[c] = o[c];
            }
            if ((k & 3) == 3) quad_rescale(P); /* a window shrinks the product by < 1e-50: every 4th step is ample */
#pragma unroll
            for (int c = 0; c < 4; c++) Mc[c] = Mn[c];
            w0 = w1;
            w1 = w2;
        }
        quad_rescale(P);
        if (has_start) s_reset = 1; /* benign race: every writer stores 1 */
    }

    /* =========================== phase B: scans ============================================================== */
    /* level 1, the 8 segments of a warp: inclusive prefix / suffix products by Kogge-Stone over shuffles */
    double Ppre[4], Psuf[4];
#pragma unroll
    for (int c = 0; c < 4; c++) Ppre[c] = Psuf[c] = P[c];
#pragma unroll 1
    for (int off = 1; off < 8; off <<= 1) {
        double a[4], o[4];
        /* prefix: P_g <- P_{g-off} * P_g : row q of the left factor comes from the quad `off` below, the right factor is
         * this quad's own matrix */
#pragma unroll
        for (int c = 0; c < 4; c++) a[c] = __shfl_up_sync(FULL, Ppre[c], 4 * off);
        row_times_quadmat(a, Ppre, qbase, o);
        if (quad >= off) {
#pragma unroll
            for (int c = 0; c < 4; c++) Ppre[c] = o[c];
        }
        quad_rescale(Ppre);
        /* suffix: P_g <- P_g * P_{g+off} */
        row_times_quadmat(Psuf, Psuf, qbase + 4 * off, o);
        if (quad + off < 8) {
#pragma unroll
            for (int c = 0; c < 4; c++) Psuf[c] = o[c];
        }
        quad_rescale(Psuf);
    }
    /* the warp's product = inclusive prefix of its last quad = inclusive suffix of its first: seeds of the second level */
    if (quad == 7) {
        double2 *d = reinterpret_cast<double2 *>(ks + (size_t) warp * 16 + q * 4);
        d[0] = make_double2(Ppre[0], Ppre[1]);
        d[1] = make_double2(Ppre[2], Ppre[3]);
        d = reinterpret_cast<double2 *>(ks + (size_t) (2 * WARPS + warp) * 16 + q * 4);
        d[0] = make_double2(Ppre[0], Ppre[1]);
        d[1] = make_double2(Ppre[2], Ppre[3]);
    }
    /* exclusive products of this quad inside its warp */
    {
        double t[4];
#pragma unroll
        for (int c = 0; c < 4; c++) t[c] = __shfl_up_sync(FULL, Ppre[c], 4);
#pragma unroll
        for (int c = 0; c < 4; c++) Ppre[c] = quad == 0 ? (c == q ? 1.0 : 0.0) : t[c];
#pragma unroll
        for (int c = 0; c < 4; c++) t[c] = __shfl_down_sync(FULL, Psuf[c], 4);
#pragma unroll
        for (int c = 0; c < 4; c++) Psuf[c] = quad == 7 ? (c == q ? 1.0 : 0.0) : t[c];
    }
    __syncthreads();
    /* level 2, the WARPS warp products of this CTA: Kogge-Stone through shared memory, one matrix per quad (rows over its
     * four lanes); threads [0, 4*WARPS) build the inclusive prefixes, [4*WARPS, 8*WARPS) the inclusive suffixes.  Buffers:
     * ks[(dir * 2 + buf) * WARPS + i][16]. */
    int ks_fin = 0;
    {
        const int dir = tid >= 4 * WARPS ? 1 : 0;
        const int i = (tid - dir * 4 * WARPS) >> 2; /* matrix index (meaningful for tid < 8*WARPS) */
        const bool scanner = tid < 8 * WARPS;
        int cur = 0;
#pragma unroll 1
        for (int off = 1; off < WARPS; off <<= 1) {
            if (scanner) {
                const double *src = ks + (size_t) (dir * 2 + cur) * WARPS * 16;
                double a[4], o[4];
                const double *own = src + (size_t) i * 16;
                bool upd;
                if (dir == 0) {
                    upd = i >= off;
                    const double *left = src + (size_t) (upd ? i - off : i) * 16 + q * 4;
#pragma unroll
                    for (int c = 0; c < 4; c++) a[c] = left[c];
                    row_times_mat(a, own, o);
                } else {
                    upd = i + off < WARPS;
#pragma unroll
                    for (int c = 0; c < 4; c++) a[c] = own[q * 4 + c];
                    row_times_mat(a, src + (size_t) (upd ? i + off : i) * 16, o);
                }
                if (!upd) {
#pragma unroll
                    for (int c = 0; c < 4; c++) o[c] = own[q * 4 + c];
                }
                quad_rescale(o);
                double2 *d = reinterpret_cast<double2 *>(ks + ((size_t) (dir * 2 + (cur ^ 1)) * WARPS + i) * 16 + q * 4);
                d[0] = make_double2(o[0], o[1]);
                d[1] = make_double2(o[2], o[3]);
            }
            cur ^= 1;
            __syncthreads();
        }
        ks_fin = cur;
    }
    const double *ks_pre = ks + (size_t) (0 * 2 + ks_fin) * WARPS * 16; /* inclusive prefix products of the warps */
    const double *ks_suf = ks + (size_t) (1 * 2 + ks_fin) * WARPS * 16; /* inclusive suffix products */
    if (tid < 16) {
        A.block_tot[(size_t) blockIdx.x * 16 + tid] = ks_pre[(size_t) (WARPS - 1) * 16 + tid];
        if (tid == 0) A.block_reset[blockIdx.x] = s_reset;
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

    /* messages entering this segment: (CTA message) x (exclusive product of the warps before) x (exclusive product of the
     * quads before), and the mirror image for the backward direction */
    double *stash = qstash + (size_t) (tid >> 2) * 8; /* v_in | u_in, read back after a __syncwarp */
    double f[4];
    {
        double t[4], o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = blk_vec[i];
        if (warp > 0) {
            row_times_mat(t, ks_pre + (size_t) (warp - 1) * 16, o);
            vec_rescale(o);
#pragma unroll
            for (int i = 0; i < 4; i++) t[i] = o[i];
        }
        /* v_in[c] = sum_r t[r] * Ppre[r][c]: this lane holds row q, the sum over rows is a sum over the quad's lanes */
        const double tq = sel4(t, q);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double pc = tq * Ppre[c];
            pc += __shfl_xor_sync(FULL, pc, 1);
            pc += __shfl_xor_sync(FULL, pc, 2);
            f[c] = pc;
        }
        vec_normalize(f); /* f^ of the window before the segment: sums to one, as the reference's scaled forward does */

#pragma unroll
        for (int i = 0; i < 4; i++) t[i] = blk_vec[4 + i];
        if (warp < WARPS - 1) {
            mat_times_col(ks_suf + (size_t) (warp + 1) * 16, t, o);
            vec_rescale(o);
#pragma unroll
            for (int i = 0; i < 4; i++) t[i] = o[i];
        }
        /* u_in[r] = sum_c Psuf[r][c] * t[c]: lane r computes its component, the quad exchanges them */
        double ur = fma(Psuf[3], t[3], fma(Psuf[2], t[2], fma(Psuf[1], t[1], Psuf[0] * t[0])));
        const double us = ur + __shfl_xor_sync(FULL, ur, 1);
        ur *= 1.0 / (us + __shfl_xor_sync(FULL, us, 2));
        stash[q] = sel4(f, q);
        stash[4 + q] = ur;
    }

    /* =========================== phase C1: forward inside the segment ======================================== */
    /* Every lane carries the whole forward vector (identical bits in the four lanes); lane s adds column s of M, read as
     * row s of the transposed table.  The vector is carried UNNORMALISED, f_k = (f_{k-1} M_k) 2^{e_k} with e_k != 0 every
     * fourth window: the scales of hmm.c:410-419 are c_k = sum(f_{k-1} M_k) / sum(f_{k-1}), so that
     * sum_k log c_k = log(sum f_last) - ln 2 sum_k e_k (the entering vector sums to one; a chunk start restarts the product
     * with c = sum of its first column, which is the same formula). */
    {
        int esum = 0;
        double fsum = 1.0; /* sum of the current vector */
        uint32_t w0 = len > 0 ? __ldg(wk) : 0u;
        uint32_t w1 = len > 1 ? __ldg(wk + cap) : 0u;
        double Mc[4], Mn[4];
        ld_row(A.tabMT + (size_t) HFG_KEY_ID(w0) * 16 + q * 4, Mc);
        double *ft = A.scrFT + ((size_t) j * 4 + q); /* segment-transposed [k][j][4]: the 8 quads of a warp write 256 contiguous bytes */
#pragma unroll 1
        for (int k = 0; k < kmax; k++) {
            const uint32_t w2 = k + 2 < len ? __ldg(wk + (size_t) (k + 2) * cap) : 0u;
            ld_row(A.tabMT + (size_t) HFG_KEY_ID(w1) * 16 + q * 4, Mn);
            const bool start = (w0 & HFG_KEY_CHUNK_START) != 0;
            double fo;
            if (start) {
                /* EM_fillFirstColumnForward: f[0][s] = e * start probability = any row of the rank-1 matrix */
                fo = Mc[0];
            } else {
                /* f[i][s] = sum_pre f[i-1][pre] * (tProb * eProb), preState ascending (hmm.c:386-408) */
                double a = 0.0;
                a += f[0] * Mc[0];
                a += f[1] * Mc[1];
                a += f[2] * Mc[2];
                a += f[3] * Mc[3];
                fo = a;
            }
            double fn[4];
#pragma unroll
            for (int i = 0; i < 4; i++) fn[i] = __shfl_sync(FULL, fo, qbase + i);
            if (k < len) {
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
                fsum = c * sc;
                st_cg(ft + (size_t) k * cap * 4, fo * sc);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) Mc[c] = Mn[c];
            w0 = w1;
            w1 = w2;
        }
        /* sum_k log c_k of this segment */
        const double loglik = len > 0 ? log(fsum) - (double) esum * 0.6931471805599453 : 0.0;
        if (q == 0) A.seg_loglik[j] = loglik;
    }
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 4] = clock64(); /* thread 0's own C1 end (no barrier here) */

    /* =========================== phase C2: backward + decode ================================================= */
    const bool lab_smem = A.smax <= HFGQ_LAB_SMAX;
    const int cta_w0 = A.seg_start[min(blockIdx.x * QUADS, max(A.n_seg - 1, 0))]; /* first window of this CTA */
    const int lab_shift = cta_w0 & 15; /* shared-memory offset == global offset (mod 16): the copy-out moves aligned vectors */
    if (!A.forward_only) {
        __syncwarp(); /* the forward vectors written by the other lanes of the quad are read below */
        /* f[] holds the forward vector of the segment's last window.  b is a direction, carried unnormalised with exact
         * rescaling: the decode needs it up to a positive factor, the statistics normalise their record themselves
         * (sum_{pre,s} f_{i-1}[pre] M_i[pre][s] b_i[s] = 1, the invariant of the reference's scaling, hmm.c:452-467,613-614). */
        double bh[4] = {1.0, 1.0, 1.0, 1.0};
        bool done = len == 0;
        /* decode of one window: EM_getPosterior / EM_getMostProbableState (hmm.c:671-692), first maximum
         * (common.c:292-303); a common positive factor does not change the order */
        auto decode = [&](const double (&fw)[4], const double (&bw)[4], int gi) {
            double g0 = fw[0] * bw[0], g1 = fw[1] * bw[1], g2 = fw[2] * bw[2], g3 = fw[3] * bw[3];
            if (A.posteriors) {
                const double tot = ((g0 + g1) + g2) + g3;
                g0 /= tot;
                g1 /= tot;
                g2 /= tot;
                g3 /= tot;
                const double gq = q == 0 ? g0 : (q == 1 ? g1 : (q == 2 ? g2 : g3));
                A.posteriors[(size_t) gi * 4 + q] = gq;
            }
            int best = 0;
            double m = g0;
            if (m < g1) { m = g1; best = 1; }
            if (m < g2) { m = g2; best = 2; }
            if (m < g3) { m = g3; best = 3; }
            if (q == 0) {
                if (lab_smem) lab_s[gi - cta_w0 + lab_shift] = (int8_t) best;
                else A.labels[gi] = (int8_t) best;
            }
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
                for (int s = 0; s < 4; s++) bh[s] = stash[4 + s];
            }
            decode(f, bh, seg_first + len - 1); /* the segment's last window; every other window is decoded at the end of
                                                   the step that produces its b */
        }
        double bh_own = sel4(bh, q);
        const uint32_t *wp = A.wposT + j; /* position of the window's record in key-list order, or HFGQ_NOPOS */
        uint32_t w0 = (kmax > 0 && kmax - 1 < len) ? __ldg(wk + (size_t) (kmax - 1) * cap) : 0u;
        uint32_t w1 = (kmax > 1 && kmax - 2 < len) ? __ldg(wk + (size_t) (kmax - 2) * cap) : 0u;
        double Mc[4], Mn[4];
        ld_row(A.tabM + (size_t) HFG_KEY_ID(w0) * 16 + q * 4, Mc);
        const double *ft = A.scrFT + (size_t) j * 4;
#pragma unroll 1
        for (int k = kmax - 1; k >= 0; k--) {
            const uint32_t w2 = (k >= 2 && k - 2 < len) ? __ldg(wk + (size_t) (k - 2) * cap) : 0u;
            ld_row(A.tabM + (size_t) HFG_KEY_ID(w1) * 16 + q * 4, Mn);
            const bool act = k < len && !done;
            const int gi = seg_first + k;
            const uint32_t pos = k < len ? __ldg(wp + (size_t) k * cap) : HFGQ_NOPOS;
            /* forward vector of the previous window (last window of the previous segment == the entering message) */
            double fp[4];
            if (k > 0 && k < len) {
                ld_row_cg(ft + (size_t) (k - 1) * cap * 4, fp);
            } else {
#pragma unroll
                for (int s = 0; s < 4; s++) fp[s] = stash[s];
            }
            /* b[i-1][pre] = sum_s tProb*eProb*b[i][s] (hmm.c:493-520): lane pre forms its component */
            const double bo = ((Mc[0] * bh[0] + Mc[1] * bh[1]) + Mc[2] * bh[2]) + Mc[3] * bh[3];
            double bn[4];
#pragma unroll
            for (int i = 0; i < 4; i++) bn[i] = __shfl_sync(FULL, bo, qbase + i);
            if (act) {
                if (w0 & HFG_KEY_CHUNK_START) {
                    done = true; /* first window of a chunk: nothing to the left */
                } else {
                    if (pos != HFGQ_NOPOS) {
                        /* the window's record for the statistics: (f_{i-1}[q], b_i[q] / (f_{i-1} . M_i b_i)), interleaved so
                         * that the quad writes 64 contiguous bytes */
                        const double dot = ((fp[0] * bn[0] + fp[1] * bn[1]) + fp[2] * bn[2]) + fp[3] * bn[3];
                        st_cg2(A.scrXB + (size_t) pos * 8 + 2 * q, sel4(fp, q), bh_own * (1.0 / dot));
                    }
                    double sc = 1.0;
                    if ((k & 3) == 0) sc = pow2_of(2046 - max_exp4(bn));
#pragma unroll
                    for (int s = 0; s < 4; s++) bh[s] = bn[s] * sc;
                    bh_own = bo * sc;
                    if (k > 0) decode(fp, bh, gi - 1);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) Mc[c] = Mn[c];
            w0 = w1;
            w1 = w2;
        }
    }
    __syncthreads(); /* this CTA's labels are final */
    if (!A.forward_only && lab_smem) {
        /* labels of this CTA's windows (one contiguous range: segments are in genome order): shared memory -> global memory
         * and, for the blocking calls, -> the caller's page-locked buffer over PCIe, as aligned 16-byte vectors (posted
         * writes: they drain while the statistics phase runs) */
        const int s0 = blockIdx.x * QUADS;
        if (s0 < A.n_seg) {
            const long long w_begin = cta_w0;
            const long long w_end = s0 + QUADS < A.n_seg ? (long long) A.seg_start[s0 + QUADS] : (long long) A.n_windows;
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
        const int s0 = blockIdx.x * QUADS;
        if (s0 < A.n_seg) {
            const long long w_begin = cta_w0;
            const long long w_end = s0 + QUADS < A.n_seg ? (long long) A.seg_start[s0 + QUADS] : (long long) A.n_windows;
            const int t = (warp - (WARPS - 2)) * 32 + lane; /* 0..63 */
            for (long long o = w_begin + t; o < w_end; o += 64) A.labels_host[o] = __ldcg(A.labels + o);
        }
    }
    if (uf_flag) atomicOr(A.err_flags, 1);
    grid.sync(); /* the records of every window are in place */
    if (tid == 0) A.phase_clock[blockIdx.x * HFG_PC_STRIDE + 5] = clock64();

    /* =========================== phases S and D, region by region ============================================ */
    /* statistic rows: 0..15 transition counts, 16..17 truncated exponential, 18+3g.. (meanNum, den, varNum) of Gaussian
     * component g, last row the log-likelihood (kept in region 0).  Every CTA takes a contiguous, equal share of the tiles
     * (they are sorted by region): a CTA meets one or two regions, not all R. */
    const long long n_tiles_all = A.forward_only ? 0 : __ldg(&A.region_tile_begin[HFG_MAX_REGIONS]);
    const int blk_t0 = (int) (n_tiles_all * blockIdx.x / gridDim.x), blk_t1 = (int) (n_tiles_all * (blockIdx.x + 1) / gridDim.x);
    for (int r = 0; r < R; r++) {
        const int t_begin = max(blk_t0, __ldg(&A.region_tile_begin[r])), t_end = min(blk_t1, __ldg(&A.region_tile_begin[r + 1]));
        if (r > 0 && t_begin >= t_end) { /* (CTA-uniform) none of this region's tiles here; region 0 carries the log-likelihood */
            for (int qq = tid; qq < NSTAT; qq += THREADS) A.partials[((size_t) r * NSTAT + qq) * gridDim.x + blockIdx.x] = 0.0;
            continue;
        }
        double *ws = wstat + (size_t) warp * NSTAT; /* this warp's row: one writer per entry */
        for (int qq = lane; qq < NSTAT; qq += 32) ws[qq] = 0.0;
        __syncwarp();
        if (r == 0) {
            const double ll = warp_sum(q == 0 ? ld_cg(A.seg_loglik + j) : 0.0); /* written by this lane in C1 */
            if (lane == 0) ws[NSTAT - 1] = ll;
        }
#pragma unroll 1
        for (int t0 = t_begin + warp * 8; t0 < t_end; t0 += QUADS) {
            const int t = t0 + quad;
            const bool act = t < t_end;
            const int p = act ? __ldg(&A.tile_key[t]) : 0, lb = act ? __ldg(&A.tile_begin[t]) : 0, ln = act ? __ldg(&A.tile_cnt[t]) : 0;
            /* The tile's records are consecutive in key-list order.  Lane q takes windows q, q+4, ... of the tile and sums
             * their outer products f (x) b (the four lanes of a quad read 256 contiguous bytes per step); the quad then
             * folds the four sums so that lane pre = q ends with row pre of S[pre][s] = sum_i f_{i-1}[pre] b_i[s]. */
            double S[4];
            {
                double O[16];
#pragma unroll
                for (int i = 0; i < 16; i++) O[i] = 0.0;
                for (int i = q; i < ln; i += 4) {
                    const double *rec = A.scrXB + (size_t) (lb + i) * 8; /* f0 b0 f1 b1 | f2 b2 f3 b3 */
                    double lo[4], hi[4];
                    ld_row_cg(rec, lo);
                    ld_row_cg(rec + 4, hi);
                    const double fw[4] = {lo[0], lo[2], hi[0], hi[2]}, bw[4] = {lo[1], lo[3], hi[1], hi[3]};
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) O[a * 4 + b] = fma(fw[a], bw[b], O[a * 4 + b]);
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
        __syncthreads();
        /* deterministic CTA reduction: the warps' rows in warp order */
        for (int qq = tid; qq < NSTAT; qq += THREADS) {
            double sum = 0.0;
#pragma unroll 4
            for (int wv = 0; wv < WARPS; wv++) sum += wstat[(size_t) wv * NSTAT + qq];
            A.partials[((size_t) r * NSTAT + qq) * gridDim.x + blockIdx.x] = sum; /* [R][NSTAT][grid] */
        }
        __syncthreads();
    }
    if (nan_flag) atomicOr(A.err_flags, 2);

    hfg_estep_tail<THREADS>(A, wstat, ks, &s_last);
}
