// Stand-alone timing of the pieces of the device rate fit (one warp, cold then warm): nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __noinline__ double hfg_objective_dev(double rate, double trunc, double sum_x, double sum_w) {
    return sum_w * log(rate) - sum_w * log(1.0 - exp(-rate * trunc)) - sum_x * rate;
}
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
    /* who holds the value at x1 / x2: 0, 1 the round's y1, y2; 2 + t the point of step t */
    int i1 = 0, i2 = 1;
#pragma unroll 1
    for (int t = 0; t < L; t++) {
        /* both points on one side of the maximum: the nearer one is higher; x* between them: the polynomial decides */
        int b;
        if (sx2 <= xs) b = 0;
        else if (sx1 >= xs) b = 1;
        else {
            const double d1 = sx1 - xs, d2 = sx2 - xs;
            const double p1 = d1 * fma(d1, fma(d1, fma(d1, c4, c3), c2), c1), p2 = d2 * fma(d2, fma(d2, fma(d2, c4, c3), c2), c1);
            b = p1 > p2;
        }
        int word = i1 | (i2 << 6) | (b << 12); /* the operands and the predicted outcome of step t's comparison */
        sspan = inv_phi * sspan;
        double xn;
        if (b) {
            shi = sx2; sx2 = sx1; i2 = i1;
            sx1 = slo + inv_phi2 * sspan;
            xn = sx1;
            i1 = 2 + t;
        } else {
            slo = sx1; sx1 = sx2; i1 = i2;
            sx2 = slo + inv_phi * sspan;
            xn = sx2;
            i2 = 2 + t;
        }
        word |= (i1 << 13) | (i2 << 19); /* the holders after step t */
        if (lane == 0) {
            const unsigned row = snap_s + 64u * (unsigned) t;
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(row), "d"(xn), "d"(slo) : "memory");
            asm volatile("st.shared.v2.f64 [%0+16], {%1, %2};" ::"r"(row), "d"(sx1), "d"(sx2) : "memory");
            asm volatile("st.shared.v2.f64 [%0+32], {%1, %2};" ::"r"(row), "d"(shi), "d"(sspan) : "memory");
            asm volatile("st.shared.u32 [%0+48], %1;" ::"r"(row), "r"(word) : "memory");
        }
    }
    __syncwarp();
    if (clk && lane == 0) clk[14] = clock64();
    double4 ra;
    double2 rb;
    int word;
    {
        const unsigned row = snap_s + 64u * (unsigned) (lane < L ? lane : 0);
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ra.x), "=d"(ra.y) : "r"(row));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(ra.z), "=d"(ra.w) : "r"(row));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(rb.x), "=d"(rb.y) : "r"(row));
        asm volatile("ld.shared.u32 %0, [%1+48];" : "=r"(word) : "r"(row));
    }
    const int ca = word & 63, cb = (word >> 6) & 63, pb = (word >> 12) & 1, a1 = (word >> 13) & 63, a2 = (word >> 19) & 63;
    const double ynew = lane < L ? hfg_objective_dev(ra.x, trunc, sum_x, sum_w) : 0.0;
    if (clk && lane == 0) clk[13] = clock64();
    /* the evaluated values behind every comparison, then the first step whose prediction was wrong */
    const double ya_s = __shfl_sync(FULLW, ynew, ca < 2 ? 0 : ca - 2), yb_s = __shfl_sync(FULLW, ynew, cb < 2 ? 0 : cb - 2);
    const double ya = ca == 0 ? y1 : ca == 1 ? y2 : ya_s, yb = cb == 0 ? y1 : cb == 1 ? y2 : yb_s;
    const unsigned wrong = __ballot_sync(FULLW, lane < L && (int) (ya > yb) != pb);
    const int m = wrong ? __ffs(wrong) - 1 : L;
    const double n1_s = __shfl_sync(FULLW, ynew, a1 < 2 ? 0 : a1 - 2), n2_s = __shfl_sync(FULLW, ynew, a2 < 2 ? 0 : a2 - 2);
    const double n1 = a1 == 0 ? y1 : a1 == 1 ? y2 : n1_s, n2 = a2 == 0 ? y1 : a2 == 1 ? y2 : n2_s;
    if (m > 0) {
        st->lo = __shfl_sync(FULLW, ra.y, m - 1);
        st->x1 = __shfl_sync(FULLW, ra.z, m - 1);
        st->x2 = __shfl_sync(FULLW, ra.w, m - 1);
        st->hi = __shfl_sync(FULLW, rb.x, m - 1);
        st->span = __shfl_sync(FULLW, rb.y, m - 1);
        st->y1 = __shfl_sync(FULLW, n1, m - 1);
        st->y2 = __shfl_sync(FULLW, n2, m - 1);
    }
    __syncwarp();
    return m;
}


__global__ void k(long long *out, double *sink, int L) {
    __shared__ __align__(16) double snap[256];
    const int lane = threadIdx.x & 31;
    for (int rep = 0; rep < 4; rep++) {
        double co[5];
        const long long t0 = clock64();
        hfg_fit_newton_dev(1.7, 0.0, 7.25, 7.25, 1000.0, 2000.0, co);
        const long long t1 = clock64();
        HfgFitState st = {0.0, 7.25 * 0.381966, 7.25 * 0.618034, 7.25, 7.25, 0.0, 0.0};
        st.y1 = hfg_objective_dev(st.x1, 7.25, 1000.0, 2000.0);
        st.y2 = hfg_objective_dev(st.x2, 7.25, 1000.0, 2000.0);
        const long long t2 = clock64();
        const int m = hfg_fit_pred_round_dev(&st, co, 7.25, 1000.0, 2000.0, lane, L, snap, out + 32);
        const long long t3 = clock64();
        if (threadIdx.x == 0) { out[rep * 4] = t1 - t0; out[rep * 4 + 1] = t2 - t1; out[rep * 4 + 2] = t3 - t2; out[rep * 4 + 3] = m; sink[rep] = st.lo + co[0]; }
    }
}
int main() {
    long long *out, h[64]; double *sink;
    cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 64);
    for (int warps = 1; warps <= 8; warps *= 8) {
        k<<<1, 32 * warps>>>(out, sink, 32);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        for (int r = 0; r < 4; r++) printf("warps %d rep %d: newton %lld, two objective evaluations %lld, predicted round (32 steps) %lld cycles, m = %lld; walk %lld\n", warps, r, h[r*4], h[r*4+1], h[r*4+2], h[r*4+3], 0LL);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
