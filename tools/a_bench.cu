// Replica of the E-step kernel's phase A inner loop (segment product P <- P * M' with lane-rotated shared-memory reads),
// to see what bounds it: tools/a_bench (B200).  Variants switch off the loads, the permutation, the arithmetic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define STEPS 64
__device__ __forceinline__ double selp(bool c, double a, double b) {
    double r;
    asm("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n selp.f64 %0, %1, %2, p;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int) c));
    return r;
}
__device__ __forceinline__ void xorperm4(double (&x)[4], bool d0, bool d1) {
    const double t0 = selp(d0, x[1], x[0]), t1 = selp(d0, x[0], x[1]), t2 = selp(d0, x[3], x[2]), t3 = selp(d0, x[2], x[3]);
    x[0] = selp(d1, t2, t0); x[1] = selp(d1, t3, t1); x[2] = selp(d1, t0, t2); x[3] = selp(d1, t1, t3);
}
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) bench(const double *gtab, int n_rows, long long *cyc, double *sink) {
    extern __shared__ double tab[];
    const int tid = threadIdx.x, lane = tid & 31, li = lane & 7;
    for (int i = tid; i < n_rows * 16; i += THREADS) tab[i] = gtab[i];
    __syncthreads();
    const bool d0 = (li & 2) != 0, d1 = ((li & 4) != 0) != ((li & 1) != 0);
    uint32_t s = 12345u + 977u * tid + blockIdx.x;
    double P[16];
#pragma unroll
    for (int i = 0; i < 16; i++) P[i] = (i % 5 == 0) ? 1.0 : 0.0;
    double M[16];
#pragma unroll
    for (int i = 0; i < 16; i++) M[i] = 0.25 + 1e-3 * i;
    const long long t0 = clock64();
#pragma unroll 1
    for (int k = 0; k < STEPS; k++) {
        s = s * 1664525u + 1013904223u;
        const int row = (s >> 8) % n_rows;
        if (MODE != 2) {
            const double2 *p = reinterpret_cast<const double2 *>(tab + (size_t) row * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double2 v = p[c ^ li];
                M[2 * c] = v.x;
                M[2 * c + 1] = v.y;
            }
        }
        if (MODE == 3) { // loads only
            P[0] += M[0] + M[3] + M[5] + M[7] + M[9] + M[11] + M[13] + M[15];
            continue;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double t[4];
#pragma unroll
            for (int c = 0; c < 4; c++) t[c] = fma(P[r * 4 + 3], M[12 + c], fma(P[r * 4 + 2], M[8 + c], fma(P[r * 4 + 1], M[4 + c], P[r * 4] * M[c])));
            if (MODE != 1) xorperm4(t, d0, d1);
#pragma unroll
            for (int c = 0; c < 4; c++) P[r * 4 + c] = t[c];
        }
        if ((k & 3) == 3) {
            int hi = __double2hiint(P[0]);
#pragma unroll
            for (int i = 1; i < 16; i++) hi = max(hi, __double2hiint(P[i]));
            const double sc = __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
#pragma unroll
            for (int i = 0; i < 16; i++) P[i] *= sc;
        }
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    double a = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) a += P[i];
    sink[blockIdx.x * THREADS + tid] = a;
}
template <int MODE, int THREADS>
static void run(const char *name, const double *gtab, int n_rows, long long *d_cyc, double *d_sink) {
    const size_t smem = (size_t) n_rows * 128;
    cudaFuncSetAttribute(bench<MODE, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    for (int rep = 0; rep < 2; rep++) bench<MODE, THREADS><<<148, THREADS, smem>>>(gtab, n_rows, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148.0;
    printf("%-44s threads %4d: %7.0f cycles per step-round (all warps one step) = %6.1f per warp-step per SM  %s\n", name, THREADS,
           avg / STEPS, avg / STEPS / (THREADS / 32), e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    const int n_rows = 1300;
    double *gtab, *d_sink;
    long long *d_cyc;
    cudaMalloc(&gtab, (size_t) n_rows * 128);
    cudaMemset(gtab, 0, (size_t) n_rows * 128);
    cudaMalloc(&d_sink, 148 * 1024 * 8);
    cudaMalloc(&d_cyc, 148 * 8);
    run<0, 512>("full (loads + 64 DFMA + permutation)", gtab, n_rows, d_cyc, d_sink);
    run<1, 512>("no permutation", gtab, n_rows, d_cyc, d_sink);
    run<2, 512>("no loads", gtab, n_rows, d_cyc, d_sink);
    run<3, 512>("loads only", gtab, n_rows, d_cyc, d_sink);
    run<0, 256>("full", gtab, n_rows, d_cyc, d_sink);
    run<2, 256>("no loads", gtab, n_rows, d_cyc, d_sink);
    run<0, 640>("full", gtab, n_rows, d_cyc, d_sink);
    run<2, 128>("no loads", gtab, n_rows, d_cyc, d_sink);
    return 0;
}
