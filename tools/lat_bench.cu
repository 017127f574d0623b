// Dependent-issue latency of the double-precision pipe and friends on one warp (B200, sm_100a): cycles per dependent instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench tools/lat_bench.cu && ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int MODE>
__global__ void k(double *out, long long *cyc, double a, double b, int warps) {
    double x = a + threadIdx.x, y = b, z = a * 3, w = b * 5;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (MODE == 0) x = fma(x, y, y);                                   // 1 chain
        if (MODE == 1) { x = fma(x, y, y); z = fma(z, y, y); }             // 2 chains
        if (MODE == 2) { x = fma(x, y, y); z = fma(z, y, y); w = fma(w, y, y); a = fma(a, y, y); } // 4 chains
        if (MODE == 3) x = x + y;                                          // DADD
        if (MODE == 4) x = x * y;                                          // DMUL
        if (MODE == 5) x = 1.0 / x;                                        // division (reciprocal)
        if (MODE == 6) x = exp(x) * 1e-300;                                // exp
        if (MODE == 7) x = log(x + 2.0);                                   // log
        if (MODE == 8) { float f = (float) x; f = fmaf(f, 1.0001f, 0.5f); x = f; } // conversions + FFMA
        if (MODE == 9) x = x > y ? z : w, z = z + 1.0;                     // compare + select
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + z + w + a;
}
template <int MODE> void run(const char *name, int per, double *out, long long *cyc) {
    for (int warps = 1; warps <= 16; warps *= 4) {
        k<MODE><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, warps);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-44s warps/CTA %2d: %7.2f cycles per loop body (%d instruction(s) per chain step)\n", name, warps, (double) h / N, per);
    }
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    run<0>("DFMA, 1 dependent chain", 1, out, cyc);
    run<1>("DFMA, 2 independent chains", 1, out, cyc);
    run<2>("DFMA, 4 independent chains", 1, out, cyc);
    run<3>("DADD dependent", 1, out, cyc);
    run<4>("DMUL dependent", 1, out, cyc);
    run<5>("1.0 / x dependent", 1, out, cyc);
    run<6>("exp dependent", 1, out, cyc);
    run<7>("log dependent", 1, out, cyc);
    run<8>("D2F + FFMA + F2D dependent", 3, out, cyc);
    run<9>("DSETP + select + DADD", 2, out, cyc);
    return 0;
}
