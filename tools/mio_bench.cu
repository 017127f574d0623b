// Microbenchmark of the SM's memory-pipe instruction costs that decide the E-step kernel's layout (B200, sm_100a):
// cycles per warp-instruction per SM for SHFL, LDS (per-lane random rows, quad broadcast), LDG hits with 8 / 32 lines.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mio_bench tools/mio_bench.cu && ./mio_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(const double *__restrict__ gtab, int n_rows, long long *cycles, double *sink) {
    extern __shared__ double tab[]; // [n_rows][16] (+ swizzle by mode)
    const int tid = threadIdx.x, lane = tid & 31, q = tid & 3;
    for (int i = tid; i < n_rows * 16; i += blockDim.x) tab[i] = gtab[i];
    __syncthreads();
    uint32_t s = 12345u + 977u * (MODE >= 20 ? (tid >> 2) : tid) + blockIdx.x; // quad-uniform stream for the quad modes
    double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        const int row = rng(s) % n_rows;
        if (MODE == 0) { // 8 SHFL.32 (4 doubles from the lanes of the quad)
            acc0 += __shfl_sync(0xffffffffu, acc1 + row, (lane & ~3) + 0);
            acc1 += __shfl_sync(0xffffffffu, acc2 + row, (lane & ~3) + 1);
            acc2 += __shfl_sync(0xffffffffu, acc3 + row, (lane & ~3) + 2);
            acc3 += __shfl_sync(0xffffffffu, acc0 + row, (lane & ~3) + 3);
        } else if (MODE == 1) { // per-lane random row, whole matrix: 8 LDS.128, row stride 128 B, XOR swizzle of the 16-byte chunk
            const double2 *p = reinterpret_cast<const double2 *>(tab + row * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) { const double2 v = p[c ^ (row & 7)]; acc0 += v.x; acc1 += v.y; }
        } else if (MODE == 2) { // per-lane random row, conflict-free by rotation (lane-dependent chunk order)
            const double2 *p = reinterpret_cast<const double2 *>(tab + row * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) { const double2 v = p[(c ^ lane) & 7]; acc0 += v.x; acc1 += v.y; }
        } else if (MODE == 3) { // per-lane random row, no swizzle (8-way conflicts expected)
            const double2 *p = reinterpret_cast<const double2 *>(tab + row * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) { const double2 v = p[c]; acc0 += v.x; acc1 += v.y; }
        } else if (MODE == 20) { // quad broadcast: the 4 lanes of a quad read the whole matrix of the quad's row (8 LDS.128), swizzled
            const double2 *p = reinterpret_cast<const double2 *>(tab + row * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) { const double2 v = p[c ^ (row & 7)]; acc0 += v.x; acc1 += v.y; }
        } else if (MODE == 21) { // quad: lane q reads row q of the quad's matrix (2 LDS.128), swizzled
            const double2 *p = reinterpret_cast<const double2 *>(tab + row * 16);
            const double2 v = p[(2 * q) ^ (row & 7)], w = p[(2 * q + 1) ^ (row & 7)];
            acc0 += v.x; acc1 += v.y; acc2 += w.x; acc3 += w.y;
        } else if (MODE == 22) { // quad: lane q reads column q (4 LDS.64)
            const double *p = tab + row * 16;
#pragma unroll
            for (int r = 0; r < 4; r++) acc0 += p[(((2 * r + (q >> 1)) ^ (row & 7)) << 1) + (q & 1)];
        } else if (MODE == 23) { // quad: lane q reads row q from GLOBAL (L1-resident table): one LDG.256, 8 lines
            const double *p = gtab + row * 16 + q * 4;
            double a, b, c, d;
            asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
            acc0 += a; acc1 += b; acc2 += c; acc3 += d;
        } else if (MODE == 24) { // quad: exchange of one double per lane through shared memory (STS.64 + 2 LDS.128) instead of 8 SHFL
            double *x = tab + n_rows * 16 + (tid & ~3);
            x[q] = acc0 + row;
            __syncwarp();
            const double2 v = reinterpret_cast<const double2 *>(x)[0], w = reinterpret_cast<const double2 *>(x)[1];
            __syncwarp();
            acc0 += v.x; acc1 += v.y; acc2 += w.x; acc3 += w.y;
        } else if (MODE == 4) { // per-lane random row from GLOBAL (L1-resident): 4 LDG.256, 32 lines each
            const double *p = gtab + row * 16;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                double a, b, c, d;
                asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + 4 * r));
                acc0 += a; acc1 += b; acc2 += c; acc3 += d;
            }
        } else if (MODE == 5) { // coalesced store + load of 8 bytes per lane (scratch traffic)
            double *x = sink + 4096 + (size_t) blockIdx.x * 65536 + ((it & 31) * 1024 + tid);
            asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(x), "d"(acc0) : "memory");
        } else if (MODE == 6) { // nothing but the generator (baseline to subtract)
            acc0 += row;
        }
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 1024 + tid] = acc0 + acc1 + acc2 + acc3;
}

template <int MODE>
static void run(const char *name, const double *gtab, int n_rows, long long *d_cyc, double *d_sink, double base) {
    const size_t smem = (size_t) n_rows * 128 + 1024 * 8;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    bench<MODE><<<148, 1024, smem>>>(gtab, n_rows, d_cyc, d_sink);
    bench<MODE><<<148, 1024, smem>>>(gtab, n_rows, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += h[i];
    avg /= 148.0;
    // 32 warps per SM execute the loop body; cycles per warp-level loop body per SM:
    const double per = avg / ITERS / 32.0;
    printf("%-72s %8.2f cycles per warp-iteration per SM (net of generator %6.2f)  %s\n", name, per, per - base,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int n_rows = 1400;
    double *gtab, *d_sink;
    long long *d_cyc;
    cudaMalloc(&gtab, (size_t) n_rows * 128);
    cudaMemset(gtab, 0, (size_t) n_rows * 128);
    cudaMalloc(&d_sink, (size_t) (4096 + 148 * 65536) * 8 + 148 * 1024 * 8);
    cudaMalloc(&d_cyc, 148 * 8);
    long long h[148];
    {
        const size_t smem = (size_t) n_rows * 128 + 1024 * 8;
        cudaFuncSetAttribute(bench<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        bench<6><<<148, 1024, smem>>>(gtab, n_rows, d_cyc, d_sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    }
    double base = 0;
    for (int i = 0; i < 148; i++) base += h[i];
    base = base / 148.0 / ITERS / 32.0;
    printf("generator only: %.2f cycles per warp-iteration per SM\n", base);
    run<0>("8 SHFL.32 (all-gather of one double in a quad)", gtab, n_rows, d_cyc, d_sink, base);
    run<1>("thread: 8 LDS.128, random row per lane, XOR-swizzled chunks", gtab, n_rows, d_cyc, d_sink, base);
    run<2>("thread: 8 LDS.128, random row per lane, lane-rotated chunks (conflict-free)", gtab, n_rows, d_cyc, d_sink, base);
    run<3>("thread: 8 LDS.128, random row per lane, no swizzle", gtab, n_rows, d_cyc, d_sink, base);
    run<4>("thread: 4 LDG.256, random row per lane (L1 hits, 32 lines)", gtab, 512, d_cyc, d_sink, base);
    run<20>("quad: 8 LDS.128 broadcast, whole matrix, random row per quad", gtab, n_rows, d_cyc, d_sink, base);
    run<21>("quad: 2 LDS.128, lane q reads row q", gtab, n_rows, d_cyc, d_sink, base);
    run<22>("quad: 4 LDS.64, lane q reads column q", gtab, n_rows, d_cyc, d_sink, base);
    run<23>("quad: 1 LDG.256, lane q reads row q (L1 hits, 8 lines)", gtab, 512, d_cyc, d_sink, base);
    run<24>("quad: STS.64 + 2 LDS.128 (exchange through shared memory)", gtab, n_rows, d_cyc, d_sink, base);
    run<5>("coalesced STG.64 per lane (256 B per warp)", gtab, n_rows, d_cyc, d_sink, base);
    return 0;
}
