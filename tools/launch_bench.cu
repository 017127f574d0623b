// What one launch costs on the stream, event to event, for a kernel shaped like the E-step kernel (148 x 512 threads, 226 KB of
// dynamic shared memory, 1.5 KB of parameters): plain launch against cooperative launch, alone and back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_bench tools/launch_bench.cu && ./launch_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
struct Args { long long pad[180]; int *out; };
__global__ void __launch_bounds__(512, 1) k(const Args a) {
    extern __shared__ double sm[];
    if (threadIdx.x == 0 && a.pad[3] == 12345) a.out[blockIdx.x] = (int) sm[0];
}
struct Small { long long pad[5]; int *out; };
__global__ void __launch_bounds__(512, 1) ks(const Small a) {
    extern __shared__ double sm[];
    if (threadIdx.x == 0 && a.pad[3] == 12345) a.out[blockIdx.x] = (int) sm[0];
}
__global__ void __launch_bounds__(512, 1) kc(const Args a) {
    extern __shared__ double sm[];
    cooperative_groups::grid_group g = cooperative_groups::this_grid();
    if (a.pad[4] == 777) g.sync();
    if (threadIdx.x == 0 && a.pad[3] == 12345) a.out[blockIdx.x] = (int) sm[0];
}
int main() {
    int *out; cudaMalloc(&out, 4096);
    Args a = {}; a.out = out;
    const int smem = 226 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void *args[] = {&a};
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            float alone = 0, chain = 0;
            for (int i = 0; i < 20; i++) { // alone: event pair around one launch, stream idle before
                cudaStreamSynchronize(s);
                cudaEventRecord(e0, s);
                if (mode == 0) k<<<148, 512, smem, s>>>(a); else cudaLaunchCooperativeKernel((void *) kc, dim3(148), dim3(512), args, smem, s);
                cudaEventRecord(e1, s);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); alone += ms;
            }
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            for (int i = 0; i < 200; i++) {
                if (mode == 0) k<<<148, 512, smem, s>>>(a); else cudaLaunchCooperativeKernel((void *) kc, dim3(148), dim3(512), args, smem, s);
            }
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&chain, e0, e1);
            printf("%s launch: alone %.2f us event to event, back to back %.2f us per launch\n", mode ? "cooperative" : "plain      ", 1e3 * alone / 20, 1e3 * chain / 200);
        }
    }
    // variants: shared memory size, parameter size, grid size (plain launches, alone)
    cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Small sa = {}; sa.out = out;
    struct { const char *name; int big_args, smem, grid, threads; } v[] = {
        {"1.5 KB args, 226 KB smem, 148 x 512", 1, smem, 148, 512}, {"1.5 KB args, 0 smem, 148 x 512", 1, 0, 148, 512},
        {"48 B args, 226 KB smem, 148 x 512", 0, smem, 148, 512},  {"48 B args, 0 smem, 148 x 512", 0, 0, 148, 512},
        {"48 B args, 0 smem, 1 x 32", 0, 0, 1, 32},                  {"48 B args, 226 KB smem, 1 x 32", 0, smem, 1, 32},
        {"1.5 KB args, 226 KB smem, 18 x 512", 1, smem, 18, 512}};
    for (auto &c : v) {
        float alone = 0;
        for (int i = 0; i < 30; i++) {
            cudaStreamSynchronize(s);
            cudaEventRecord(e0, s);
            if (c.big_args) k<<<c.grid, c.threads, c.smem, s>>>(a); else ks<<<c.grid, c.threads, c.smem, s>>>(sa);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 10) alone += ms;
        }
        // after a memset on the stream (the L2 flush of the bench): does the preceding kernel's shared-memory configuration matter?
        float after = 0;
        for (int i = 0; i < 30; i++) {
            cudaMemsetAsync(out, 0, 4096, s);
            cudaEventRecord(e0, s);
            if (c.big_args) k<<<c.grid, c.threads, c.smem, s>>>(a); else ks<<<c.grid, c.threads, c.smem, s>>>(sa);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 10) after += ms;
        }
        printf("%-40s alone %.2f us, right behind a memset %.2f us\n", c.name, 1e3 * alone / 20, 1e3 * after / 20);
    }
    // everything queued behind a LONG memset (the bench's 256 MiB L2 flush): the host's launch latency is hidden, what remains
    // is the device's own event -> kernel -> event cost; then the same sequence as a captured graph
    {
        void *big; cudaMalloc(&big, 256 << 20);
        float q = 0;
        for (int i = 0; i < 30; i++) {
            cudaMemsetAsync(big, 0x5a, 256 << 20, s);
            cudaEventRecord(e0, s);
            cudaLaunchCooperativeKernel((void *) kc, dim3(148), dim3(512), args, smem, s);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 10) q += ms;
        }
        printf("cooperative, queued behind a 256 MiB memset: %.2f us event to event\n", 1e3 * q / 20);
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        cudaMemsetAsync(big, 0x5a, 256 << 20, s);
        cudaEventRecordWithFlags(e0, s, cudaEventRecordExternal);
        cudaLaunchCooperativeKernel((void *) kc, dim3(148), dim3(512), args, smem, s);
        cudaEventRecordWithFlags(e1, s, cudaEventRecordExternal);
        cudaStreamEndCapture(s, &g);
        cudaGraphInstantiate(&ge, g, 0);
        q = 0;
        for (int i = 0; i < 30; i++) {
            cudaGraphLaunch(ge, s);
            cudaEventSynchronize(e1);
            cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 10) q += ms;
        }
        printf("the same as a graph (memset, event, kernel, event): %.2f us event to event\n", 1e3 * q / 20);
        cudaFree(big);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
