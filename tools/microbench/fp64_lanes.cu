// fp64_lanes.cu -- microbenchmark: what does a partially filled warp cost on the FP64 pipe of sm_100a?
// Register-resident DFMA chains (8 independent chains per thread) with only some lanes of each warp active.
// If a warp instruction whose upper (or lower) 16 lanes are all inactive takes one pipe pass instead of two, a
// kernel that packs its work into half-warps loses nothing to granularity.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(double* out, int iters, unsigned lane_mask, double seed) {
    const int lane = threadIdx.x & 31;
    if (!((lane_mask >> lane) & 1u)) return;
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000, blocks = sms * 8, threads = 256;
    struct { const char* name; unsigned mask; } cfg[] = {
        {"all 32 lanes", 0xffffffffu}, {"lower 16", 0x0000ffffu}, {"upper 16", 0xffff0000u}, {"even lanes (16)", 0x55555555u},
        {"lower 8", 0x000000ffu}, {"lanes 0-23", 0x00ffffffu}, {"1 lane", 0x1u}};
    for (auto& c : cfg) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        chain<<<blocks, threads>>>(out, 100, c.mask, 1.0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        chain<<<blocks, threads>>>(out, iters, c.mask, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warp_inst = (double)blocks * (threads / 32) * iters * 64.0;
        printf("%-18s %8.3f ms   %.3f warp-DFMA/clk/SM at 1.965 GHz\n", c.name, ms, warp_inst / (ms * 1e-3) / 1.965e9 / sms);
    }
    // latency: one warp per SM sub-partition, a single dependent chain
    return 0;
}
