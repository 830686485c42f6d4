// cuda_emul.h -- TEST INFRASTRUCTURE.  A minimal host emulation of the CUDA execution model so
// that the *exact* kernel source (celeste.jl_b200/csrc/celeste_kernels.cuh) can be run on a
// machine without a GPU and compared with the oracle: one std::thread per CUDA thread of a
// block, std::barrier for __syncthreads/__syncwarp, exchange buffers for shuffles/ballots;
// blocks run one after another.  Never linked into the product library.
#ifndef CUDA_EMUL_H
#define CUDA_EMUL_H
#define CELESTE_HOST_EMULATION 1

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)

struct uint3_emul { unsigned x, y, z; };
struct int2 { int x, y; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

namespace cuda_emul {
struct BlockCtx {
    int nthreads;
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<double> xchg_d;     // per thread
    std::vector<int> xchg_i;
    std::vector<unsigned char> dyn;
};
inline BlockCtx*& ctx() { static BlockCtx* c = nullptr; return c; }
inline double* dynamic_smem() { return reinterpret_cast<double*>(ctx()->dyn.data()); }
}  // namespace cuda_emul

inline thread_local uint3_emul threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { cuda_emul::ctx()->block_bar->arrive_and_wait(); }
inline void __syncwarp() { cuda_emul::ctx()->warp_bar[threadIdx.x / 32]->arrive_and_wait(); }
inline unsigned __ballot_sync(unsigned, bool pred) {
    auto* c = cuda_emul::ctx();
    const int w = threadIdx.x / 32;
    c->xchg_i[threadIdx.x] = pred ? 1 : 0;
    c->warp_bar[w]->arrive_and_wait();
    unsigned r = 0;
    for (int l = 0; l < 32 && w * 32 + l < c->nthreads; ++l) r |= (unsigned)c->xchg_i[w * 32 + l] << l;
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
inline double __shfl_xor_sync(unsigned, double v, int o) {
    auto* c = cuda_emul::ctx();
    const int w = threadIdx.x / 32;
    c->xchg_d[threadIdx.x] = v;
    c->warp_bar[w]->arrive_and_wait();
    const double r = c->xchg_d[(threadIdx.x ^ o)];
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
inline int __shfl_sync(unsigned, int v, int src) {
    auto* c = cuda_emul::ctx();
    const int w = threadIdx.x / 32;
    c->xchg_i[threadIdx.x] = v;
    c->warp_bar[w]->arrive_and_wait();
    const int r = c->xchg_i[w * 32 + src];
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
inline double __shfl_sync(unsigned, double v, int src) {
    auto* c = cuda_emul::ctx();
    const int w = threadIdx.x / 32;
    c->xchg_d[threadIdx.x] = v;
    c->warp_bar[w]->arrive_and_wait();
    const double r = c->xchg_d[w * 32 + src];
    c->warp_bar[w]->arrive_and_wait();
    return r;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == (threadIdx.x / 32 * 32 + 32 <= (unsigned)cuda_emul::ctx()->nthreads ? 0xffffffffu : ((1u << (cuda_emul::ctx()->nthreads % 32)) - 1u)); }
inline double __ldg(const double* p) { return *p; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
using std::isfinite;
using std::isnan;
using std::max;
using std::min;

namespace cuda_emul {
template <typename F, typename... Args>
void launch(F kernel, int grid, int block, size_t smem_bytes, Args... args) {
    for (int b = 0; b < grid; ++b) {
        BlockCtx c;
        c.nthreads = block;
        c.block_bar = std::make_unique<std::barrier<>>(block);
        for (int w = 0; w < (block + 31) / 32; ++w)
            c.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32, block - 32 * w)));
        c.xchg_d.assign(block, 0.0);
        c.xchg_i.assign(block, 0);
        c.dyn.assign(smem_bytes + 16, 0);
        ctx() = &c;
        std::vector<std::thread> th;
        for (int t = 0; t < block; ++t)
            th.emplace_back([=]() {
                threadIdx = {(unsigned)t, 0, 0};
                blockIdx = {(unsigned)b, 0, 0};
                blockDim = {(unsigned)block, 1, 1};
                gridDim = {(unsigned)grid, 1, 1};
                kernel(args...);
            });
        for (auto& x : th) x.join();
        ctx() = nullptr;
    }
}
}  // namespace cuda_emul
#endif
