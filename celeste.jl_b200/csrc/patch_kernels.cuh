// patch_kernels.cuh -- ImagePatch construction and neighbour discovery on the device.
//
//   ImagePatch(img, box)  (src/model/imaged_sources.jl:80-117), the parts that touch pixels or the PSF stamp:
//     active_pixel_bitmap = [!isnan(img.pixels[x, y]) for x in box[1], y in box[2]]              (:92-95)
//     grid_psf = max.(psfmap(center), 0) + 1e-6, normalised, softpluslike                       (:97-105)
//     itp_psf  = interpolate(grid_psf, BSpline(Cubic(Line())), OnGrid())                        (:107)
//   with psfmap either handed over as a raw stamp or rasterised here from the image's Gaussian-mixture PSF
//   (render_psf, src/model/psf_model.jl:61-75).  The box arithmetic and the WCS linearisation stay on the host
//   (WCS.jl is a host library); they are a handful of flops per patch.
//   find_neighbors(patches, target)  (imaged_sources.jl:232-244): sources whose box overlaps the target's in any image.
//
// Interpolations.jl is an un-vendored dependency (REQUIRE:21): the prefilter is restated (same statement as
// model.cubic_bspline_prefilter, which the tests compare this kernel against): coefficients padded by one per
// side, interior rows (1/6, 2/3, 1/6), "Line" boundary rows (1, -2, 1) -- which reduce to c_1 = a_1, c_n = a_n, a
// constant-coefficient tridiagonal system for c_2..c_{n-1} (Thomas algorithm), c_0 = 2 c_1 - c_2 and
// c_{n+1} = 2 c_n - c_{n-1}; separable, first axis then second.  PARITY UNPINNED like the host version.
#ifndef CELESTE_PATCH_KERNELS_CUH
#define CELESTE_PATCH_KERNELS_CUH

#include "celeste_kernels.cuh"

namespace celeste {

constexpr int PB_THREADS = 64;
constexpr int PB_MAX_GRID = 54;      // stamp side limit: n x n + (n + 2) x n doubles must fit 48 KB of shared memory

struct SplineJob {
    const double* raw;     // grid_n x grid_n raw psfmap stamp (device, column-major), or null: rasterise `psf`
    const double* psf;     // K x 7 (alphaBar, xiBar[2], tauBar[4]) when raw == null
    int K;
    int grid_n;
    double* coefs;         // out: (grid_n + 2)^2, column-major
};

// one tridiagonal solve of the prefilter along a strided line: a[0..n-1] (stride sa) -> c[0..n+1] (stride sc);
// cpv[i] = the data-independent Thomas factors.  `dp` is caller scratch of n doubles (stride 1, per thread).
__device__ inline void prefilter_line(const double* a, int sa, double* c, int sc, int n, const double* cpv) {
    // unknowns u_i = c_{i+1}, i = 0..n-1 (grid index i+1): u_0 = a_0, u_{n-1} = a_{n-1}
    const double first = a[0], last = a[(n - 1) * sa];
    c[1 * sc] = first;
    c[n * sc] = last;
    if (n > 2) {
        // forward sweep over i = 1..n-2, storing dp in the output line itself
        double prev = 0.0;
        for (int i = 1; i <= n - 2; ++i) {
            double rhs = 6.0 * a[i * sa];
            if (i == 1) rhs -= first;
            if (i == n - 2) rhs -= last;
            const double denom_inv = cpv[i];                     // 1 / (4 - cp[i-1]),  cp[0] = 0
            prev = (rhs - (i > 1 ? prev : 0.0)) * denom_inv;
            c[(i + 1) * sc] = prev;
        }
        for (int i = n - 3; i >= 1; --i) c[(i + 1) * sc] -= cpv[i] * c[(i + 2) * sc];
    }
    c[0] = 2.0 * c[1 * sc] - c[2 * sc];
    c[(n + 1) * sc] = 2.0 * c[n * sc] - c[(n - 1) * sc];
}

__global__ void __launch_bounds__(PB_THREADS) spline_build_kernel(const SplineJob* __restrict__ jobs) {
    CEL_DYNAMIC_SMEM(smem);
    __shared__ double red[PB_THREADS];
    __shared__ double cpv[PB_MAX_GRID + 2];
    __shared__ double s_total;
    const SplineJob job = jobs[blockIdx.x];
    const int n = job.grid_n, n2 = n + 2, tid = threadIdx.x;
    double* g = smem;              // n x n, column-major
    double* T = smem + n * n;      // (n + 2) x n: prefiltered along the first axis

    // 1. the stamp: raw or render_psf (psf_model.jl:61-75: centred at (dims + 1) / 2), then max(., 0) + 1e-6
    const double c0 = 0.5 * (n + 1);
    double part = 0.0;
    for (int idx = tid; idx < n * n; idx += PB_THREADS) {
        double v;
        if (job.raw) {
            v = job.raw[idx];
        } else {
            const int i = idx % n, j = idx / n;
            v = 0.0;
            for (int k = 0; k < job.K; ++k) {
                const double* pc = job.psf + 7 * k;
                const double t11 = pc[3], t21 = pc[4], t12 = pc[5], t22 = pc[6];
                const double det = t11 * t22 - t12 * t21;
                const double dx = (double)(i + 1) - c0 - pc[1], dy = (double)(j + 1) - c0 - pc[2];
                const double q = (t22 * dx * dx - (t12 + t21) * dx * dy + t11 * dy * dy) / det;
                v += pc[0] * exp(-0.5 * q) / (2.0 * 3.14159265358979323846 * sqrt(det));
            }
        }
        v = fmax(v, 0.0) + 1e-6;
        g[idx] = v;
        part += v;
    }
    red[tid] = part;
    if (tid == 0) {
        // Thomas factors of the constant (1, 4, 1) system: cpv[i] = 1 / (4 - cpv[i-1]), cpv[0] = 0
        cpv[0] = 0.0;
        for (int i = 1; i < n; ++i) cpv[i] = 1.0 / (4.0 - cpv[i - 1]);
    }
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < PB_THREADS; ++k) t += red[k];        // fixed order
        s_total = t;
    }
    __syncthreads();
    // 2. normalise, softpluslike (fsm_util.jl:221)
    const double total = s_total;
    for (int idx = tid; idx < n * n; idx += PB_THREADS) {
        const double x = 1000.0 * (g[idx] / total);
        g[idx] = x > 1.0 ? x - 1.0 : log(x);
    }
    __syncthreads();
    // 3. prefilter along the first axis (one column per thread), then along the second (one row per thread)
    for (int j = tid; j < n; j += PB_THREADS) prefilter_line(g + (size_t)j * n, 1, T + (size_t)j * n2, 1, n, cpv);
    __syncthreads();
    for (int r = tid; r < n2; r += PB_THREADS) prefilter_line(T + r, n2, job.coefs + r, n2, n, cpv);
}

struct BitmapJob {
    int image, off_h, off_w, H2, W2;
    unsigned char* bitmap;
};

__global__ void bitmap_build_kernel(const ImageDev* __restrict__ images, const BitmapJob* __restrict__ jobs) {
    const BitmapJob job = jobs[blockIdx.x];
    const ImageDev img = images[job.image];
    const int npix = job.H2 * job.W2;
    for (int idx = threadIdx.x; idx < npix; idx += blockDim.x) {
        const int h2 = idx % job.H2, w2 = idx / job.H2;
        const float x = img.pixels[(size_t)(job.off_h + h2) + (size_t)(job.off_w + w2) * img.H];
        job.bitmap[idx] = isnan(x) ? 0 : 1;
    }
}

// find_neighbors for every target at once.  One warp per target; fill == 0: counts[t] = number of neighbours;
// fill == 1: the neighbours (ascending source index, 0-based) go to out[ptr[t] ..].
__global__ void neighbor_kernel(const PatchDev* __restrict__ patches, int S, int N, int fill, int* __restrict__ counts,
                                const int* __restrict__ ptr, int* __restrict__ out) {
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (t >= S) return;
    int count = 0;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        bool hit = false;
        if (i < S && i != t) {
            for (int n = 0; n < N && !hit; ++n) {
                const PatchDev& a = patches[t + (size_t)n * S];
                const PatchDev& b = patches[i + (size_t)n * S];
                // boxes_overlap (imaged_sources.jl:36-40) on the ranges off+1 : off+size (empty ranges never overlap)
                hit = (a.off_h + 1 <= b.off_h + b.H2) && (b.off_h + 1 <= a.off_h + a.H2) && (a.off_w + 1 <= b.off_w + b.W2) &&
                      (b.off_w + 1 <= a.off_w + a.W2);
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (fill && hit) out[ptr[t] + count + __popc(ballot & ((1u << lane) - 1u))] = i;
        count += __popc(ballot);
    }
    if (!fill && lane == 0) counts[t] = count;
}

}  // namespace celeste
#endif
