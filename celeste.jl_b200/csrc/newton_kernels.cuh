// newton_kernels.cuh -- the trust-region subproblem of the batched Newton driver.
//
// Optim.NewtonTrustRegion (un-vendored dependency of ElboMaximize.jl:105-108,235) solves, per source and
// per iterate,   min_s  g's + 1/2 s'Hs   s.t. |s| <= delta   for the 41 x 41 free-space Hessian.  cuSOLVER's
// batched FP64 eigensolver and LAPACK-on-the-host both cost ~0.5 s per 1000 sources -- 100x the ELBO
// evaluation itself -- so the subproblem gets its own kernel: ONE block per source,
//   1. cyclic two-sided Jacobi eigen-decomposition in shared memory with the round-robin ("tournament")
//      ordering: n/2 disjoint rotations per round are computed and applied in parallel;
//   2. q = V'g; the secular equation |s(lam)| = delta by Newton on 1/|s| from just above the pole
//      (monotone), interior and hard cases handled as in elbo_maximize.solve_tr_subproblem (the torch
//      restatement the tests compare against);
//   3. s = V coef, predicted change m = sum_j (q_j c_j + 1/2 ev_j c_j^2).
#ifndef CELESTE_NEWTON_KERNELS_CUH
#define CELESTE_NEWTON_KERNELS_CUH

#include "celeste_kernels.cuh"

namespace celeste {

constexpr int TR_MAXN = 48;          // padded dimension limit (41 free parameters -> 42)
constexpr int TR_LD = TR_MAXN + 1;   // leading dimension (odd: conflict-free column walks)
constexpr int TR_THREADS = 128;
constexpr int TR_MAX_SWEEPS = 14;

// shared-memory workspace of one subproblem (one block)
struct TrShared {
    double A[TR_MAXN * TR_LD];
    double V[TR_MAXN * TR_LD];
    double rc[TR_MAXN / 2], rs[TR_MAXN / 2];
    int rp[TR_MAXN / 2], rq[TR_MAXN / 2];
    double ev[TR_MAXN], qg[TR_MAXN], coef[TR_MAXN], gsh[TR_MAXN];
    double red[TR_THREADS / 32];
};

// Solve the subproblem held in S (S.A: symmetric n x n matrix zero-padded to the even size np, S.gsh: gradient
// zero-padded) with all TR_THREADS threads of the block; S.A is destroyed.  s_out (n doubles, shared or global),
// *m_out and *interior_out are written by the block; the caller synchronises before reading them.
__device__ inline void tr_solve_block(TrShared& S, int n, double delta, double* s_out, double* m_out, int* interior_out) {
    double* A = S.A;
    double* V = S.V;
    double* rc = S.rc;
    double* rs = S.rs;
    int* rp = S.rp;
    int* rq = S.rq;
    double* ev = S.ev;
    double* qg = S.qg;
    double* coef = S.coef;
    double* gsh = S.gsh;
    double* red = S.red;
    const int tid = threadIdx.x;
    const int np = (n + 1) & ~1;          // even padded size
    const int half = np / 2;
    for (int i = tid; i < np * np; i += TR_THREADS) {
        const int r = i / np, c = i % np;
        V[r * TR_LD + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();

    // block-wide sum helper (fixed order)
    auto block_sum = [&](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < TR_THREADS / 32; ++w) t += red[w];
        __syncthreads();
        return t;
    };

    double frob = 0.0;
    for (int i = tid; i < np * np; i += TR_THREADS) {
        const double a = A[(i / np) * TR_LD + (i % np)];
        frob += a * a;
    }
    frob = block_sum(frob);

    for (int sweep = 0; sweep < TR_MAX_SWEEPS; ++sweep) {
        double off = 0.0;
        for (int i = tid; i < np * np; i += TR_THREADS) {
            const int r = i / np, c = i % np;
            if (r != c) {
                const double a = A[r * TR_LD + c];
                off += a * a;
            }
        }
        off = block_sum(off);
        if (off <= 1e-30 * frob || frob == 0.0) break;
        for (int round = 0; round < np - 1; ++round) {
            // tournament pairing: player np-1 fixed, the others rotate
            if (tid < half) {
                int p, q;
                if (tid == 0) {
                    p = np - 1;
                    q = round;
                } else {
                    p = (round + tid) % (np - 1);
                    q = (round - tid + (np - 1)) % (np - 1);
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                const double apq = A[p * TR_LD + q];
                double c = 1.0, s = 0.0;
                if (fabs(apq) > 1e-300) {
                    const double tau = (A[q * TR_LD + q] - A[p * TR_LD + p]) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                rp[tid] = p;
                rq[tid] = q;
                rc[tid] = c;
                rs[tid] = s;
            }
            __syncthreads();
            // A <- J' A J by 2 x 2 blocks: block (k', k) = rows (p', q') of pair k', columns (p, q) of pair k; each
            // element is read and written once per round (column rotation k, then row rotation k', in registers)
            for (int i = tid; i < half * half; i += TR_THREADS) {
                const int k = i / half, kr = i - k * half;
                const int p = rp[k], q = rq[k], pr = rp[kr], qr = rq[kr];
                const double c = rc[k], s = rs[k], cr = rc[kr], sr = rs[kr];
                const double a00 = A[pr * TR_LD + p], a01 = A[pr * TR_LD + q];
                const double a10 = A[qr * TR_LD + p], a11 = A[qr * TR_LD + q];
                const double b00 = c * a00 - s * a01, b01 = s * a00 + c * a01;     // columns
                const double b10 = c * a10 - s * a11, b11 = s * a10 + c * a11;
                A[pr * TR_LD + p] = cr * b00 - sr * b10;                             // rows
                A[qr * TR_LD + p] = sr * b00 + cr * b10;
                A[pr * TR_LD + q] = cr * b01 - sr * b11;
                A[qr * TR_LD + q] = sr * b01 + cr * b11;
            }
            // V <- V J (columns only)
            for (int i = tid; i < half * np; i += TR_THREADS) {
                const int k = i / np, r = i % np;
                const int p = rp[k], q = rq[k];
                const double c = rc[k], s = rs[k];
                const double vp = V[r * TR_LD + p], vq = V[r * TR_LD + q];
                V[r * TR_LD + p] = c * vp - s * vq;
                V[r * TR_LD + q] = s * vp + c * vq;
            }
            __syncthreads();
        }
    }

    // eigenvalues / q = V'g
    if (tid < n) {
        ev[tid] = A[tid * TR_LD + tid];
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += V[i * TR_LD + tid] * gsh[i];
        qg[tid] = t;
    }
    __syncthreads();

    // secular equation: warp 0, lanes hold entries j and j + 32
    if (tid < 32) {
        const double d2 = delta * delta;
        const int j0 = tid, j1 = tid + 32;
        const bool h0 = j0 < n, h1 = j1 < n;
        const double e0 = h0 ? ev[j0] : 0.0, e1 = h1 ? ev[j1] : 0.0;
        const double q0 = h0 ? qg[j0] : 0.0, q1 = h1 ? qg[j1] : 0.0;
        auto wsum = [&](double v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            return v;
        };
        auto wmin = [&](double v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
            return v;
        };
        auto wmax = [&](double v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
            return v;
        };
        const double big = 1.797e308;
        const double lam_min = wmin(fmin(h0 ? e0 : big, h1 ? e1 : big));
        const double evabs_max = wmax(fmax(h0 ? fabs(e0) : 0.0, h1 ? fabs(e1) : 0.0));
        auto pnorm2 = [&](double lam) {
            const double a0 = h0 ? q0 / (e0 + lam) : 0.0, a1 = h1 ? q1 / (e1 + lam) : 0.0;
            return wsum(a0 * a0 + a1 * a1);
        };
        const bool pos_def = lam_min >= 1e-8;
        const bool interior = pos_def && (pnorm2(0.0) <= d2);
        const double lam_lb = fmax(-lam_min, 0.0);
        const double tiny = 1e-12 * (1.0 + evabs_max);
        // hard case
        const double scale = wmax(fmax(fabs(q0), fabs(q1))) + 1e-300;
        const bool m0 = h0 && fabs(e0 - lam_min) <= 1e-12 * (1.0 + fabs(e0));
        const bool m1 = h1 && fabs(e1 - lam_min) <= 1e-12 * (1.0 + fabs(e1));
        const bool ok0 = !m0 || fabs(q0) <= 1e-12 * scale, ok1 = !m1 || fabs(q1) <= 1e-12 * scale;
        const bool ortho = __all_sync(0xffffffffu, ok0 && ok1);
        const double c0h = (h0 && !m0) ? -q0 / (e0 - lam_min) : 0.0;
        const double c1h = (h1 && !m1) ? -q1 / (e1 - lam_min) : 0.0;
        const double ph2 = wsum(c0h * c0h + c1h * c1h);
        const bool hard = !interior && ortho && lam_min <= 1e-8 && ph2 <= d2;
        double lam = lam_lb + tiny;
        if (!interior && !hard) {
            for (int it = 0; it < 60; ++it) {
                const double dn0 = e0 + lam, dn1 = e1 + lam;
                const double a0 = h0 ? q0 / dn0 : 0.0, a1 = h1 ? q1 / dn1 : 0.0;
                const double p2 = wsum(a0 * a0 + a1 * a1);
                const double dp2 = -2.0 * wsum((h0 ? a0 * a0 / dn0 : 0.0) + (h1 ? a1 * a1 / dn1 : 0.0));
                const double pn = sqrt(p2);
                const double step = (pn - delta) / delta * p2 / (-0.5 * dp2 + 1e-300);
                double nw = lam + step;
                if (nw <= lam_lb) nw = 0.5 * (lam + lam_lb) + tiny;
                if (fabs(step) <= 1e-12 * (1.0 + fabs(lam))) break;
                lam = nw;
            }
        }
        if (interior) lam = 0.0;
        if (hard) lam = -lam_min;
        double c0 = 0.0, c1 = 0.0;
        if (hard) {
            c0 = c0h;
            c1 = c1h;
            // move along one eigenvector of lam_min to the boundary: the lowest-index lane holding lam_min
            const unsigned who = __ballot_sync(0xffffffffu, m0) ;
            const unsigned who1 = __ballot_sync(0xffffffffu, m1);
            const double tau = sqrt(fmax(d2 - ph2, 0.0));
            if (who) {
                if (tid == __ffs(who) - 1) c0 = tau;
            } else if (who1) {
                if (tid == __ffs(who1) - 1) c1 = tau;
            }
        } else {
            double dn0 = e0 + lam, dn1 = e1 + lam;
            if (fabs(dn0) < 1e-300) dn0 = 1e-300;
            if (fabs(dn1) < 1e-300) dn1 = 1e-300;
            c0 = h0 ? -q0 / dn0 : 0.0;
            c1 = h1 ? -q1 / dn1 : 0.0;
        }
        if (h0) coef[j0] = c0;
        if (h1) coef[j1] = c1;
        const double m = wsum((h0 ? q0 * c0 + 0.5 * e0 * c0 * c0 : 0.0) + (h1 ? q1 * c1 + 0.5 * e1 * c1 * c1 : 0.0));
        if (tid == 0) {
            *m_out = m;
            *interior_out = interior ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid < n) {
        double t = 0.0;
        for (int j = 0; j < n; ++j) t += V[tid * TR_LD + j] * coef[j];
        s_out[tid] = t;
    }
}

// mask (nullable): sources with mask[b] == 0 are skipped and their outputs left untouched
__global__ void __launch_bounds__(TR_THREADS) tr_subproblem_kernel(int n, const double* __restrict__ g_all,
                                                                   const double* __restrict__ H_all,
                                                                   const double* __restrict__ delta_all,
                                                                   const unsigned char* __restrict__ mask,
                                                                   double* __restrict__ s_all, double* __restrict__ m_all,
                                                                   int* __restrict__ interior_all) {
    __shared__ TrShared S;
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    if (mask && !mask[b]) return;
    const int np = (n + 1) & ~1;
    const double* H = H_all + (size_t)b * n * n;
    const double* g = g_all + (size_t)b * n;
    for (int i = tid; i < np * np; i += TR_THREADS) {
        const int r = i / np, c = i % np;
        // symmetrised load (the free-space Hessian is symmetric by construction; this guards round-off)
        S.A[r * TR_LD + c] = (r < n && c < n) ? 0.5 * (H[(size_t)r * n + c] + H[(size_t)c * n + r]) : 0.0;
    }
    if (tid < np) S.gsh[tid] = tid < n ? g[tid] : 0.0;
    __syncthreads();
    tr_solve_block(S, n, delta_all[b], s_all + (size_t)b * n, m_all + b, interior_all + b);
}

}  // namespace celeste
#endif
