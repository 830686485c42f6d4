// newton_kernels.cuh -- the trust-region subproblem of the batched Newton driver.
//
// Optim.NewtonTrustRegion (un-vendored dependency of ElboMaximize.jl:105-108,235) solves, per source and
// per iterate,   min_s  g's + 1/2 s'Hs   s.t. |s| <= delta   for the 41 x 41 free-space Hessian.  cuSOLVER's
// batched FP64 eigensolver and LAPACK-on-the-host both cost ~0.5 s per 1000 sources -- 100x the ELBO
// evaluation itself -- so the subproblem gets its own code, ONE block per source.  Two implementations:
// tr_solve_block (the product: Householder tridiagonalisation + O(n) work on the tridiagonal matrix, further down)
// and tr_solve_block_jacobi (round 1: a full eigen-decomposition; kept as the A/B reference, -DCELESTE_TR_JACOBI):
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
__device__ inline void tr_solve_block_jacobi(TrShared& S, int n, double delta, double* s_out, double* m_out, int* interior_out) {
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


// 1 / x for the sequential tridiagonal recurrences: one reciprocal per step IS their critical path.  __drcp_rn is a
// correctly rounded subroutine (~100 cycles of dependent instructions); by default (CELESTE_TR_FAST_RCP = 1) the
// hardware's ~20-bit approximation (MUFU.RCP64H) is refined by NR Newton steps of two dependent FMAs each: 2 steps = full
// double precision to an ulp (not correctly rounded), 1 step (~1e-12) is enough where only the sign of the result is
// used.  |x| is always a normal number here (the callers clamp the pivots to 1e-300).  Measured: newton_step_kernel
// 2.97 -> 2.6 ms per 10 000 sources, maximize leg 49.9 k -> 53.9 k sources/s (-DCELESTE_TR_FAST_RCP=0: __drcp_rn).
#ifndef CELESTE_TR_FAST_RCP
#define CELESTE_TR_FAST_RCP 1
#endif
template <int NR = 2>
__device__ __forceinline__ double tr_rcp(double x) {
#if defined(__CUDA_ARCH__) && CELESTE_TR_FAST_RCP
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const double e = fma(-x, r, 1.0);
        r = fma(r, e, r);
    }
    return r;
#elif defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// ------------------------------------------------------------------------------------------------
// The same subproblem WITHOUT an eigen-decomposition (the Jacobi version above costs ~8 sweeps x n^3 and was as
// long as the ELBO evaluation it follows: 10 ms per 10 000 sources):
//   1. Householder tridiagonalisation  H = Q T Q'  (4/3 n^3 flops, all threads of the block; reflectors kept in the
//      lower triangle of S.A);
//   2. everything that depended on eigenvalues is done on T in O(n) per evaluation by warp 0:
//      lam_min by 32-way multisection on Sturm counts; |s(lam)| and d|s|^2/dlam by two solves with the LDL'
//      factorisation of T + lam I (positive definite for lam > -lam_min); the secular equation by the SAME Newton
//      iteration on 1/|s| from just above the pole, with the same stopping rule, as the eigenbasis version and the
//      torch restatement (elbo_maximize.solve_tr_subproblem): in exact arithmetic the iterates coincide;
//      the hard case (g orthogonal to the eigenvector z of lam_min, |s(-lam_min)| <= delta) with z from inverse
//      iteration on T;
//   3. s = Q s~ by the reflectors; predicted change m = g~'s~ + 1/2 s~'T s~.
// Same interface as tr_solve_block_jacobi; S.A and S.V are destroyed.
__device__ inline void tr_solve_block(TrShared& S, int n, double delta, double* s_out, double* m_out, int* interior_out) {
#ifdef CELESTE_TR_JACOBI               // A/B knob: the eigen-decomposition version
    tr_solve_block_jacobi(S, n, delta, s_out, m_out, interior_out);
    return;
#endif
    double* A = S.A;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // work vectors carved out of S.V (free here)
    double* hv = S.V;                 // current reflector v (length m)
    double* hp = S.V + TR_MAXN;       // p = beta A22 v, then w
    double* tau = S.V + 2 * TR_MAXN;  // beta_k of every reflector
    double* v0s = S.V + 3 * TR_MAXN;  // v_0 of every reflector (its slot in A holds the sub-diagonal entry)
    double* sc = S.V + 9 * TR_MAXN;   // scalars: [0] beta, [1] alpha, [2] K
    double* dg = S.ev;                // diagonal of T
    double* od = S.qg;                // off-diagonal: od[i] couples i and i + 1
    double* gt = S.coef;              // g~ = Q'g, later s~
    double* gsh = S.gsh;
    auto wsum = [&](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };

    // ---- 1. tridiagonalisation -------------------------------------------------------------------------------
    for (int k = 0; k + 2 < n; ++k) {
        const int m = n - k - 1;                  // x = A[k+1 .. n-1][k]
        if (warp == 0) {
            const int i0 = lane, i1 = lane + 32;
            const double x0 = i0 < m ? A[(k + 1 + i0) * TR_LD + k] : 0.0;
            const double x1 = i1 < m ? A[(k + 1 + i1) * TR_LD + k] : 0.0;
            const double tail = wsum((i0 >= 1 ? x0 * x0 : 0.0) + x1 * x1);
            const double xf = __shfl_sync(0xffffffffu, x0, 0);
            double beta = 0.0, alpha = xf;
            if (tail > 0.0) {
                alpha = -copysign(sqrt(xf * xf + tail), xf);
                const double v0 = xf - alpha;
                beta = 2.0 / (v0 * v0 + tail);
                if (i0 < m) hv[i0] = i0 == 0 ? v0 : x0;
                if (i1 < m) hv[i1] = x1;
            } else {
                if (i0 < m) hv[i0] = 0.0;
                if (i1 < m) hv[i1] = 0.0;
            }
            if (lane == 0) {
                sc[0] = beta;
                sc[1] = alpha;
                tau[k] = beta;
            }
        }
        __syncthreads();
        const double beta = sc[0];
        if (beta != 0.0) {                         // block-uniform
            {
                // p = beta A22 v: two threads per row (TR_THREADS / 2 >= TR_MAXN rows), halves of the dot product
                const int r = tid >> 1, part = tid & 1;
                double t = 0.0;
                if (r < m) {
                    const double* row = A + (k + 1 + r) * TR_LD + (k + 1);
                    for (int j = part; j < m; j += 2) t = fma(row[j], hv[j], t);
                }
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                if (r < m && part == 0) hp[r] = beta * t;
            }
            __syncthreads();
            if (warp == 0) {
                const double t = wsum((lane < m ? hv[lane] * hp[lane] : 0.0) + (lane + 32 < m ? hv[lane + 32] * hp[lane + 32] : 0.0));
                if (lane == 0) sc[2] = 0.5 * beta * t;
            }
            __syncthreads();
            const double K = sc[2];
            // A22 <- A22 - v w' - w v',  w = p - K v: thread (j, i0) owns column j of the rows i0, i0 + 2, ...
            // (no integer division by m in the inner loop)
            {
                const int j = tid & 63, i0 = tid >> 6;
                if (j < m) {
                    const double vj = hv[j], wj = hp[j] - K * vj;
                    for (int i = i0; i < m; i += TR_THREADS / 64) {
                        const double vi = hv[i], wi = hp[i] - K * vi;
                        A[(k + 1 + i) * TR_LD + (k + 1 + j)] -= vi * wj + wi * vj;
                    }
                }
            }
        }
        // keep the reflector in column k (below the sub-diagonal position) and the sub-diagonal entry
        if (tid < m) A[(k + 1 + tid) * TR_LD + k] = tid == 0 ? sc[1] : hv[tid];
        if (tid == 0) v0s[k] = hv[0];
        __syncthreads();
    }
    if (tid < n) dg[tid] = A[tid * TR_LD + tid];
    if (tid + 1 < n) od[tid] = A[(tid + 1) * TR_LD + tid];
    if (tid < n) gt[tid] = gsh[tid];
    __syncthreads();

    if (warp == 0) {
        // reflector k acts on entries k+1 .. n-1: v = (v0, A[k+2..][k])
        auto apply = [&](int k, double* x) {
            const double beta = tau[k];
            if (beta == 0.0) return;
            const int m = n - k - 1;
            const int i0 = lane, i1 = lane + 32;
            const double v0 = i0 < m ? (i0 == 0 ? v0s[k] : A[(k + 1 + i0) * TR_LD + k]) : 0.0;
            const double v1 = i1 < m ? A[(k + 1 + i1) * TR_LD + k] : 0.0;
            const double t = beta * wsum((i0 < m ? v0 * x[k + 1 + i0] : 0.0) + (i1 < m ? v1 * x[k + 1 + i1] : 0.0));
            if (i0 < m) x[k + 1 + i0] -= t * v0;
            if (i1 < m) x[k + 1 + i1] -= t * v1;
            __syncwarp();
        };
        // ---- g~ = Q'g = H_{n-3} ... H_0 g
        for (int k = 0; k + 2 < n; ++k) apply(k, gt);

        // ---- 2. scalars on T -------------------------------------------------------------------------------
        // Gershgorin bounds
        double glo = 1.797e308, ghi = -1.797e308;
        for (int i = lane; i < n; i += 32) {
            const double r = (i > 0 ? fabs(od[i - 1]) : 0.0) + (i + 1 < n ? fabs(od[i]) : 0.0);
            glo = fmin(glo, dg[i] - r);
            ghi = fmax(ghi, dg[i] + r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            glo = fmin(glo, __shfl_xor_sync(0xffffffffu, glo, o));
            ghi = fmax(ghi, __shfl_xor_sync(0xffffffffu, ghi, o));
        }
        const double evabs_max = fmax(fabs(glo), fabs(ghi));
        // number of eigenvalues of T below x (Sturm count by the pivots of T - x I)
        double* od2 = S.V + 10 * TR_MAXN;            // squared off-diagonals
        for (int i = lane; i + 1 < n; i += 32) od2[i] = od[i] * od[i];
        __syncwarp();
        auto below = [&](double x) {
            int cnt = 0;
            double q = dg[0] - x;
            if (q < 0.0) ++cnt;
            for (int i = 1; i < n; ++i) {
                if (fabs(q) < 1e-300) q = q < 0.0 ? -1e-300 : 1e-300;
                q = fma(-od2[i - 1], tr_rcp<1>(q), dg[i] - x);
                if (q < 0.0) ++cnt;
            }
            return cnt;
        };
        // Positive definite (no eigenvalue below 1e-8)?  One Sturm count.  lam_min itself is needed only when T is not:
        // then 32-way multisection of [glo, min(ghi, 1e-8)]
        const bool pos_def = below(1e-8) == 0;
        double lam_min = 1e-8;
        if (!pos_def) {
            double lo = glo - 1e-12 * (1.0 + evabs_max), hi = fmin(ghi, 1e-8) + 1e-12 * (1.0 + evabs_max);
            for (int round = 0; round < 14; ++round) {
                const double x = lo + (hi - lo) * (double)(lane + 1) / 33.0;
                const unsigned hit = __ballot_sync(0xffffffffu, below(x) >= 1);
                const int first = hit ? __ffs(hit) - 1 : 32;             // first lane whose x has an eigenvalue below it
                const double nlo = first == 0 ? lo : lo + (hi - lo) * (double)first / 33.0;
                const double nhi = first == 32 ? hi : lo + (hi - lo) * (double)(first + 1) / 33.0;
                lo = nlo;
                hi = nhi;
                if (!(hi - lo > 4.0e-16 * (1.0 + evabs_max))) break;
            }
            lam_min = 0.5 * (lo + hi);
        }

        const double d2 = delta * delta;
        // LDL' solves on lane 0 (results broadcast); x_out may alias rhs
        double* sv = S.V + 4 * TR_MAXN;  // s(lam)
        double* wv = S.V + 5 * TR_MAXN;  // u = L^-1 s  (s'(T + lam I)^-1 s = sum u_i^2 / q_i)
        double* piv = S.V + 6 * TR_MAXN; // reciprocal pivots of T + lam I
        double* lf = S.V + 11 * TR_MAXN; // its unit lower bidiagonal factor
        auto solve_pair = [&](double lam, bool need_w, double& p2, double& sw) {
            // lane 0: T + lam I = L D L', s = -(T + lam I)^-1 g~, and optionally u = L^-1 s for
            // s'(T + lam I)^-1 s = sum u_i^2 / q_i.  Every recurrence carries its running value in a register (the
            // chain per step is then one rcp + fma for the pivots, one fma for a forward solve, fma + mul for the back
            // solve -- not a shared-memory store / load round trip), and the forward solve of g~ rides along with the
            // factorisation (two independent chains).
            if (lane == 0) {
                double q = dg[0] + lam;
                if (fabs(q) < 1e-300) q = 1e-300;
                double rq = tr_rcp(q);             // piv[i] = 1 / q_i, lf[i] = od[i - 1] / q_{i-1} (the L factor)
                piv[0] = rq;
                double y = -gt[0];
                sv[0] = y;
                for (int i = 1; i < n; ++i) {
                    const double l = od[i - 1] * rq;
                    lf[i] = l;
                    q = fma(-od2[i - 1], rq, dg[i] + lam);
                    if (fabs(q) < 1e-300) q = 1e-300;
                    rq = tr_rcp(q);
                    piv[i] = rq;
                    y = fma(-l, y, -gt[i]);        // forward: y_i = b_i - l_i y_{i-1}
                    sv[i] = y;
                }
                double x = y * rq;                 // back: x_i = (y_i - od_i x_{i+1}) / q_i
                sv[n - 1] = x;
                for (int i = n - 2; i >= 0; --i) {
                    x = fma(-od[i], x, sv[i]) * piv[i];
                    sv[i] = x;
                }
                if (need_w) {
                    double u = sv[0];
                    wv[0] = u;
                    for (int i = 1; i < n; ++i) {
                        u = fma(-lf[i], u, sv[i]);
                        wv[i] = u;
                    }
                }
            }
            __syncwarp();
            double a = 0.0, b = 0.0;
            for (int i = lane; i < n; i += 32) {
                a = fma(sv[i], sv[i], a);
                if (need_w) b = fma(wv[i] * wv[i], piv[i], b);
            }
            p2 = wsum(a);
            sw = wsum(b);
            __syncwarp();
        };
        double p2 = 0.0, sw = 0.0;
        bool interior = false;
        if (pos_def) {
            solve_pair(0.0, false, p2, sw);
            interior = p2 <= d2;
        }
        const double lam_lb = fmax(-lam_min, 0.0);
        const double tiny = 1e-12 * (1.0 + evabs_max);
        // Hard case (g orthogonal to the eigenvector z of lam_min and |s(-lam_min)| <= delta): detected by the first
        // evaluation just above the pole -- any component of g~ along z would make |s| there astronomically large.
        // Only then is z computed (inverse iteration on T) and the step completed along it to the boundary.
        bool hard = false;
        double* zv = S.V + 7 * TR_MAXN;
        double* sh = S.V + 8 * TR_MAXN;
        double ph2 = 0.0;
        double lam = lam_lb + tiny;
        if (!interior) {
            solve_pair(lam, true, p2, sw);
            if (!pos_def && p2 <= d2) {
                hard = true;
                if (lane == 0) {
                    const double shift = -lam_min + 4.0 * tiny * 1e-3;
                    double q = dg[0] + shift;
                    if (fabs(q) < 1e-300) q = 1e-300;
                    double rq = tr_rcp(q);
                    piv[0] = rq;
                    for (int i = 1; i < n; ++i) {
                        lf[i] = od[i - 1] * rq;
                        q = fma(-od[i - 1], lf[i], dg[i] + shift);
                        if (fabs(q) < 1e-300) q = 1e-300;
                        rq = tr_rcp(q);
                        piv[i] = rq;
                    }
                    for (int i = 0; i < n; ++i) zv[i] = ((i & 1) ? 1.0 : 0.73);
                    for (int it = 0; it < 3; ++it) {
                        for (int i = 1; i < n; ++i) zv[i] = fma(-lf[i], zv[i - 1], zv[i]);
                        zv[n - 1] *= piv[n - 1];
                        for (int i = n - 2; i >= 0; --i) zv[i] = fma(-od[i], zv[i + 1], zv[i]) * piv[i];
                        double nr = 0.0;
                        for (int i = 0; i < n; ++i) nr = fma(zv[i], zv[i], nr);
                        nr = 1.0 / sqrt(nr);
                        for (int i = 0; i < n; ++i) zv[i] *= nr;
                    }
                }
                __syncwarp();
                // s_h = s(lam_lb + tiny) with its z component removed
                double c = 0.0;
                for (int i = lane; i < n; i += 32) c = fma(zv[i], sv[i], c);
                c = wsum(c);
                double a = 0.0;
                for (int i = lane; i < n; i += 32) {
                    sh[i] = sv[i] - c * zv[i];
                    a = fma(sh[i], sh[i], a);
                }
                ph2 = wsum(a);
                __syncwarp();
            } else {
                for (int it = 0; it < 60; ++it) {
                    if (it > 0) solve_pair(lam, true, p2, sw);
                    const double pn = sqrt(p2);
                    // d|s|^2 / dlam = -2 s'(T + lam I)^-1 s
                    const double step = (pn - delta) / delta * p2 / (sw + 1e-300);
                    double nw = lam + step;
                    if (nw <= lam_lb) nw = 0.5 * (lam + lam_lb) + tiny;
                    if (fabs(step) <= 1e-12 * (1.0 + fabs(lam))) break;
                    lam = nw;
                }
            }
        }
        if (interior) {
            lam = 0.0;
            solve_pair(0.0, false, p2, sw);
        } else if (!hard) {
            solve_pair(lam, false, p2, sw);
        }
        // s~ -> gt's place is still needed for m: keep s~ in sv (or sh + tau z in the hard case)
        if (hard) {
            const double tz = sqrt(fmax(d2 - ph2, 0.0));
            for (int i = lane; i < n; i += 32) sv[i] = sh[i] + tz * zv[i];
            __syncwarp();
        }
        // m = g~'s~ + 1/2 s~'T s~
        double mm = 0.0;
        for (int i = lane; i < n; i += 32) {
            double ts = dg[i] * sv[i];
            if (i > 0) ts = fma(od[i - 1], sv[i - 1], ts);
            if (i + 1 < n) ts = fma(od[i], sv[i + 1], ts);
            mm += sv[i] * (gt[i] + 0.5 * ts);
        }
        mm = wsum(mm);
        if (lane == 0) {
            *m_out = mm;
            *interior_out = interior ? 1 : 0;
        }
        // ---- 3. s = Q s~ = H_0 H_1 ... H_{n-3} s~
        __syncwarp();
        for (int k = n - 3; k >= 0; --k) apply(k, sv);
        for (int i = lane; i < n; i += 32) s_out[i] = sv[i];
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
