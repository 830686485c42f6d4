// maximize_kernels.cuh -- one lock-step Newton trust-region iteration for every source of a batch, fused into one
// kernel per iteration (SURVEY.md 8, rows f.1 + f.2 + f.3).
//
// What the reference does per iterate of one source on the host (ElboMaximize.evaluate!, ElboMaximize.jl:161-172,
// then Optim.NewtonTrustRegion, :105-108,235):
//     to_bound!  ->  elbo (likelihood - KL)  ->  propagate_derivatives!  ->  trust-region bookkeeping + subproblem
// Here the likelihood of ALL sources is one plan evaluation (celeste_kernels.cuh) and everything else is
// newton_step_kernel, one block per source, nothing leaving the device:
//   1. -KL(q || p) with gradient and Hessian in closed form (elbo_kl.jl:94-154: categorical a and k, Gaussian
//      flux, diag-MVN vs the 8-component colour mixture, radius prior) added to the 44-space likelihood result;
//   2. propagate_derivatives! (ConstraintTransforms.jl:373-396): J'g and J'HJ + sum_i g_i d2b_i for the 26 scaled
//      logistic boxes and the three softmax simplexes (constraint table ElboMaximize.jl:63-93), symmetrised;
//   3. the Optim.NewtonTrustRegion update (rho test, radius update, accept/reject, x/f/g convergence tests of
//      Optim.Options(x_tol 1e-7, f_tol 1e-6, g_tol 1e-8), ElboMaximize.jl:95-103);
//   4. for sources still iterating: the exact subproblem (tr_solve_block, newton_kernels.cuh), the candidate
//      x + s, and to_bound! of the candidate written straight into the plan's parameter array.
// The `active` bytes double as the plan's task mask: the likelihood kernels skip converged sources.
// The torch code in elbo_maximize.py / kl.py / constraint_transforms.py is the same algorithm for CPU tensors and
// is what the tests compare this kernel against.  PARITY UNPINNED vs Optim.jl internals (un-vendored dependency).
#ifndef CELESTE_MAXIMIZE_KERNELS_CUH
#define CELESTE_MAXIMIZE_KERNELS_CUH

#include "newton_kernels.cuh"

namespace celeste {

constexpr int NW_BOUND = 44, NW_FREE = 41, NW_BOX = 26;
// packed prior (kl.KLTerm.packed): log pi_a[2], flux_mean[2], flux_var[2], log pi_k[2][8], mu[2][8][4],
// precision[2][8][4][4], logdet[2][8], radius mean, radius var
constexpr int PR_LOG_A = 0, PR_FLUX_MEAN = 2, PR_FLUX_VAR = 4, PR_LOG_K = 6, PR_MU = 22, PR_PREC = 86, PR_LOGDET = 342,
              PR_RAD = 358, PR_LEN = 360;

struct NewtonDev {
    // accepted iterate, minimisation form (f = -ELBO), free coordinates
    double* x;          // B x 41
    double* f;          // B
    double* g;          // B x 41
    double* H;          // B x 41 x 41
    double* delta;      // B
    // candidate under evaluation and what the subproblem predicted for it
    double* x_new;      // B x 41
    double* m_pred;     // B
    int* interior;      // B
    unsigned char* active;     // B   (also the plan's task mask)
    unsigned char* converged;  // B
    int* iters;         // B
    int* f_calls;       // B   ELBO evaluations consumed
    const double* lo;   // B x 26
    const double* hi;   // B x 26
    // the plan's outputs at the candidate (44-space likelihood)
    const double* v;    // B
    const double* d;    // B x 44
    const double* h;    // B x 44 x 44
    const int* flags;   // B
    double* vp_all;     // n_slots x 44: bound parameters the plan reads
    const long long* aslot;    // B: slot of each source's own parameters in vp_all
    const double* prior;       // PR_LEN doubles, or null: no KL term
    long long h_layout;        // 0: h is B x 44 x 44; 1: B x 406 (CELESTE_HESS_PACKED28)
};

constexpr double NW_ETA = 0.1, NW_RHO_LOWER = 0.25, NW_RHO_UPPER = 0.75;
constexpr double NW_X_TOL = 1e-7, NW_F_TOL = 1e-6, NW_G_TOL = 1e-8;
constexpr double NW_INITIAL_DELTA = 1.0, NW_DELTA_HAT = 1e9;

__host__ __device__ inline int simplex_first(int s) { return s == 0 ? 26 : (s == 1 ? 28 : 36); }
__host__ __device__ inline int simplex_n(int s) { return s == 0 ? 2 : 8; }
__host__ __device__ inline double simplex_lower(int s) { return s == 0 ? 0.005 : 0.01 / 8; }
__host__ __device__ inline int simplex_free0(int s) { return s == 0 ? 26 : (s == 1 ? 27 : 34); }
__host__ __device__ inline int simplex_pp0(int s) { return s == 0 ? 0 : (s == 1 ? 2 : 10); }   // offset into pp[18]
__host__ __device__ inline int simplex_of_free(int a) { return a == 26 ? 0 : (a < 34 ? 1 : 2); }
// canonical (0-based) id of local KL coordinate t of source type i: (a, flux_loc, flux_scale, color_mean[4],
// color_var[4], k[8])
__host__ __device__ inline int kl_index(int i, int t) {
    if (t == 0) return 26 + i;
    if (t == 1) return 6 + i;
    if (t == 2) return 8 + i;
    if (t < 7) return 10 + 4 * i + (t - 3);
    if (t < 11) return 18 + 4 * i + (t - 7);
    return 28 + 8 * i + (t - 11);
}

// to_bound! (ConstraintTransforms.jl:67-70, 89-111) of the 41 free values xf (shared) -> out[44]; also leaves the
// logistic values in sig[26] and the simplex probabilities in pp[18].  Block-wide; caller synchronises after.
__device__ inline void to_bound_block(const double* xf, const double* lo, const double* hi, double* sig, double* pp,
                                      double* out) {
    const int tid = threadIdx.x;
    if (tid < NW_BOX) {
        const double sg = 1.0 / (1.0 + exp(-xf[tid]));
        sig[tid] = sg;
        out[tid] = sg * (hi[tid] - lo[tid]) + lo[tid];
    } else if (tid >= 32 && tid < 35) {
        const int s = tid - 32, n = simplex_n(s), m = n - 1, f0 = simplex_free0(s), first = simplex_first(s);
        const double lower = simplex_lower(s);
        double* p = pp + simplex_pp0(s);
        double mx = xf[f0];                    // the reference's max runs over the free entries only (:97)
        for (int j = 1; j < m; ++j) mx = fmax(mx, xf[f0 + j]);
        double tot = 0.0;
        for (int j = 0; j < n; ++j) {
            const double e = exp((j < m ? xf[f0 + j] : 0.0) - mx);
            p[j] = e;
            tot += e;
        }
        for (int j = 0; j < n; ++j) {
            p[j] /= tot;
            out[first + j] = (1.0 - n * lower) * p[j] + lower;
        }
    }
}

// J[row][a]: derivative of bound parameter (first + j) of simplex s w.r.t. its free logit ja
__device__ inline double simplex_jac(const double* p, double alpha, int j, int ja) {
    return alpha * p[j] * ((j == ja ? 1.0 : 0.0) - p[ja]);
}

// phase 0: first evaluation (at x);  phase 1: evaluation of the candidate x_new;  phase 2: to_bound!(x) -> vp_all
#ifndef CELESTE_NEWTON_MINB
#define CELESTE_NEWTON_MINB 4
#endif
__global__ void __launch_bounds__(TR_THREADS, CELESTE_NEWTON_MINB) newton_step_kernel(NewtonDev nb, int phase) {
    __shared__ TrShared S;
    __shared__ double gb[NW_BOUND], bnd[NW_BOUND], xe[NW_FREE], sig[NW_BOX], d1[NW_BOX], d2[NW_BOX], pp[18], gnew[TR_MAXN];
    __shared__ double sbuf[TR_MAXN], xcand[NW_FREE];
    __shared__ double k_th[19], k_lk[8], k_Pd[32], k_dDv[32], k_D[8], k_Gg[18], k_sc[4];
    __shared__ double sc[8];
    __shared__ int flg[4];

    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const double* lo = nb.lo + (size_t)b * NW_BOX;
    const double* hi = nb.hi + (size_t)b * NW_BOX;
    double* vp_own = nb.vp_all + (size_t)nb.aslot[b] * NW_BOUND;

    if (phase == 2) {                                   // maximize! :239-240
        if (tid < NW_FREE) xe[tid] = nb.x[(size_t)b * NW_FREE + tid];
        __syncthreads();
        to_bound_block(xe, lo, hi, sig, pp, bnd);
        __syncthreads();
        if (tid < NW_BOUND) vp_own[tid] = bnd[tid];
        return;
    }
    if (phase == 1 && !nb.active[b]) return;

    // ---- load the evaluation: 44-space likelihood value / gradient / Hessian, the point it was taken at
    const double* xsrc = (phase == 0 ? nb.x : nb.x_new) + (size_t)b * NW_FREE;
    const bool bad = nb.flags[b] != 0;
    if (tid < NW_FREE) xe[tid] = xsrc[tid];
    if (tid < NW_BOUND) {
        gb[tid] = nb.d[(size_t)b * NW_BOUND + tid];
        bnd[tid] = vp_own[tid];
    }
    double* Hb = S.A;                                   // 44 x 44, leading dimension TR_LD
    if (nb.h_layout == 1) {
        // packed upper triangle of the live 28 x 28 block; every other entry of the 44 x 44 matrix is zero
        const double* h = nb.h + (size_t)b * HESS_PACKED_LEN;
        for (int e = tid; e < NW_BOUND * NW_BOUND; e += TR_THREADS) {
            const int r = e / NW_BOUND, c = e % NW_BOUND;
            const int lo_i = r < c ? r : c, hi_i = r < c ? c : r;
            Hb[r * TR_LD + c] = hi_i < NLIVE ? h[hess_packed_index(lo_i, hi_i)] : 0.0;
        }
    } else {
        const double* h = nb.h + (size_t)b * NW_BOUND * NW_BOUND;
        for (int e = tid; e < NW_BOUND * NW_BOUND; e += TR_THREADS) Hb[(e / NW_BOUND) * TR_LD + (e % NW_BOUND)] = h[e];
    }
    if (tid == 0) sc[0] = 0.0;                          // - KL value
    __syncthreads();

    // ---- 1. subtract_kl (elbo_kl.jl:143-154) in closed form
    if (nb.prior) {
        const double* pr = nb.prior;
        for (int i = 0; i < 2; ++i) {
            const double* P = pr + PR_PREC + i * 128;   // [d][j][k]
            const double* mu = pr + PR_MU + i * 32;     // [d][j]
            if (tid < 19) k_th[tid] = bnd[kl_index(i, tid)];
            __syncthreads();
            const double a = k_th[0], r = k_th[1], s2 = k_th[2];
            const double* c = k_th + 3;
            const double* var = k_th + 7;
            const double* k = k_th + 11;
            if (tid < 32) {
                const int dd = tid >> 2, j = tid & 3;
                double t = 0.0;
                for (int q = 0; q < 4; ++q) t += P[dd * 16 + j * 4 + q] * (mu[dd * 4 + q] - c[q]);
                k_Pd[tid] = t;
                k_dDv[tid] = 0.5 * (P[dd * 16 + j * 5] - 1.0 / var[j]);
            }
            __syncthreads();
            if (tid < 8) {
                double t = -4.0 + pr[PR_LOGDET + i * 8 + tid];
                for (int j = 0; j < 4; ++j)
                    t += P[tid * 16 + j * 5] * var[j] + (mu[tid * 4 + j] - c[j]) * k_Pd[tid * 4 + j] - log(var[j]);
                k_D[tid] = 0.5 * t;
                k_lk[tid] = log(k[tid]) - pr[PR_LOG_K + i * 8 + tid];
            }
            __syncthreads();
            const double M = pr[PR_FLUX_MEAN + i], Vf = pr[PR_FLUX_VAR + i];
            if (tid == 0) {
                const double R = 0.5 * (log(Vf) - log(s2) + (s2 + (r - M) * (r - M)) / Vf - 1.0);
                double G = R;
                for (int dd = 0; dd < 8; ++dd) G += k[dd] * (k_lk[dd] + k_D[dd]);
                const double la = log(a) - pr[PR_LOG_A + i];
                k_sc[0] = la;
                k_sc[1] = G;
                k_Gg[0] = (r - M) / Vf;
                k_Gg[1] = 0.5 * (1.0 / Vf - 1.0 / s2);
                sc[0] -= a * (la + G);
            } else if (tid >= 1 && tid <= 4) {
                const int j = tid - 1;
                double t = 0.0;
                for (int dd = 0; dd < 8; ++dd) t -= k[dd] * k_Pd[dd * 4 + j];
                k_Gg[2 + j] = t;
            } else if (tid >= 5 && tid <= 8) {
                const int j = tid - 5;
                double t = 0.0;
                for (int dd = 0; dd < 8; ++dd) t += k[dd] * k_dDv[dd * 4 + j];
                k_Gg[6 + j] = t;
            } else if (tid >= 9 && tid <= 16) {
                const int dd = tid - 9;
                k_Gg[10 + dd] = k_lk[dd] + 1.0 + k_D[dd];
            }
            __syncthreads();
            if (tid < 19) gb[kl_index(i, tid)] -= (tid == 0 ? k_sc[0] + 1.0 + k_sc[1] : a * k_Gg[tid - 1]);
            double ksum = 0.0;
            for (int dd = 0; dd < 8; ++dd) ksum += k[dd];
            for (int e = tid; e < 19 * 19; e += TR_THREADS) {
                const int p1 = e / 19, q1 = e % 19;
                double hl;
                if (p1 == 0 && q1 == 0) {
                    hl = 1.0 / a;
                } else if (p1 == 0 || q1 == 0) {
                    hl = k_Gg[(p1 == 0 ? q1 : p1) - 1];
                } else {
                    int p = p1 - 1, q = q1 - 1;
                    if (p > q) {
                        const int t = p;
                        p = q;
                        q = t;
                    }
                    double gh = 0.0;            // d2 G / d(local p) d(local q), p <= q
                    if (q < 2) {
                        gh = (p == q) ? (p == 0 ? 1.0 / Vf : 0.5 / (s2 * s2)) : 0.0;
                    } else if (q < 6) {
                        if (p >= 2)
                            for (int dd = 0; dd < 8; ++dd) gh += k[dd] * P[dd * 16 + (p - 2) * 4 + (q - 2)];
                    } else if (q < 10) {
                        if (p == q) gh = ksum * 0.5 / (var[q - 6] * var[q - 6]);
                    } else {
                        const int dd = q - 10;
                        if (p == q)
                            gh = 1.0 / k[dd];
                        else if (p >= 2 && p < 6)
                            gh = -k_Pd[dd * 4 + (p - 2)];
                        else if (p >= 6 && p < 10)
                            gh = k_dDv[dd * 4 + (p - 6)];
                    }
                    hl = a * gh;
                }
                Hb[kl_index(i, p1) * TR_LD + kl_index(i, q1)] -= hl;
            }
            __syncthreads();
        }
        if (tid == 0) {                                 // source_e_log_prob (:132): the radius prior
            const double rm = pr[PR_RAD], rv = pr[PR_RAD + 1], x = bnd[5];
            sc[0] += -0.5 * (log(2.0 * 3.14159265358979323846) + log(rv) + (x - rm) * (x - rm) / rv);
            gb[5] += -(x - rm) / rv;
            Hb[5 * TR_LD + 5] += -1.0 / rv;
        }
        __syncthreads();
    }

    // ---- 2. propagate_derivatives!: Jacobian pieces at xe
    {
        double* scratch = sbuf;                         // to_bound_block's `out` (44 <= TR_MAXN); values not needed here
        to_bound_block(xe, lo, hi, sig, pp, scratch);
        __syncthreads();
        if (tid < NW_BOX) {
            const double sg = sig[tid];
            d1[tid] = sg * (1.0 - sg) * (hi[tid] - lo[tid]);
            d2[tid] = d1[tid] * (1.0 - 2.0 * sg);
        }
        __syncthreads();
    }
    double* T = S.V;                                    // T = Hb J, 44 x 41
    for (int e = tid; e < NW_BOUND * NW_FREE; e += TR_THREADS) {
        const int i = e / NW_FREE, a = e % NW_FREE;
        double t;
        if (a < NW_BOX) {
            t = Hb[i * TR_LD + a] * d1[a];
        } else {
            const int s = simplex_of_free(a), n = simplex_n(s), first = simplex_first(s), ja = a - simplex_free0(s);
            const double alpha = 1.0 - n * simplex_lower(s);
            const double* p = pp + simplex_pp0(s);
            t = 0.0;
            for (int j = 0; j < n; ++j) t += Hb[i * TR_LD + first + j] * simplex_jac(p, alpha, j, ja);
        }
        T[i * TR_LD + a] = t;
    }
    if (tid < NW_FREE) {                                // g_free = J' g
        const int a = tid;
        double t;
        if (a < NW_BOX) {
            t = d1[a] * gb[a];
        } else {
            const int s = simplex_of_free(a), n = simplex_n(s), first = simplex_first(s), ja = a - simplex_free0(s);
            const double alpha = 1.0 - n * simplex_lower(s);
            const double* p = pp + simplex_pp0(s);
            t = 0.0;
            for (int j = 0; j < n; ++j) t += simplex_jac(p, alpha, j, ja) * gb[first + j];
        }
        gnew[a] = -t;
    }
    if (tid >= NW_FREE && tid < TR_MAXN) gnew[tid] = 0.0;
    __syncthreads();
    double* Hf = S.A;                                   // J' T + C  (Hb is dead)
    constexpr int NP = (NW_FREE + 1) & ~1;
    for (int e = tid; e < NP * NP; e += TR_THREADS) {
        const int a = e / NP, c = e % NP;
        double t = 0.0;
        if (a < NW_FREE && c < NW_FREE) {
            if (a < NW_BOX) {
                t = d1[a] * T[a * TR_LD + c];
                if (a == c) t += gb[a] * d2[a];
            } else {
                const int s = simplex_of_free(a), n = simplex_n(s), first = simplex_first(s), f0 = simplex_free0(s);
                const int ja = a - f0, m = n - 1;
                const double alpha = 1.0 - n * simplex_lower(s);
                const double* p = pp + simplex_pp0(s);
                for (int j = 0; j < n; ++j) t += simplex_jac(p, alpha, j, ja) * T[(first + j) * TR_LD + c];
                if (c >= f0 && c < f0 + m) {
                    // sum_i g_i d2 p_i / dz_ja dz_jc = sum_i gi p_i (d_ija - p_ja)(d_ijc - p_jc) - S p_ja (d_jajc - p_jc)
                    const int jc = c - f0;
                    double t1 = 0.0, ssum = 0.0;
                    for (int i = 0; i < n; ++i) {
                        const double gi = gb[first + i] * alpha * p[i];
                        t1 += gi * ((i == ja ? 1.0 : 0.0) - p[ja]) * ((i == jc ? 1.0 : 0.0) - p[jc]);
                        ssum += gi;
                    }
                    t += t1 - ssum * ((ja == jc ? p[ja] : 0.0) - p[ja] * p[jc]);
                }
            }
        }
        // T lives in S.V, Hf overwrites S.A: no hazard.  Stored negated (minimisation form).
        Hf[a * TR_LD + c] = -t;
    }
    __syncthreads();
    for (int e = tid; e < NP * NP; e += TR_THREADS) {   // symmetrize! (:452-457)
        const int a = e / NP, c = e % NP;
        if (a < c) {
            const double hm = 0.5 * (Hf[a * TR_LD + c] + Hf[c * TR_LD + a]);
            Hf[a * TR_LD + c] = hm;
            Hf[c * TR_LD + a] = hm;
        }
    }
    __syncthreads();

    // ---- 3. Optim.NewtonTrustRegion bookkeeping (thread 0)
    if (tid == 0) {
        const double f_new = -(nb.v[b] + sc[0]);
        double gmax = 0.0;
        bool gnan = false;
        for (int a = 0; a < NW_FREE; ++a) {
            gmax = fmax(gmax, fabs(gnew[a]));
            gnan = gnan || isnan(gnew[a]);
        }
        int accept, still;
        double delta;
        if (phase == 0) {
            const bool act = !bad && !gnan && gmax >= NW_G_TOL;       // initial g_tol check
            accept = 1;
            still = act ? 1 : 0;
            delta = NW_INITIAL_DELTA;
            nb.converged[b] = (!bad && !act) ? 1 : 0;
            nb.iters[b] = 0;
            nb.f_calls[b] = 1;
        } else {
            const double f_old = nb.f[b], m = nb.m_pred[b];
            delta = nb.delta[b];
            const double eps = 2.220446049250313e-16;
            double rho = fabs(m) <= eps ? 1.0 : (m > 0.0 ? NW_RHO_LOWER - 1.0 : (f_old - f_new) / (-m));
            if (bad || !isfinite(f_new) || isnan(rho)) rho = NW_RHO_LOWER - 1.0;
            if (rho < NW_RHO_LOWER)
                delta *= 0.25;
            else if (rho > NW_RHO_UPPER && !nb.interior[b])
                delta = fmin(2.0 * delta, NW_DELTA_HAT);
            accept = rho > NW_ETA ? 1 : 0;
            // convergence is assessed only on accepted steps
            double dx = 0.0;
            const double* xo = nb.x + (size_t)b * NW_FREE;
            for (int a = 0; a < NW_FREE; ++a) dx = fmax(dx, fabs(xe[a] - xo[a]));
            const bool x_conv = dx < NW_X_TOL;
            const bool f_conv = fabs(f_new - f_old) <= NW_F_TOL * fabs(f_new);
            const bool g_conv = gmax < NW_G_TOL;
            const bool newly = accept && (x_conv || f_conv || g_conv);
            nb.iters[b] += 1;
            nb.f_calls[b] += 1;                         // the candidate's evaluation was consumed
            if (newly) nb.converged[b] = 1;
            const bool dead = delta < 1e-14;            // a collapsed trust region cannot make progress
            still = (!newly && !dead) ? 1 : 0;
        }
        nb.delta[b] = delta;
        if (accept) nb.f[b] = f_new;
        nb.active[b] = (unsigned char)still;
        flg[0] = accept;
        flg[1] = still;
        sc[1] = delta;
    }
    __syncthreads();
    const bool accept = flg[0] != 0, still = flg[1] != 0;
    double* xs = nb.x + (size_t)b * NW_FREE;
    double* gs = nb.g + (size_t)b * NW_FREE;
    double* Hs = nb.H + (size_t)b * NW_FREE * NW_FREE;
    if (accept) {
        if (tid < NW_FREE) {
            if (phase != 0) xs[tid] = xe[tid];
            gs[tid] = gnew[tid];
        }
        for (int e = tid; e < NW_FREE * NW_FREE; e += TR_THREADS) Hs[e] = Hf[(e / NW_FREE) * TR_LD + (e % NW_FREE)];
        if (tid < TR_MAXN) S.gsh[tid] = gnew[tid];
    } else if (still) {
        for (int e = tid; e < NP * NP; e += TR_THREADS) {
            const int a = e / NP, c = e % NP;
            S.A[a * TR_LD + c] = (a < NW_FREE && c < NW_FREE) ? Hs[a * NW_FREE + c] : 0.0;
        }
        if (tid < TR_MAXN) S.gsh[tid] = tid < NW_FREE ? gs[tid] : 0.0;
        if (tid < NW_FREE) xe[tid] = xs[tid];           // the candidate is abandoned: step again from x
    }
    if (!still) return;
    __syncthreads();

    // ---- 4. next candidate: exact subproblem, x + s, to_bound! -> the plan's parameter array
    tr_solve_block(S, NW_FREE, sc[1], sbuf, &sc[2], &flg[2]);
    __syncthreads();
    if (tid < NW_FREE) {
        xcand[tid] = xe[tid] + sbuf[tid];
        nb.x_new[(size_t)b * NW_FREE + tid] = xcand[tid];
    }
    if (tid == 0) {
        nb.m_pred[b] = sc[2];
        nb.interior[b] = flg[2];
    }
    __syncthreads();
    to_bound_block(xcand, lo, hi, sig, pp, bnd);
    __syncthreads();
    if (tid < NW_BOUND) vp_own[tid] = bnd[tid];
}

}  // namespace celeste
#endif
