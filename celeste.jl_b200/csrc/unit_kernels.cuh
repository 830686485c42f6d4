// unit_kernels.cuh -- the ELBO hot loop (add_pixel_term!, elbo_objective.jl:330-392) with ONE WARP per "unit" (rows
// [h2_lo, h2_hi) of one (active source, image)), pulled from a device-side queue (heaviest first): every mode of the
// production shape (Sa = 1, K = 2).  The Hessian mode (ElboMaximize.jl:150-152 always asks for the Hessian) is the
// one the formulation below is about; modes 0 / 1 are its phase A alone.
//
// The reference -- and pixel_kernel<2> -- push every one of the 28 Gaussian components of every pixel through the
// second-order chain rule (27 sums, ~96 FP64 per component-pixel).  Here the second-order work is split by what is
// LINEAR in the components and what is not:
//
//   H_(c,y) = sum_pix [ Jz' Lzz Jz  +  L4 d2f0  ]        (needs only the 7 first-order mixture sums per pixel)
//           + sum_pix   L5(pix) d2f1(pix)                 (linear in the components)
//
//  phase A  walks the rows of the patch exactly like march_kernel (a lane PAIR per walk, exp recurrence along the
//           row, 17 FP64 per component-pixel for f1 and its 6 first derivatives), finishes the pixel term with its
//           full second-order part EXCEPT L5 d2f1, and stores L5 = dL/df1 of every pixel in the unit's plane of plan.l5.
//  phase B  re-walks the patch with one LANE PER COMPONENT: f_c(pix) by the same recurrence (2 multiplications),
//           v = L5(pix) f_c(pix), and the five row moments sum v d2^b; at the end of a row they are folded into the
//           15 moments D[a][b] = sum L5 f_c d1^a d2^b (a + b <= 4, d = x - mu_c).  ~12 FP64 per component-pixel.
//  fold     every second-order mixture sum of gal_group<2> (bvn_xsig_h / bvn_sigsig_h, BivariateNormals.jl:293-316)
//           is f_c times a polynomial of degree <= 4 in Lambda_c d: its L5-weighted pixel sum is a fixed linear
//           combination of the component's 15 moments (hess_moments.inc, generated with sympy by
//           tools/gen_hess_moments.py), evaluated ONCE per component and reduced over the warp by shuffles.
//
// About 1.2 k FP64 per pixel instead of 2.9 k, mathematically the same function (floating-point reassociation
// plus the recurrence's rounding, ~1e-13 relative).
//
// Output: one NAcc<MODE>-vector of (c, y)-space sums per (sub, image) in plan.partials -- the layout
// epilogue_kernel<MODE> consumes (chunk_ptr = identity), so the raw -> parameter chain rule is unchanged.
//
// Three kernels on one stream, each a persistent grid whose warps pull units from a device-side queue and never
// synchronise with each other:
//   unit_bg_kernel      neighbours (value only) of the units that have any -> (E_bg, V_bg) planes in plan.bg
//   unit_walk_kernel<M> phase A (MODE 0 / 1: the whole value / gradient evaluation)
//   unit_moment_kernel  phase B + fold (MODE 2 only)
// A fused single kernel was measured first (profiles/ncu_unit_kernel_hess_r02_v1.txt): 10 k SASS instructions with the
// twelve warps of an SM spread over three different loops -- 31 % of its warp samples were instruction-cache misses.
#ifndef CELESTE_UNIT_KERNELS_CUH
#define CELESTE_UNIT_KERNELS_CUH

#include "march_kernels.cuh"

namespace celeste {

#ifndef CELESTE_UNIT_MINB
#define CELESTE_UNIT_MINB 3
#endif
#ifndef CELESTE_UNIT_MINB_GRAD
#define CELESTE_UNIT_MINB_GRAD CELESTE_UNIT_MINB      // value / gradient instantiations (few shared-memory accumulators)
#endif
#ifndef CELESTE_UNIT_ROWS
#define CELESTE_UNIT_ROWS 16          // rows of a patch per unit (build_unit_list)
#endif
#ifndef CELESTE_UNIT_BG_MINB
#define CELESTE_UNIT_BG_MINB 4
#endif
#ifndef CELESTE_UNIT_MOM_MINB
#define CELESTE_UNIT_MOM_MINB 6
#endif
constexpr int UNIT_WARPS = 4;
// unit_bg_kernel: shared pixels (summed over neighbours) per piece, about -- large pieces on a plan that fills the GPU
// (fewer per-neighbour prologues), small ones on a small plan (balance)
constexpr int UNIT_BG_PIXELS_BIG = 1400, UNIT_BG_PIXELS_SMALL = 350;
constexpr int MOMENT_ROWBLOCK = 16;    // unit_moment_kernel: rows between exact starts of the row-direction recurrence
constexpr int UNIT_THREADS = 32 * UNIT_WARPS;
// per-(source, image) constants of the warp (shared memory): march's SI_* plus the second-derivative spline weights
constexpr int SU_DDWX = 28, SU_DDWY = 32, SU_STRIDE = 36;
// compact per-thread accumulator slots (shared memory, stride 32): G 6 | C1 4 | HH 21 | CC 7 live | CR 24.
// ACC index (elbo_math.cuh) = slot + 3 for slot < 38, slot + 6 for the CR block; ACC_CC + 7..9 are identically
// zero (L is linear in V, so d2L/dB dB = 0).
constexpr int UA_G = 0, UA_C1 = 6, UA_HH = 10, UA_CC = 31, UA_CR = 38;
template <int MODE> struct NUAcc { static constexpr int value = MODE == 0 ? 0 : (MODE == 1 ? 10 : 62); };
CEL_HD constexpr int unit_acc_index(int slot) { return slot < UA_CR ? slot + 3 : slot + 6; }

struct UnitHdr {        // one unit = rows [h2_lo, h2_hi) of one (sub, image); built on the host (build_unit_list), heaviest first
    int aslot, slot0, slot1;
    int field, sub, task;
    int n;              // image
    int nseg;           // column segments per row of the active patch (phase A)
    int hasbg;          // some other source of the task reaches this image
    int pidx;           // partial vector of this unit in plan.partials
    int nbpix;          // pixels shared with neighbours (cost of the (sub, image) in unit_bg_kernel)
    int tn;             // sub * N + n: index of the (sub, image) planes (plan.pix / plan.l5 / plan.bg)
    int h2_lo, h2_hi;   // rows of the active patch this unit walks (a large plan: all of them; a small plan is cut
                        // finer so that one source's evaluation spreads over many SMs)
    int first;          // 1: the unit that holds row 0 (it carries the (sub, image)'s neighbour counter)
    int bgp0, bgp1;     // first unit: the (sub, image)'s pieces in unit_bg_kernel's list (their counters: plan.bg_cnt);
                        // a piece of that list: bgp0 = its own slot
    int pad;
};

// One active pixel as the walk reads it: 16 bytes, stored per unit in WALK ORDER (row-major inside the patch: a row
// walk reads consecutive records), packed once per plan by unit_pack_kernel.  x = NaN marks a pixel that is masked
// (NaN in the image, elbo_objective.jl:459) or not in the active bitmap (:445).
struct PixRec {
    float x, sky;
    double pixconst;      // x log(iota) - lgamma(x + 1)
};

// geometry and pointers of the unit's image / active patch (per-warp shared memory)
struct UnitImg {
    const PixRec* pix;      // H2 x W2 records, row-major
    const float* iota;
    const double* coefs;
    const double* bg;       // (E_bg, V_bg) pairs of the unit, row-major like pix, or null
    int H2, W2, off_h, off_w, n1, n2, band0, pad;
};

// Kernel-tuning knob, OFF by default: value / gradient instantiations have shared memory to spare, so the window of the
// spline coefficient table that a unit can touch ((rows + 3) x (columns + 3) doubles) can be staged there
// (-DCELESTE_UNIT_WIN=1600).  Measured: unit_walk_kernel<1> 1.97 -> 2.19 ms -- the 12.8 KB per warp come out of L1, and
// the pixel records then miss more than the taps gain (profiles/tuning_r02.md).
#ifndef CELESTE_UNIT_WIN
#define CELESTE_UNIT_WIN 0
#endif
template <int MODE> struct UnitWin { static constexpr int value = MODE == 2 ? 0 : CELESTE_UNIT_WIN; };
// How the walk gets its pixel records (profiles/tuning_r02.md):
//   CELESTE_UNIT_CPASYNC = 2 (default)  value / gradient instantiations: each lane streams its own 16-byte records
//                              global -> shared with cp.async (LDGSTS), three iterations deep (1.5 KB ring per warp);
//                              the Hessian instantiation, whose shared memory is full of accumulators, keeps
//                              prefetch.global.L1 + a plain load.  Measured: unit_walk_kernel<1> 1.963 -> 1.936 ms
//                        = 1  every mode (unit_walk_kernel<2>: 2.93 -> 3.26 ms: the ring costs it L1)
//                        = 0  never
//   CELESTE_UNIT_PF_AHEAD = k  the spline taps are prefetched k iterations further ahead (measured: slower, off)
//   CELESTE_UNIT_BG_PF = 1 (default)  unit_bg_kernel prefetches the next iteration's bitmap byte / record / sums / taps
//                              (0.289 -> 0.284 ms)
#ifndef CELESTE_UNIT_CPASYNC
#define CELESTE_UNIT_CPASYNC 2
#endif
#ifndef CELESTE_UNIT_BG_PF
#define CELESTE_UNIT_BG_PF 1
#endif
#ifndef CELESTE_UNIT_PF_AHEAD
#define CELESTE_UNIT_PF_AHEAD 0
#endif
constexpr int UNIT_RING = 3;                                     // cp.async ring slots per lane
template <int MODE> struct UnitCpAsync {
    static constexpr bool value = CELESTE_UNIT_CPASYNC == 1 || (CELESTE_UNIT_CPASYNC == 2 && MODE <= 1);
};
template <int MODE> struct UnitWarpDoubles {
    static constexpr size_t value = (size_t)NUAcc<MODE>::value * 32 + (size_t)NC2 * MREC + SU_STRIDE + (sizeof(UnitImg) + 7) / 8 +
                                    (size_t)UnitWin<MODE>::value + (UnitCpAsync<MODE>::value ? UNIT_RING * 32 * 2 : 0);
};
template <int MODE>
constexpr size_t unit_smem_bytes() { return UnitWarpDoubles<MODE>::value * UNIT_WARPS * sizeof(double); }
constexpr size_t unit_bg_smem_bytes() { return (size_t)(NC2 * MREC + SU_STRIDE) * UNIT_WARPS * sizeof(double); }
constexpr size_t unit_moment_smem_bytes() { return (size_t)(NC2 * MREC) * UNIT_WARPS * sizeof(double); }

// per-(source, image) constants of a walk (march's SI_* layout plus the second-derivative spline weights), staged by
// FOUR lanes in parallel (which = 0..3): the brightness scalars of the image's band from the slot's precomputed
// moments (slotbr_kernel; a_i E_l[i][b] in the summation order of brightness_values, as march_stage_srcimg computes
// them), m_pos (linear_world_to_pix, wcs_utils.jl:14-18), the spline weights at the patch's fractional offsets.
template <int MODE>
__device__ inline void unit_stage_srcimg(const PatchDev& p, const double* vs, const double* br, int band0, double* si,
                                         int which) {
    if (which == 0) {
        si[SI_CB + 0] = br[20] * br[band0];
        si[SI_CB + 1] = br[21] * br[5 + band0];
        si[SI_CB + 2] = br[20] * br[10 + band0];
        si[SI_CB + 3] = br[21] * br[15 + band0];
        si[SI_THETA] = vs[2];
        si[SI_THETA + 1] = 0.0;
    } else if (which == 1) {
        double m1, m2;
        march_m_pos(p, vs, m1, m2);
        si[SI_M] = m1;
        si[SI_M + 1] = m2;
        for (int i = 0; i < 4; ++i) si[SI_J + i] = p.J[i];
    } else {
        double m1, m2;
        march_m_pos(p, vs, m1, m2);
        const double a = which == 2 ? (double)(p.off_h + 1) - m1 + 26.0 : (double)(p.off_w + 1) - m2 + 26.0;
        double w[4], dw[4], ddw[4];
        cubic_weights<MODE>(a - floor(a), w, dw, ddw);
        const int o = which == 2 ? 0 : SI_WY - SI_WX;
        for (int i = 0; i < 4; ++i) {
            si[SI_WX + o + i] = w[i];
            if (MODE >= 1) si[SI_DWX + o + i] = dw[i];
            if (MODE >= 2) si[(which == 2 ? SU_DDWX : SU_DDWY) + i] = ddw[i];
        }
    }
}

// The pixel term (add_elbo_log_term! + add_scaled_sfs!, elbo_objective.jl:274-327, :383-391) of one covered or
// uncovered active pixel: the same expressions as pixel_accumulate<MODE> (elbo_math.cuh) without the
// L5 * d2f1 part of the Hessian, which phase B supplies.  acc: this thread's slots (stride 32).  Returns L5.
template <int MODE>
__device__ __forceinline__ double unit_pixel_term(double* acc, const PixelConsts& pc, double Ebg, double Vbg, bool covered,
                                                  const double* cb, double f0, const double* g0, const double* h0, double f1,
                                                  const double* r, double& val, const double* logtab) {
    const double A1 = cb[0], A2 = cb[1], B1 = cb[2], B2 = cb[3];
    const double m = covered ? (A1 * f0 + A2 * f1) : 0.0;
    const double E = Ebg + m;
    const double V = covered ? (Vbg + B1 * f0 * f0 + B2 * f1 * f1 - m * m) : Vbg;
    const double iE = 1.0 / E;
    const double iE2 = iE * iE;
    val += pc.x * (log_tab(E, logtab) - 0.5 * V * iE2) - pc.iota * E + pc.pixconst;
    if (MODE == 0 || !covered) return 0.0;
    constexpr int S = 32;
    const double gE = pc.x * (iE + V * iE2 * iE) - pc.iota;
    const double gV = -0.5 * pc.x * iE2;
    const double Ez[6] = {f0, f1, 0.0, 0.0, A1, A2};
    const double Vz[6] = {-2.0 * m * f0, -2.0 * m * f1, f0 * f0, f1 * f1, 2.0 * (B1 * f0 - m * A1), 2.0 * (B2 * f1 - m * A2)};
    double Lz[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) Lz[i] = gE * Ez[i] + gV * Vz[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double g = Lz[5] * r[k];
        if (k < 2) g += Lz[4] * g0[k];
        acc[(UA_G + k) * S] += g;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[(UA_C1 + c) * S] += Lz[c];
    if (MODE == 1) return 0.0;

    // Second order.  With e = dE/dz, v = dV/dz (z = A1 A2 B1 B2 f0 f1) the Hessian of the pixel term in z is
    //     Lzz = e b' + b e' + S,     b = (LEE / 2 - gV) e + LEV v,
    // a symmetric rank-2 matrix plus the six sparse entries S = (gE - 2 gV m) Ezz + gV Bpart (L is linear in V, so there
    // is no v v' term; combine_sfs_hessian!, SensitiveFloats.jl:99-126, written out).  Pushed through
    // Jz = d z / d(c, y) -- identity on c, g0 (x only) for f0, r for f1 -- the dense part stays rank 2:
    //     H_(c,y) = et bt' + bt et' + Jz' S Jz,   et = Jz' e,   bt = Jz' b.
    const double LEE = -pc.x * (iE2 + 3.0 * V * iE2 * iE2);
    const double LEV = pc.x * iE2 * iE;
    const double ha = 0.5 * LEE - gV;
    double b[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = (i == 2 || i == 3) ? LEV * Vz[i] : fma(ha, Ez[i], LEV * Vz[i]);
    // et_y = A2 r + A1 g0 and bt_y = b5 r + b4 g0 (g0 in the x rows only), so the y-y block is
    //     HH = al r r' + be (g0 r' + r g0') + ga g0 g0' + L4 h0,
    // and row c of the c-y block is  cr_c r' + cg_c g0'  -- scalars first, then one FMA per entry and vector.
    const double sig = gE - 2.0 * gV * m;           // Ezz entries (A1, f0), (A2, f1)
    const double al = 2.0 * (A2 * b[5] + gV * B2);
    const double be = fma(A1, b[5], A2 * b[4]);
    const double ga = 2.0 * (A1 * b[4] + gV * B1);
    double ar[6], bg0[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) ar[k] = al * r[k];
    bg0[0] = be * g0[0];
    bg0[1] = be * g0[1];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int l = k; l < 6; ++l) {
            double a = acc[(UA_HH + tri6(k, l)) * S];
            a = fma(ar[k], r[l], a);
            if (k < 2) a = fma(bg0[k], r[l], a);
            if (l < 2) {
                a = fma(r[k], bg0[l], a);
                a = fma(ga * g0[k], g0[l], a);
                a = fma(Lz[4], h0[k + l], a);              // h0 packed xx, xy, yy
            }
            acc[(UA_HH + tri6(k, l)) * S] = a;
        }
    acc[(UA_CC + tri4(0, 0)) * S] += 2.0 * f0 * b[0];
    acc[(UA_CC + tri4(0, 1)) * S] += fma(f0, b[1], f1 * b[0]);
    acc[(UA_CC + tri4(0, 2)) * S] += f0 * b[2];
    acc[(UA_CC + tri4(0, 3)) * S] += f0 * b[3];
    acc[(UA_CC + tri4(1, 1)) * S] += 2.0 * f1 * b[1];
    acc[(UA_CC + tri4(1, 2)) * S] += f1 * b[2];
    acc[(UA_CC + tri4(1, 3)) * S] += f1 * b[3];
    // c-y block: e = (f0, f1, 0, 0) on the c rows
    const double cr[4] = {fma(f0, b[5], b[0] * A2), fma(f1, b[5], fma(b[1], A2, sig)), b[2] * A2, fma(b[3], A2, 2.0 * gV * f1)};
    const double cg[4] = {fma(f0, b[4], fma(b[0], A1, sig)), fma(f1, b[4], b[1] * A1), fma(b[2], A1, 2.0 * gV * f0), b[3] * A1};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double a = acc[(UA_CR + c * 6 + k) * S];
            a = fma(cr[c], r[k], a);
            if (k < 2) a = fma(cg[c], g0[k], a);
            acc[(UA_CR + c * 6 + k) * S] = a;
        }
    return Lz[5];
}

// One component's share of the 20 second-order mixture sums from its 15 L5-weighted moments (see the file header)
__device__ __forceinline__ void unit_moments_to_sums(double l11, double l12, double l22, const double* D, double* OUT) {
    const double D00 = D[0], D01 = D[1], D02 = D[2], D03 = D[3], D04 = D[4];
    const double D10 = D[5], D11 = D[6], D12 = D[7], D13 = D[8];
    const double D20 = D[9], D21 = D[10], D22 = D[11];
    const double D30 = D[12], D31 = D[13];
    const double D40 = D[14];
#include "hess_moments.inc"
}

// Once per plan: the pixel records of every unit.  One block per unit.
__global__ void unit_pack_kernel(PlanDev plan, const UnitHdr* __restrict__ units, int n_units, PixRec* __restrict__ out) {
    const int u = blockIdx.x;
    if (u >= n_units) return;
    const UnitHdr uh = units[u];
    if (!uh.first) return;
    const FieldDev field = plan.fields[uh.field];
    const ImageDev img = field.images[uh.n];
    const PatchDev& pa = field.patches[plan.src_row[uh.aslot] + (size_t)uh.n * field.S_tot];
    const int H2 = pa.H2, W2 = pa.W2;
    PixRec* dst = out + plan.l5_ptr[uh.tn];
    for (int i = threadIdx.x; i < H2 * W2; i += blockDim.x) {
        const int h2 = i / W2, w2 = i - h2 * W2;
        const size_t ipix = (size_t)(pa.off_h + h2) + (size_t)(pa.off_w + w2) * img.H;
        PixRec r;
        r.x = pa.bitmap[h2 + (size_t)w2 * H2] ? img.pixels[ipix] : nanf("");
        r.sky = img.sky[ipix];
        r.pixconst = img.pixconst[ipix];
        dst[i] = r;
    }
}

// Neighbouring sources (value only, elbo_objective.jl:38-40,69): E_bg += E_s, V_bg += E2_s - E_s^2 over the pixels a
// neighbour shares with the active patch, one neighbour at a time in slot order (fixed summation order), into the
// unit's (E_bg, V_bg) planes in plan.bg.  One warp per unit that has a neighbour; the walk is march_kernel's.
// MINB: blocks per SM the registers are budgeted for -- 4 (128 registers) is faster on a plan that fills the GPU, 3 (168,
// no spills) on a small, latency-bound plan (profiles/tuning_r02.md).
template <int MINB>
__global__ void __launch_bounds__(UNIT_THREADS, MINB)
    unit_bg_kernel(PlanDev plan, const UnitHdr* __restrict__ units, int n_units, int* __restrict__ queue,
                   const double* __restrict__ vp) {
    CEL_DYNAMIC_SMEM(smem);
    __shared__ double s_exptab[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kk = tid & 1;
    (void)kk;
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    double* wbase = smem + (size_t)warp * (NC2 * MREC + SU_STRIDE);
    double* s_rec = wbase;
    double* s_si = s_rec + NC2 * MREC;
    __syncthreads();
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const UnitHdr uh = units[u];
        if (plan.task_mask && !plan.task_mask[uh.task]) continue;
        const FieldDev field = plan.fields[uh.field];
        const int n = uh.n, aslot = uh.aslot;
        const PatchDev* prow = field.patches + (size_t)n * field.S_tot;       // patches of image n, by source row
        const PatchDev& pa = prow[plan.src_row[aslot]];
        const int band0 = field.images[n].band - 1;
        if (!uh.hasbg) continue;
        double* my_scratch = plan.bg + plan.bg_ptr[uh.tn];
        const PixRec* arec = plan.pix + plan.l5_ptr[uh.tn];      // x = NaN: masked or not in the active bitmap
        double cnt_inactive = 0.0;
        {
            {
                const int i0 = 2 * uh.h2_lo * pa.W2, i1 = 2 * uh.h2_hi * pa.W2;  // this piece's rows of the planes
                for (int i = i0 + lane; i < i1; i += 32) my_scratch[i] = 0.0;   // (E_bg, V_bg) pairs, row-major
            }
            for (int s = uh.slot0; s < uh.slot1; ++s) {
                if (s == aslot) continue;
                const PatchDev& p = prow[plan.src_row[s]];
                // active pixels: rows off+1..off+H2, columns off+1..off+W2; the neighbour covers columns
                // off+1..off+W2-1 only (strict `w2 < W2`, elbo_objective.jl:349)
                // ... restricted to the rows h2_lo .. h2_hi - 1 of the active patch that this piece owns
                const int h_lo = max(pa.off_h + uh.h2_lo, p.off_h) + 1, h_hi = min(pa.off_h + uh.h2_hi, p.off_h + p.H2);
                const int w_lo = max(pa.off_w, p.off_w) + 1, w_hi = min(pa.off_w + pa.W2, p.off_w + p.W2 - 1);
                if (h_hi < h_lo || w_hi < w_lo) continue;            // warp-uniform
                const int bnh = h_hi - h_lo + 1, bnw = w_hi - w_lo + 1;
                __syncwarp();                                        // the previous neighbour's walks are done with s_rec / s_si
                if (lane < NC2) {
                    const double* vs = vp + (size_t)NPARAM * s;
                    double xx[3];
                    galaxy_xixi(vs[3], vs[4], vs[5], xx[0], xx[1], xx[2]);
                    march_make_record(p, vs, xx, s_rec, lane);
                }
                if (lane >= NC2) unit_stage_srcimg<0>(p, vp + (size_t)NPARAM * s, plan.slotbr + (size_t)s * SLOTBR_STRIDE, band0, s_si, lane - NC2);
                __syncwarp();
                const int nsg = (bnw + MARCH_MAXSEG - 1) / MARCH_MAXSEG;
                const int total = bnh * nsg;
                const int segw = (bnw + nsg - 1) / nsg;
                for (int ub = 0; ub < total; ub += NPW) {            // warp-uniform
                    const int uu = ub + (lane >> 1);
                    const bool has = uu < total;
                    const int ul = has ? uu : 0;
                    const int seg = ul / bnh, row = ul - seg * bnh;
                    const int c0 = seg * segw;
                    const int len = has ? max(min(segw, bnw - c0), 0) : 0;
                    const int nit = warp_max_int((len + 1) >> 1);
                    if (nit == 0) continue;
                    const int h = h_lo + row, w0 = w_lo + c0;       // 1-based image coordinates
                    const int aW2 = pa.W2, nH2 = p.H2, n1 = p.n1, n2 = p.n2;
                    const double* coefs = p.coefs;
                    const double* si = s_si;
                    const double* recs = s_rec + kk * MREC;          // this lane's PSF component
                    double fp[NPROTO], rr[NPROTO];
                    const double ax = (double)h - si[SI_M] + 26.0, ay0 = (double)w0 - si[SI_M + 1] + 26.0;
                    const int ixf = (int)floor(ax), iy0 = (int)floor(ay0);
                    const bool fast = len > 0 && ixf >= 1 && ixf <= n1 - 3 && iy0 >= 1 && iy0 + len - 1 <= n2 - 3;
                    double R0 = 0.0, R1 = 0.0;
                    const double* ccol = coefs + (size_t)(fast ? iy0 - 1 + kk : 0) * n1 + (fast ? ixf - 1 : 0);
                    if (fast && kk < len) {
                        const double wx0 = si[SI_WX], wx1 = si[SI_WX + 1], wx2 = si[SI_WX + 2], wx3 = si[SI_WX + 3];
                        R0 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                        ccol += n1;
                        R1 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                        ccol += n1;
                    }
                    const int ah2 = h - pa.off_h - 1, nh2 = h - p.off_h - 1;
                    const int acol = w0 - pa.off_w - 1 + kk, ncol = w0 - p.off_w - 1 + kk;      // own first column, 0-based
                    const uint8_t* nbit = p.bitmap + nh2 + (size_t)ncol * nH2;
                    const PixRec* px = arec + (size_t)ah2 * aW2 + acol;
                    double* bgE = my_scratch + 2 * ((size_t)ah2 * aW2 + acol);
                    const double theta = si[SI_THETA];
                    int t = 0;
                    while (t < nit) {
                        const bool asleep = march_start(recs, s_exptab, (double)h, (double)(w0 + 2 * t), fp, rr);
                        const bool careful = __ballot_sync(0xffffffffu, asleep && len > 0) != 0u;
                        const int tend = careful ? min(nit, t + MARCH_CAREFUL_COLS / 2) : nit;
                        for (; t < tend; ++t) {
                            const bool own = 2 * t + kk < len;
                            unsigned char nb = 0;
                            float xv = nanf("");
                            double b0 = 0.0, b1 = 0.0;            // the sums so far: requested now, used after the mixture
                            if (own) {
                                nb = *nbit;
                                xv = px->x;
                                b0 = bgE[0];
                                b1 = bgE[1];
#if CELESTE_UNIT_BG_PF
                                CEL_PREFETCH_L1(nbit + 2 * nH2);       // the next iteration's loads
                                CEL_PREFETCH_L1(px + 2);
                                CEL_PREFETCH_L1(bgE + 4);
                                if (fast) {
                                    CEL_PREFETCH_L1(ccol + 2 * n1);
                                    CEL_PREFETCH_L1(ccol + 3 * n1 + 3);
                                }
#endif
                            }
                            double R2 = 0.0, R3 = 0.0;
                            if (fast && own) {
                                const double wx0 = si[SI_WX], wx1 = si[SI_WX + 1], wx2 = si[SI_WX + 2], wx3 = si[SI_WX + 3];
                                R2 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                                ccol += n1;
                                R3 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                                ccol += n1;
                            }
                            double S[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                            for (int pix = 0; pix < 2; ++pix) {
#pragma unroll
                                for (int j = 0; j < NPROTO; ++j) {
                                    S[pix][j < NPROTO_DEV ? 0 : 1] += fp[j];
                                    fp[j] *= rr[j];
                                    rr[j] *= recs[j * 2 * MREC + 3];
                                }
                            }
                            double Fd = kk == 0 ? S[0][0] : S[1][0], Fe = kk == 0 ? S[0][1] : S[1][1];
                            Fd += __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][0] : S[0][0], 1);
                            Fe += __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][1] : S[0][1], 1);
                            if (nb && !isnan(xv)) {
                                double f0;
                                if (fast) {
                                    const double v = si[SI_WY] * R0 + si[SI_WY + 1] * R1 + si[SI_WY + 2] * R2 + si[SI_WY + 3] * R3;
                                    f0 = v < 0 ? 1e-3 * exp_nonpos(v) : 1e-3 * (v + 1.0);     // softpluslikeinv, fsm_util.jl:222
                                } else {
                                    double gd[2], hd[3];
                                    star_eval<0>(LdGlobal(), coefs, n1, n2, ax, ay0 + (double)(2 * t + kk), f0, gd, hd);
                                }
                                const double f1 = theta * Fd + (1.0 - theta) * Fe;
                                const double Es = si[SI_CB] * f0 + si[SI_CB + 1] * f1;
                                const double E2s = si[SI_CB + 2] * f0 * f0 + si[SI_CB + 3] * f1 * f1;
                                bgE[0] = b0 + Es;
                                bgE[1] = b1 + (E2s - Es * Es);
                                cnt_inactive += 1.0;                                          // elbo_objective.jl:353-357
                            }
                            nbit += 2 * nH2;
                            px += 2;
                            bgE += 4;
                            R0 = R2;
                            R1 = R3;
                        }
                    }
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) cnt_inactive += __shfl_xor_sync(0xffffffffu, cnt_inactive, o);
        if (lane == 0) plan.bg_cnt[uh.bgp0] = cnt_inactive;
        __syncwarp();
    }
}

// cp.async (LDGSTS) of one 16-byte record, for the CELESTE_UNIT_CPASYNC variant
__device__ __forceinline__ void unit_cp_async16(void* smem_dst, const void* gsrc) {
#ifndef CELESTE_HOST_EMULATION
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#else
    memcpy(smem_dst, gsrc, 16);
#endif
}
__device__ __forceinline__ void unit_cp_async_commit() {
#ifndef CELESTE_HOST_EMULATION
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N> __device__ __forceinline__ void unit_cp_async_wait() {
#ifndef CELESTE_HOST_EMULATION
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// Phase A: the active source of a unit, row walks by lane pairs (see the file header).
template <int MODE>
__global__ void __launch_bounds__(UNIT_THREADS, MODE == 2 ? CELESTE_UNIT_MINB : CELESTE_UNIT_MINB_GRAD)
    unit_walk_kernel(PlanDev plan, const UnitHdr* __restrict__ units, int n_units, int* __restrict__ queue,
                     const double* __restrict__ vp) {
    constexpr int NUA = NUAcc<MODE>::value;
    constexpr int NS = MODE == 0 ? 2 : 7;
    constexpr int NACC = NAcc<MODE>::value;
    CEL_DYNAMIC_SMEM(smem);
    __shared__ double s_exptab[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kk = tid & 1;
    (void)kk;
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    __shared__ double s_logtab[256];
    double* wbase = smem + (size_t)warp * UnitWarpDoubles<MODE>::value;
    double* acc = wbase + lane;                             // NUA x 32
    double* s_rec = wbase + (size_t)NUA * 32;               // NC2 x MREC
    double* s_si = s_rec + NC2 * MREC;                      // SU_STRIDE
    UnitImg& mi = *reinterpret_cast<UnitImg*>(s_si + SU_STRIDE);
    constexpr int WCAP = UnitWin<MODE>::value;
    double* s_win = s_si + SU_STRIDE + (sizeof(UnitImg) + 7) / 8;     // WCAP doubles
    (void)s_win;
    // ring[slot][lane], 16 bytes each; every term of the per-warp layout before it is an even number of doubles
    constexpr bool CPA = UnitCpAsync<MODE>::value;
    PixRec* ring = reinterpret_cast<PixRec*>(s_win + WCAP) + lane;
    (void)ring;
#ifdef CELESTE_HOST_EMULATION
    for (int i = tid; i < 256; i += UNIT_THREADS) s_logtab[i] = h_logtab[i];
#else
    for (int i = tid; i < 256; i += UNIT_THREADS) s_logtab[i] = g_logtab[i];
#endif
    __syncthreads();
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const UnitHdr uh = units[u];
        if (plan.task_mask && !plan.task_mask[uh.task]) continue;
        const FieldDev field = plan.fields[uh.field];
        const int n = uh.n, aslot = uh.aslot;
        const PatchDev* prow = field.patches + (size_t)n * field.S_tot;       // patches of image n, by source row
        const PatchDev& pa = prow[plan.src_row[aslot]];
        const int band0 = field.images[n].band - 1;
        __syncwarp();                                       // the previous unit's shared data is consumed
#pragma unroll
        for (int a = 0; a < NUA; ++a) acc[a * 32] = 0.0;
        if (lane == 30) {
            mi.pix = plan.pix + plan.l5_ptr[uh.tn];
            mi.iota = field.images[n].iota;
            mi.coefs = pa.coefs;
            mi.bg = uh.hasbg ? plan.bg + plan.bg_ptr[uh.tn] : nullptr;
            mi.H2 = pa.H2;
            mi.W2 = pa.W2;
            mi.off_h = pa.off_h;
            mi.off_w = pa.off_w;
            mi.n1 = pa.n1;
            mi.n2 = pa.n2;
            mi.band0 = band0;
            mi.pad = 0;
        }
        double cnt_active = 0.0, val = 0.0;
        if (lane < NC2) {
            const double* vs = vp + (size_t)NPARAM * aslot;
            double xx[3];
            galaxy_xixi(vs[3], vs[4], vs[5], xx[0], xx[1], xx[2]);      // XiXi: one sin / cos per record lane, in parallel
            march_make_record(pa, vs, xx, s_rec, lane);
        }
        if (lane >= NC2) unit_stage_srcimg<MODE>(pa, vp + (size_t)NPARAM * aslot, plan.slotbr + (size_t)aslot * SLOTBR_STRIDE, band0, s_si, lane - NC2);
        __syncwarp();

        // ---- phase A: the active source, row walks by lane pairs ------------------------------------------------
        const int H2c = max(uh.h2_hi - uh.h2_lo, 1), W2 = mi.W2;      // rows of this unit
        // window of the spline table this unit can touch: table rows wx0 .. wx0 + WR - 1, columns wy0 .. wy0 + WC - 1
        const int wx0 = (int)floor((double)(mi.off_h + uh.h2_lo + 1) - s_si[SI_M] + 26.0) - 1;
        const int wy0 = (int)floor((double)(mi.off_w + 1) - s_si[SI_M + 1] + 26.0) - 1;
        const int WR = H2c + 3, WC = W2 + 3;
        const bool winok = WCAP > 0 && WR * WC <= WCAP;
        if (winok) {
            for (int i = lane; i < WR * WC; i += 32) {
                const int c = i / WR, r = i - c * WR;
                const int tx = wx0 + r, ty = wy0 + c;
                s_win[i] = (tx >= 0 && tx < mi.n1 && ty >= 0 && ty < mi.n2) ? __ldg(mi.coefs + (size_t)ty * mi.n1 + tx) : 0.0;
            }
            __syncwarp();
        }
        const double* cbase = winok ? s_win : mi.coefs;              // the star's taps: cbase[coff + ...], column stride cstr
        const int cstr = winok ? WR : mi.n1;
        const int nseg = uh.nseg;
        const int total = (uh.h2_hi > uh.h2_lo && W2 > 0) ? (uh.h2_hi - uh.h2_lo) * nseg : 0;
        const int segw = (W2 + nseg - 1) / nseg;
        double* l5plane = MODE >= 2 ? plan.l5 + plan.l5_ptr[uh.tn] : nullptr;
        for (int ub = 0; ub < total; ub += NPW) {                      // warp-uniform
            const int uu = ub + (lane >> 1);
            const bool has = uu < total;
            const int ul = has ? uu : 0;
            const int seg = ul / H2c, h2 = uh.h2_lo + (ul - seg * H2c);
            const int c0 = seg * segw;
            const int len = has ? max(min(segw, W2 - c0), 0) : 0;
            const int nit = warp_max_int((len + 1) >> 1);
            if (nit == 0) continue;
            const int ncov = min(len, W2 - 1 - c0);                   // pixels before the (uncovered) last column, :349
            const int h = mi.off_h + h2 + 1, w0 = mi.off_w + c0 + 1;  // 1-based image coordinates
            const double* si = s_si;
            const double* recs = s_rec + kk * MREC;                   // this lane's PSF component
            double fp[NPROTO], rr[NPROTO];
            const double d1 = (double)h - recs[4];                    // x1 - mu1 of this PSF component
            double d2 = (double)w0 - recs[5];
            bool fast;
            int coff;
            {
                const int ixf = (int)floor((double)h - si[SI_M] + 26.0), iy0 = (int)floor((double)w0 - si[SI_M + 1] + 26.0);
                fast = len > 0 && ixf >= 1 && ixf <= mi.n1 - 3 && iy0 >= 1 && iy0 + len - 1 <= mi.n2 - 3;
                coff = fast ? (winok ? (iy0 - 1 + kk - wy0) * WR + (ixf - 1 - wx0) : (iy0 - 1 + kk) * mi.n1 + (ixf - 1)) : 0;
            }
            double R0 = 0.0, R1 = 0.0, D0 = 0.0, D1 = 0.0, Q0 = 0.0, Q1 = 0.0;   // row-interpolated value / d / d2 columns
            if (fast && kk < len) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const double* ccol = cbase + coff;
                    const double q0 = ccol[0], q1 = ccol[1], q2 = ccol[2], q3 = ccol[3];
                    const double r = si[SI_WX] * q0 + si[SI_WX + 1] * q1 + si[SI_WX + 2] * q2 + si[SI_WX + 3] * q3;
                    double d = 0.0, dd = 0.0;
                    if (MODE >= 1) d = si[SI_DWX] * q0 + si[SI_DWX + 1] * q1 + si[SI_DWX + 2] * q2 + si[SI_DWX + 3] * q3;
                    if (MODE >= 2) dd = si[SU_DDWX] * q0 + si[SU_DDWX + 1] * q1 + si[SU_DDWX + 2] * q2 + si[SU_DDWX + 3] * q3;
                    coff += cstr;
                    if (b == 0) {
                        R0 = r;
                        D0 = d;
                        Q0 = dd;
                    } else {
                        R1 = r;
                        D1 = d;
                        Q1 = dd;
                    }
                }
            }
            int pix = h2 * W2 + c0 + kk;                              // own pixel inside the patch (walk order)
            const double iota_h = has ? (double)mi.iota[h - 1] : 0.0; // nelec_per_nmgy of this row

            int t = 0;
            if (CPA) {
#pragma unroll
                for (int a = 0; a < UNIT_RING - 1; ++a) {              // records of iterations 0 and 1 of this row piece
                    if (2 * a + kk < len) unit_cp_async16(ring + a * 32, mi.pix + pix + 2 * a);
                    unit_cp_async_commit();
                }
            }
            while (t < nit) {
                const bool asleep = march_start(recs, s_exptab, (double)h, (double)(w0 + 2 * t), fp, rr);
                const bool careful = __ballot_sync(0xffffffffu, asleep && len > 0) != 0u;
                const int tend = careful ? min(nit, t + MARCH_CAREFUL_COLS / 2) : nit;
                for (; t < tend; ++t) {
                    const int iown = 2 * t + kk;
                    const bool own = iown < len;
                    if (CPA) {
                        if (iown + 2 * (UNIT_RING - 1) < len)
                            unit_cp_async16(ring + ((t + UNIT_RING - 1) % UNIT_RING) * 32, mi.pix + pix + 2 * (UNIT_RING - 1));
                        unit_cp_async_commit();
                    }
                    if (own) {
                        if (!CPA) CEL_PREFETCH_L1(mi.pix + pix + 2);       // the pair's next two records share a sector
                        if (fast && !winok) {
                            CEL_PREFETCH_L1(mi.coefs + coff + CELESTE_UNIT_PF_AHEAD * 2 * mi.n1);
                            CEL_PREFETCH_L1(mi.coefs + coff + CELESTE_UNIT_PF_AHEAD * 2 * mi.n1 + mi.n1 + 3);
                        }
#if CELESTE_UNIT_PF_AHEAD
                        if (mi.bg) CEL_PREFETCH_L1(mi.bg + 2 * (pix + 2));
#endif
                    }
                    // this lane's half (PSF component kk) of the mixture sums of both pixels (columns 2t and 2t + 1)
                    double S[2][NS];
#pragma unroll
                    for (int q = 0; q < NS; ++q) S[0][q] = S[1][q] = 0.0;
                    const double theta = si[SI_THETA];
#pragma unroll
                    for (int j = 0; j < NPROTO; ++j) {
                        const double* o = recs + j * 2 * MREC;
                        const double cc = o[3];
                        const double fa = fp[j];
                        const double ra = rr[j];
                        const double fb = fa * ra;               // column 2t + 1
                        const double rb = ra * cc;
                        fp[j] = fb * rb;                         // column 2t + 2
                        rr[j] = rb * cc;
                        S[0][j < NPROTO_DEV ? 0 : 1] += fa;
                        S[1][j < NPROTO_DEV ? 0 : 1] += fb;
                        if (MODE >= 1) {
                            const double l11 = o[0], l12 = o[1], l22 = o[2];
                            const double p1a = fma(l12, d2, l11 * d1), p1b = p1a + l12;
                            const double p2a = fma(l22, d2, l12 * d1), p2b = p2a + l22;
                            const double tw = j < NPROTO_DEV ? theta : 1.0 - theta;
                            const double wa = fa * tw, wb = fb * tw;
                            S[0][2] = fma(wa, p1a, S[0][2]);
                            S[1][2] = fma(wb, p1b, S[1][2]);
                            S[0][3] = fma(wa, p2a, S[0][3]);
                            S[1][3] = fma(wb, p2b, S[1][3]);
                            const double na = wa * c_proto_nu[j], nb = wb * c_proto_nu[j];
                            S[0][4] = fma(na, fma(p1a, p1a, -l11), S[0][4]);      // 2 x bvn_sig_d[1], BivariateNormals.jl:267-272
                            S[1][4] = fma(nb, fma(p1b, p1b, -l11), S[1][4]);
                            S[0][5] = fma(na, fma(p1a, p2a, -l12), S[0][5]);
                            S[1][5] = fma(nb, fma(p1b, p2b, -l12), S[1][5]);
                            S[0][6] = fma(na, fma(p2a, p2a, -l22), S[0][6]);      // 2 x bvn_sig_d[3]
                            S[1][6] = fma(nb, fma(p2b, p2b, -l22), S[1][6]);
                        }
                    }
                    d2 += 2.0;
                    double T[NS];
#pragma unroll
                    for (int q = 0; q < NS; ++q)
                        T[q] = (kk == 0 ? S[0][q] : S[1][q]) + __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][q] : S[0][q], 1);

                    float xf = nanf(""), skyf = 0.f;
                    double pconst = 0.0, bE = 0.0, bV = 0.0;
                    double f0 = 0.0, g0[2] = {0.0, 0.0}, h0[3] = {0.0, 0.0, 0.0};
                    if (CPA) unit_cp_async_wait<UNIT_RING - 1>();          // this iteration's group has landed
                    if (own) {
                        const PixRec pr = CPA ? ring[(t % UNIT_RING) * 32] : mi.pix[pix];
                        xf = pr.x;
                        skyf = pr.sky;
                        pconst = pr.pixconst;
                        if (mi.bg) {
                            bE = mi.bg[2 * pix];
                            bV = mi.bg[2 * pix + 1];
                        }
                        if (fast) {
                            const double* ccol = cbase + coff;
                            const int n1 = cstr;
                            const double q0 = ccol[0], q1 = ccol[1], q2 = ccol[2], q3 = ccol[3];
                            const double q4 = ccol[n1], q5 = ccol[n1 + 1], q6 = ccol[n1 + 2], q7 = ccol[n1 + 3];
                            const double R2 = si[SI_WX] * q0 + si[SI_WX + 1] * q1 + si[SI_WX + 2] * q2 + si[SI_WX + 3] * q3;
                            const double R3 = si[SI_WX] * q4 + si[SI_WX + 1] * q5 + si[SI_WX + 2] * q6 + si[SI_WX + 3] * q7;
                            const double wy0 = si[SI_WY], wy1 = si[SI_WY + 1], wy2 = si[SI_WY + 2], wy3 = si[SI_WY + 3];
                            const double v = wy0 * R0 + wy1 * R1 + wy2 * R2 + wy3 * R3;
                            double gx = 0.0, gy = 0.0, hxx = 0.0, hxy = 0.0, hyy = 0.0;
                            if (MODE >= 1) {
                                const double D2 = si[SI_DWX] * q0 + si[SI_DWX + 1] * q1 + si[SI_DWX + 2] * q2 + si[SI_DWX + 3] * q3;
                                const double D3 = si[SI_DWX] * q4 + si[SI_DWX + 1] * q5 + si[SI_DWX + 2] * q6 + si[SI_DWX + 3] * q7;
                                gx = wy0 * D0 + wy1 * D1 + wy2 * D2 + wy3 * D3;
                                gy = si[SI_DWY] * R0 + si[SI_DWY + 1] * R1 + si[SI_DWY + 2] * R2 + si[SI_DWY + 3] * R3;
                                if (MODE >= 2) {
                                    const double Q2 = si[SU_DDWX] * q0 + si[SU_DDWX + 1] * q1 + si[SU_DDWX + 2] * q2 + si[SU_DDWX + 3] * q3;
                                    const double Q3 = si[SU_DDWX] * q4 + si[SU_DDWX + 1] * q5 + si[SU_DDWX + 2] * q6 + si[SU_DDWX + 3] * q7;
                                    hxx = wy0 * Q0 + wy1 * Q1 + wy2 * Q2 + wy3 * Q3;
                                    hxy = si[SI_DWY] * D0 + si[SI_DWY + 1] * D1 + si[SI_DWY + 2] * D2 + si[SI_DWY + 3] * D3;
                                    hyy = si[SU_DDWY] * R0 + si[SU_DDWY + 1] * R1 + si[SU_DDWY + 2] * R2 + si[SU_DDWY + 3] * R3;
                                    Q0 = Q2;
                                    Q1 = Q3;
                                }
                                D0 = D2;
                                D1 = D3;
                            }
                            R0 = R2;
                            R1 = R3;
                            if (v < 0) {                                          // softpluslikeinv, fsm_util.jl:222
                                const double e = 1e-3 * exp_nonpos(v);
                                f0 = e;
                                g0[0] = e * gx;
                                g0[1] = e * gy;
                                if (MODE >= 2) {
                                    h0[0] = e * (gx * gx + hxx);
                                    h0[1] = e * (gx * gy + hxy);
                                    h0[2] = e * (gy * gy + hyy);
                                }
                            } else {
                                f0 = 1e-3 * (v + 1.0);
                                g0[0] = 1e-3 * gx;
                                g0[1] = 1e-3 * gy;
                                if (MODE >= 2) {
                                    h0[0] = 1e-3 * hxx;
                                    h0[1] = 1e-3 * hxy;
                                    h0[2] = 1e-3 * hyy;
                                }
                            }
                        }
                    }
                    double l5 = 0.0;
                    if (!isnan(xf)) {                                        // active and not masked (elbo_objective.jl:445, :459)
                        PixelConsts pc;
                        pc.x = (double)xf;
                        pc.iota = iota_h;
                        pc.pixconst = pconst;
                        const bool covered = iown < ncov;      // the last column of the patch is not covered by its own source (:349)
                        const double f1 = theta * T[0] + (1.0 - theta) * T[1];
                        double r[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                        if (MODE >= 1) {
                            r[0] = -T[2];
                            r[1] = -T[3];
                            r[2] = 0.5 * T[4];
                            r[3] = T[5];
                            r[4] = 0.5 * T[6];
                            r[5] = T[0] - T[1];                                   // gal_frac_dev, fsm_util.jl:277-291
                        }
                        if (covered) {
                            if (!fast)
                                star_eval<MODE>(LdGlobal(), mi.coefs, mi.n1, mi.n2, (double)h - si[SI_M] + 26.0,
                                                (double)(w0 + iown) - si[SI_M + 1] + 26.0, f0, g0, h0);
                            cnt_active += 1.0;
                        }
                        const double cb[4] = {si[SI_CB], si[SI_CB + 1], si[SI_CB + 2], si[SI_CB + 3]};
                        l5 = unit_pixel_term<MODE>(acc, pc, (double)skyf + bE, bV, covered, cb, f0, g0, h0, f1, r, val, s_logtab);
                    }
                    if (MODE >= 2 && own) l5plane[pix] = l5;
                    pix += 2;
                    coff += 2 * cstr;
                }
            }
        }

        __syncwarp();

        // ---- fixed-order warp reduction -> the unit's partial vector -------------------------------------------------
        double* out = plan.partials + (size_t)uh.pidx * NACC;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            val += __shfl_xor_sync(0xffffffffu, val, o);
            cnt_active += __shfl_xor_sync(0xffffffffu, cnt_active, o);
        }
        if (lane == 0) {
            out[ACC_VAL] = val;
            out[ACC_CNT_ACTIVE] = cnt_active;
            double ci = 0.0;                                   // neighbour visits of the (sub, image): exact integer sums
            if (uh.hasbg && uh.first)
                for (int q = uh.bgp0; q < uh.bgp1; ++q) ci += plan.bg_cnt[q];
            out[ACC_CNT_INACTIVE] = ci;
        }
        if (MODE == 2 && lane < 3) out[ACC_CC + 7 + lane] = 0.0;
        for (int a = lane; a < NUA; a += 32) {
            const double* row = wbase + (size_t)a * 32;
            double s = 0.0;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) s += row[(lane + i) & 31];
            out[unit_acc_index(a)] = s;
        }
    }
}

// Phase B + fold: L5-weighted moments of every component of the active source, one lane per component, then the
// component's share of the 20 second-order mixture sums (hess_moments.inc), reduced over the warp and added to the
// unit's HH sums in plan.partials (written by unit_walk_kernel<2> earlier on the stream).
__global__ void __launch_bounds__(UNIT_THREADS, CELESTE_UNIT_MOM_MINB)
    unit_moment_kernel(PlanDev plan, const UnitHdr* __restrict__ units, int n_units, int* __restrict__ queue,
                       const double* __restrict__ vp) {
    CEL_DYNAMIC_SMEM(smem);
    __shared__ double s_exptab[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kk = tid & 1;
    (void)kk;
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    double* s_rec = smem + (size_t)warp * (NC2 * MREC);
    __syncthreads();
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const UnitHdr uh = units[u];
        if (plan.task_mask && !plan.task_mask[uh.task]) continue;
        const FieldDev field = plan.fields[uh.field];
        const int n = uh.n, aslot = uh.aslot;
        const PatchDev* prow = field.patches + (size_t)n * field.S_tot;       // patches of image n, by source row
        const PatchDev& pa = prow[plan.src_row[aslot]];
        const int band0 = field.images[n].band - 1;
        const int H2 = pa.H2, W2 = pa.W2, off_h = pa.off_h, off_w = pa.off_w;
        if (H2 <= 0 || W2 <= 1) continue;                    // no covered pixel: nothing to add
        __syncwarp();
        const double* vs = vp + (size_t)NPARAM * aslot;
        if (lane < NC2) {
            double xx[3];
            galaxy_xixi(vs[3], vs[4], vs[5], xx[0], xx[1], xx[2]);
            march_make_record(pa, vs, xx, s_rec, lane);
        }
        __syncwarp();
        const double* l5plane = plan.l5 + plan.l5_ptr[uh.tn];
        double OUT[20];
#pragma unroll
        for (int q = 0; q < 20; ++q) OUT[q] = 0.0;
        const int ncols = W2 - 1;                                // covered columns (:349)
        if (lane < NC2 && ncols > 0) {
            const double* o = s_rec + lane * MREC;
            const double l11 = o[0], l12 = o[1], l22 = o[2], cc = o[3], mu1 = o[4], mu2 = o[5], z = o[6];
            double Dm[15];
#pragma unroll
            for (int q = 0; q < 15; ++q) Dm[q] = 0.0;
            const int nsb = (ncols + MARCH_MAXSEG - 1) / MARCH_MAXSEG;
            const int sw = (ncols + nsb - 1) / nsb;
            // Start values of consecutive rows obey the same recurrence in the ROW direction (q(h + 1) - q(h) is linear
            // in h): f(h + 1, c0) = f(h, c0) gh(h), gh(h + 1) = gh(h) exp(-L11), and the column ratio
            // r(h + 1) = r(h) exp(-L12).  Used when a row is one segment, restarted exactly every MOMENT_ROWBLOCK rows and
            // whenever the component was asleep at the last exact start.
            const double ch = exp_scaled_tab(l11, -1.0, s_exptab), cl = exp_scaled_tab(fmin(fmax(l12, -700.0), 700.0), -1.0, s_exptab);
            double f0 = 0.0, gh = 0.0, r0 = 0.0;
            int since_exact = MOMENT_ROWBLOCK;                       // rows since the last exact start of column 0
            for (int h2 = uh.h2_lo; h2 < uh.h2_hi; ++h2) {
                const double d1 = (double)(off_h + h2 + 1) - mu1;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
                for (int c0 = 0; c0 < ncols; c0 += sw) {
                    const int c1 = min(c0 + sw, ncols);
                    int c = c0;
                    while (c < c1) {
                        double d2 = (double)(off_w + c + 1) - mu2;
                        double f, r;
                        bool sleep = false;
                        if (nsb == 1 && c == 0 && since_exact < MOMENT_ROWBLOCK) {
                            f0 *= gh;
                            gh *= ch;
                            r0 *= cl;
                            f = f0;
                            r = r0;
                            ++since_exact;
                        } else {
                            // exact (re)start of the recurrence at column c (the rule of march_start)
                            const double p1 = l11 * d1 + l12 * d2;
                            const double p2 = l12 * d1 + l22 * d2;
                            const double q = d1 * p1 + d2 * p2;
                            const double ra = -(p2 + 0.5 * l22), rh = -(p1 + 0.5 * l11);
                            sleep = q > MARCH_Q_SLEEP || ra > 700.0;
                            f = sleep ? 0.0 : z * exp_scaled_tab(q, -0.5, s_exptab);
                            r = sleep ? 0.0 : exp_scaled_tab(fmin(ra, 700.0), 1.0, s_exptab);
                            if (nsb == 1 && c == 0) {
                                const bool rowok = !sleep && rh < 700.0 && rh > -700.0;
                                f0 = f;
                                r0 = r;
                                gh = rowok ? exp_scaled_tab(rh, 1.0, s_exptab) : 0.0;
                                since_exact = rowok ? 1 : MOMENT_ROWBLOCK;
                            }
                        }
                        const int cend = sleep ? min(c1, c + MARCH_CAREFUL_COLS) : c1;
                        const double* lp = l5plane + (size_t)h2 * W2 + c;
#define CEL_MOMENT_STEP(L5V)                 \
    {                                        \
        const double v = f * (L5V);          \
        const double v1 = v * d2;            \
        s0 += v;                             \
        const double v2 = v1 * d2;           \
        s1 += v1;                            \
        const double v3 = v2 * d2;           \
        s2 += v2;                            \
        s3 += v3;                            \
        s4 = fma(v3, d2, s4);                \
        f *= r;                              \
        r *= cc;                             \
        d2 += 1.0;                           \
    }
                        for (; c + 4 <= cend; c += 4) {           // four L5 loads in flight
                            const double a0 = lp[0], a1 = lp[1], a2 = lp[2], a3 = lp[3];
                            CEL_MOMENT_STEP(a0)
                            CEL_MOMENT_STEP(a1)
                            CEL_MOMENT_STEP(a2)
                            CEL_MOMENT_STEP(a3)
                            lp += 4;
                        }
                        for (; c < cend; ++c) {
                            CEL_MOMENT_STEP(*lp)
                            lp += 1;
                        }
#undef CEL_MOMENT_STEP
                    }
                }
                // D[a][b] += d1^a s_b; index order 00 01 02 03 04 | 10 11 12 13 | 20 21 22 | 30 31 | 40
                const double e1 = d1, e2 = d1 * d1, e3 = e2 * d1, e4 = e2 * e2;
                Dm[0] += s0;
                Dm[1] += s1;
                Dm[2] += s2;
                Dm[3] += s3;
                Dm[4] += s4;
                Dm[5] = fma(e1, s0, Dm[5]);
                Dm[6] = fma(e1, s1, Dm[6]);
                Dm[7] = fma(e1, s2, Dm[7]);
                Dm[8] = fma(e1, s3, Dm[8]);
                Dm[9] = fma(e2, s0, Dm[9]);
                Dm[10] = fma(e2, s1, Dm[10]);
                Dm[11] = fma(e2, s2, Dm[11]);
                Dm[12] = fma(e3, s0, Dm[12]);
                Dm[13] = fma(e3, s1, Dm[13]);
                Dm[14] = fma(e4, s0, Dm[14]);
            }
            unit_moments_to_sums(l11, l12, l22, Dm, OUT);
            // weights of the component: thc (u), thc nu (xs), thc nu^2 (ss), sg (tx), sg nu (ts)
            const int j = lane >> 1;
            const double theta = vs[2];
            const double thc = j < NPROTO_DEV ? theta : 1.0 - theta;
            const double sg = j < NPROTO_DEV ? 1.0 : -1.0;
            const double nu = c_proto_nu[j];
#pragma unroll
            for (int q = 0; q < 3; ++q) OUT[q] *= thc;
#pragma unroll
            for (int q = 3; q < 9; ++q) OUT[q] *= thc * nu;
#pragma unroll
            for (int q = 9; q < 15; ++q) OUT[q] *= thc * nu * nu;
#pragma unroll
            for (int q = 15; q < 17; ++q) OUT[q] *= sg;
#pragma unroll
            for (int q = 17; q < 20; ++q) OUT[q] *= sg * nu;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 20; ++q) OUT[q] += __shfl_xor_sync(0xffffffffu, OUT[q], o);
        }
        if (lane < 20) {
            // R of GalRaw (elbo_math.cuh gal_eval): xx = (2 u1, u2, 2 u3); x-Sigma = xs; x-theta = -tx; Sigma-Sigma = ss;
            // Sigma-theta = ts.  Lane q owns sum q (every lane holds all 20 totals after the butterfly).
            double mine = 0.0;
#pragma unroll
            for (int q = 0; q < 20; ++q)
                if (lane == q) mine = OUT[q];
            const int kq[20] = {0, 0, 1, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 0, 1, 2, 3, 4};
            const int lq[20] = {0, 1, 1, 2, 3, 4, 2, 3, 4, 2, 3, 4, 3, 4, 4, 5, 5, 5, 5, 5};
            const double wq = (lane == 0 || lane == 2) ? 2.0 : ((lane == 15 || lane == 16) ? -1.0 : 1.0);
            double* out = plan.partials + (size_t)uh.pidx * NACC_MODE2 + ACC_HH + tri6(kq[lane], lq[lane]);
            *out += wq * mine;
        }
    }
}

// Host side: the unit list of a plan (heaviest first), each unit's column segmentation, and the partial vectors:
// chunk_ptr[sub * N + n] .. chunk_ptr[sub * N + n + 1] are the partials of (sub, image) for epilogue_kernel.
// A (sub, image) is cut into units of at most `unit_rows` rows (equal pieces).  The cut depends on the patch ONLY, never
// on what else is in the plan: a task's result is bit-for-bit the same alone, in a batch, or in another rank's shard.
// Small pieces also let a single source spread over many SMs (the latency of a celeste_elbo_single call) and keep the
// launch tail of a small plan (a rank of an 8-GPU run) short.
// geo(slot, n, off_h, off_w, H2, W2) -> the patch box of a slot in image n.
template <typename Geo>
inline void build_unit_list(int n_subs, int N, const int* sub_task, const int* sub_slot, const int* task_ptr,
                            const int* task_field, Geo geo, long unit_rows, long unit_pixels, int bg_pixels,
                            std::vector<UnitHdr>& units,
                            std::vector<UnitHdr>& bg_units, std::vector<int>& chunk_ptr, long long& maxpix) {
    units.clear();
    bg_units.clear();
    chunk_ptr.assign((size_t)n_subs * N + 1, 0);
    maxpix = 1;
    if (unit_rows <= 0) unit_rows = 1L << 30;                                  // 0: never cut
    std::vector<long> cost;
    for (int u = 0; u < n_subs; ++u) {
        const int t = sub_task[u];
        const int aslot = sub_slot[u], slot0 = task_ptr[t], slot1 = task_ptr[t + 1];
        for (int n = 0; n < N; ++n) {
            int oh, ow, H2, W2;
            geo(aslot, n, oh, ow, H2, W2);
            UnitHdr uh{};
            uh.aslot = aslot;
            uh.slot0 = slot0;
            uh.slot1 = slot1;
            uh.field = task_field ? task_field[t] : 0;
            uh.sub = u;
            uh.task = t;
            uh.n = n;
            uh.tn = u * N + n;
            uh.hasbg = 0;
            uh.nseg = 1;
            uh.first = 1;
            int pieces = 1;
            if (H2 > 0 && W2 > 0) {
                maxpix = std::max(maxpix, (long long)H2 * W2);
                for (int s = slot0; s < slot1; ++s) {
                    if (s == aslot) continue;
                    int ph, pw, pH2, pW2;
                    geo(s, n, ph, pw, pH2, pW2);
                    if (pH2 > 0 && pW2 > 1) {
                        const int h_lo = std::max(oh, ph) + 1, h_hi = std::min(oh + H2, ph + pH2);
                        const int w_lo = std::max(ow, pw) + 1, w_hi = std::min(ow + W2, pw + pW2 - 1);
                        if (h_hi >= h_lo && w_hi >= w_lo) {
                            uh.hasbg = 1;
                            uh.nbpix += (h_hi - h_lo + 1) * (w_hi - w_lo + 1);
                        }
                    }
                }
                pieces = (int)std::max(1L, (H2 + unit_rows - 1) / unit_rows);
                if (unit_pixels > 0)      // ... or by pixels: pieces of about unit_pixels pixels, at least 4 rows each
                    pieces = (int)std::max(1L, std::min((long)std::max(1, H2 / 4), ((long)H2 * W2 + unit_pixels / 2) / unit_pixels));
            }
            const int tn = u * N + n;
            chunk_ptr[tn + 1] = chunk_ptr[tn] + pieces;
            for (int pc = 0; pc < pieces; ++pc) {
                UnitHdr x = uh;
                if (unit_pixels > 0 || pieces == 1) {
                    x.h2_lo = (int)((long)std::max(H2, 0) * pc / pieces);
                    x.h2_hi = (int)((long)std::max(H2, 0) * (pc + 1) / pieces);
                } else {
                    // pieces of exactly unit_rows rows and a remainder: with unit_rows = the 16 walk slots of a warp a
                    // full piece is ONE round of whole-row walks (one exact start per row); equal halves of a 23-row
                    // patch would need 3 rounds of 6-pixel segments each (measured: tuning_r02.md)
                    x.h2_lo = (int)std::min((long)H2, pc * unit_rows);
                    x.h2_hi = (int)std::min((long)H2, (pc + 1) * unit_rows);
                }
                x.first = pc == 0;
                x.pidx = chunk_ptr[tn] + pc;
                const int rows = x.h2_hi - x.h2_lo;
                if (rows > 0 && W2 > 0) {
                    // column segments per row: minimise rounds of 16 walks x (iterations of two columns + an exact start)
                    const int nmin = std::max(1, (W2 + MARCH_MAXSEG - 1) / MARCH_MAXSEG);
                    long best = -1;
                    for (int cand = nmin; cand < nmin + 8; ++cand) {
                        const int L = (W2 + cand - 1) / cand;
                        const long cst = (long)((rows * cand + NPW - 1) / NPW) * (10 * ((L + 1) / 2) + 5);
                        if (best < 0 || cst < best) {
                            best = cst;
                            x.nseg = cand;
                        }
                    }
                }
                if (pc == 0 && x.hasbg) {
                    // neighbour work of the (sub, image): row pieces of at most ~UNIT_BG_PIXELS shared pixels, so that one
                    // crowded source is not the tail of unit_bg_kernel
                    const int bp = std::max(1, std::min(std::max(1, H2 / 2), (x.nbpix + bg_pixels - 1) / bg_pixels));
                    x.bgp0 = (int)bg_units.size();
                    x.bgp1 = x.bgp0 + bp;
                    for (int b = 0; b < bp; ++b) {
                        UnitHdr y = uh;
                        y.h2_lo = (int)((long)H2 * b / bp);
                        y.h2_hi = (int)((long)H2 * (b + 1) / bp);
                        y.bgp0 = x.bgp0 + b;
                        y.bgp1 = y.bgp0 + 1;
                        y.nbpix = uh.nbpix / bp;
                        bg_units.push_back(y);
                    }
                }
                units.push_back(x);
                cost.push_back((long)std::max(rows, 0) * std::max(W2, 0));
            }
        }
    }
    std::vector<int> order(units.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<UnitHdr> sorted(units.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = units[order[i]];
    units.swap(sorted);
    std::stable_sort(bg_units.begin(), bg_units.end(), [](const UnitHdr& a, const UnitHdr& b) { return a.nbpix > b.nbpix; });
}

}  // namespace celeste
#endif
