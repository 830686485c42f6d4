// march_kernels.cuh -- the value / gradient hot loop of the ELBO (add_pixel_term!, elbo_objective.jl:330-392)
// re-organised around the pixel GRID instead of the pixel: a walk moves along one row of the active source's patch
// (consecutive columns w, w+1, ...) and carries every Gaussian component with it.
//
// Why: on a regular grid the quadratic form of a bivariate normal (eval_bvn_pdf!, BivariateNormals.jl:208-222)
// changes by a LINEAR amount from one column to the next,
//       q(w + 1) - q(w) = 2 p2(w) + L22,      p2(w + 1) - p2(w) = L22,       p = Lambda (x - mu),
// so   f(w + 1) = f(w) r(w),   r(w + 1) = r(w) c,   c = exp(-L22),   r(w0) = exp(-(p2(w0) + L22 / 2)):
// after an exact start (two exp per component) every further pixel costs two multiplications per component
// instead of a quadratic form and an exp (20 of the 33 FP64 instructions a component costs in task_kernel's
// gradient mode).  The cubic B-spline of the star (star_light_density!, fsm_util.jl:225-248) has the same
// structure: along a row the fractional offsets -- hence all eight weights -- are constant and the 4 x 4 tap window
// slides, so new columns of 4 taps replace 16 taps per pixel.
// A walk is restarted exactly after at most MARCH_MAXSEG pixels (51 = the catalog patch cap,
// imaged_sources.jl:173-176), which bounds the accumulated rounding error: relative ~ n^2/2 ulp after n steps,
// 1.4e-13 at n = 51 -- five orders below the 1e-8 parity tolerance (measured against task_kernel: <= 5e-13).
//
// A walk is carried by a PAIR of adjacent lanes, one PSF component (K = 2) each: 28 doubles of state per lane.
// Every iteration advances two columns; the lanes exchange half of their mixture sums by one shuffle per sum and
// each finishes one of the two pixels (star, pixel term), so nothing is computed twice.
//
// Neighbouring sources (value only, elbo_objective.jl:38-40,69) are walked the same way over the intersection
// of their patch with the active patch, one neighbour at a time, into a per-task background buffer
// (E_bg, V_bg per active pixel) that the walk of the active source then reads.
//
// Gradients leave the block in TASK space -- world position (the -J' of each image already applied), raw
// covariance, raw gal_frac_dev, and the four c-scalars of each BAND -- so that all images of a task reduce
// into one 29-vector and march_epilogue_kernel applies the remaining chain rule once per source.
//
// Handles Sa = 1 and K = 2 (production: ParallelRun.jl:253,489, elbo_args.jl:197); every other plan keeps
// task_kernel.  Same arithmetic as there up to reassociation and the recurrence's rounding.
#ifndef CELESTE_MARCH_KERNELS_CUH
#define CELESTE_MARCH_KERNELS_CUH

#include "celeste_kernels.cuh"

#ifdef CELESTE_HOST_EMULATION
#define CEL_PREFETCH_L1(p) ((void)(p))
#else
#define CEL_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#endif

namespace celeste {

#ifndef CELESTE_MARCH_MAXSEG
#define CELESTE_MARCH_MAXSEG 51
#endif
#ifndef CELESTE_MARCH_MINB
#define CELESTE_MARCH_MINB 3
#endif
#ifndef CELESTE_MARCH_THREADS
#define CELESTE_MARCH_THREADS 128
#endif
constexpr int MARCH_THREADS = CELESTE_MARCH_THREADS;   // multiple of 32
constexpr int MARCH_NIMG = 5;
constexpr int MARCH_MAXSEG = CELESTE_MARCH_MAXSEG;
constexpr int NC2 = NPROTO * 2;       // components of a K = 2 source
constexpr int MREC = 8;               // staged component record: L11 L12 L22 c | mu1 mu2 z -
// per (source, image) constants staged in shared memory
constexpr int SI_CB = 0;              // A1 A2 B1 B2
constexpr int SI_M = 4;               // m_pos
constexpr int SI_WX = 6, SI_DWX = 10, SI_WY = 14, SI_DWY = 18;   // spline weights at the patch's fractional offsets
constexpr int SI_J = 22;              // wcs_jacobian
constexpr int SI_THETA = 26;
constexpr int SI_STRIDE = 28;
// task-space accumulators
constexpr int TA_VAL = 0, TA_CNT_ACTIVE = 1, TA_CNT_INACTIVE = 2, TA_POS = 3, TA_SIG = 5, TA_THETA = 8, TA_BAND = 9;
constexpr int NT_ACC = TA_BAND + 4 * 5;   // 29

struct MarchBox {     // intersection of a neighbour's patch with the active patch, 1-based image coordinates
    int h0, w0, nh, nw;
};

// m_pos of a source in an image: linear_world_to_pix, wcs_utils.jl:14-18 (the same expression as setup_kernel)
__device__ __forceinline__ void march_m_pos(const PatchDev& p, const double* vs, double& m1, double& m2) {
    const double d0 = vs[0] - p.wc[0], d1 = vs[1] - p.wc[1];
    m1 = (p.J[0] * d0 + p.J[2] * d1) + p.pc[0];
    m2 = (p.J[1] * d0 + p.J[3] * d1) + p.pc[1];
}

// Component record c = 2 j + k of (source, image) for marching, computed in place from the source's parameters --
// load_bvn_mixtures! (fsm_util.jl:111-169) with exactly setup_kernel's arithmetic -- plus the column ratio
// exp(-L22).  xixi: the source's XiXi (galaxy_xixi: one sin / cos per source and block, by one thread).  The march path needs no set-up launch and no 2 KB-per-(source, image) scratch round trip through HBM.
// One thread per component.
__device__ inline void march_make_record(const PatchDev& p, const double* vs, const double* xixi, double* dst, int c) {
    const int j = c >> 1, k = c & 1;
    double m1, m2, t[COMP_STRIDE];
    march_m_pos(p, vs, m1, m2);
    make_component_xi(p.psf + 7 * k, c_proto_eta[j], c_proto_nu[j], m1, m2, xixi[0], xixi[1], xixi[2], t);
    double* o = dst + c * MREC;
    o[0] = t[2];
    o[1] = t[3];
    o[2] = t[4];
    o[3] = exp(-t[4]);
    o[4] = t[0];
    o[5] = t[1];
    o[6] = t[5];
    o[7] = 0.0;
}

// per-(source, image) constants: brightness scalars of the image's band (source_brightness.jl:27-202, the sums in
// brightness_values' order), m_pos, spline weights, Jacobian
template <int MODE>
__device__ inline void march_stage_srcimg(const PatchDev& p, const double* vs, int band0, double* si) {
    double ka[10], la[10];
    band_coefs(band0, ka, la);
    for (int i = 0; i < 2; ++i) {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 10; ++k) {
            const double beta = vs[bright_id(i, k)];
            s1 += ka[k] * beta;
            s2 += la[k] * beta;
        }
        si[SI_CB + i] = vs[26 + i] * exp(s1);
        si[SI_CB + 2 + i] = vs[26 + i] * exp(s2);
    }
    double m1, m2;
    march_m_pos(p, vs, m1, m2);
    si[SI_M] = m1;
    si[SI_M + 1] = m2;
    const double ax = (double)(p.off_h + 1) - m1 + 26.0, ay = (double)(p.off_w + 1) - m2 + 26.0;
    double dd[4];
    cubic_weights<(MODE >= 1 ? 1 : 0)>(ax - floor(ax), si + SI_WX, si + SI_DWX, dd);
    cubic_weights<(MODE >= 1 ? 1 : 0)>(ay - floor(ay), si + SI_WY, si + SI_DWY, dd);
    for (int i = 0; i < 4; ++i) si[SI_J + i] = p.J[i];
    si[SI_THETA] = vs[2];
    si[SI_THETA + 1] = 0.0;
}

// A walk is carried by a PAIR of adjacent lanes: lane parity kk = PSF component (K = 2), so each lane keeps the 14
// prototype components of one PSF component in registers (28 doubles of state instead of 56: twice the warps per
// SM).  Every iteration advances two columns: both lanes accumulate their half of the mixture sums of both
// pixels, exchange one pixel's sums by a shuffle, and each lane finishes ITS pixel (star, pixel term) -- no work is
// duplicated.  All loops are warp-uniform (iteration counts are maxima over the warp), so shuffles use full masks.
constexpr int NPW = 16;                         // pairs (walks) per warp
constexpr int NPAIR = MARCH_THREADS / 2;        // walks in flight per block

// exact start at image pixel (hh, ww): f = z exp(-q/2) and the column ratio r for the 14 components of PSF
// component kk (records c = 2 j + kk)
// Returns true if some component was put to sleep: a Gaussian that is below ~1e-282 of its peak at the start
// (q / 2 > 650) cannot be carried by the recurrence -- its value is not representable, and the product of its column
// ratios towards the centre would overflow -- so it gets f = r = 0 (it contributes exactly nothing) and the caller
// restarts the walk after MARCH_CAREFUL_COLS columns, until no component sleeps.  Within that many columns a
// sleeping component stays below exp(-80) of its peak for every precision L22 <= 34 (Gaussians down to 0.17 px
// wide); sharper ones do not occur behind a sampled PSF.  Catalog-sized patches of SDSS-like PSFs never get here
// (q / 2 <= 450 at the corner of a 51 x 51 patch for a 1.2 px PSF core).
constexpr int MARCH_CAREFUL_COLS = 4;
constexpr double MARCH_Q_SLEEP = 1300.0;
__device__ __forceinline__ bool march_start(const double* recs_k, const double* etab, double hh, double ww, double* fp,
                                            double* rr) {
    bool asleep = false;
#pragma unroll
    for (int j = 0; j < NPROTO; ++j) {
        const double* o = recs_k + j * 2 * MREC;
        const double l11 = o[0], l12 = o[1], l22 = o[2], mu1 = o[4], mu2 = o[5], z = o[6];
        const double d1 = hh - mu1, d2 = ww - mu2;
        const double p1 = l11 * d1 + l12 * d2;
        const double p2 = l12 * d1 + l22 * d2;
        const double q = d1 * p1 + d2 * p2;
        // exp(-(q(w+1) - q(w)) / 2); the argument may be positive (walking towards the centre)
        const double ra = -(p2 + 0.5 * l22);
        const bool sleep = q > MARCH_Q_SLEEP || ra > 700.0;      // (false for NaN: non-finite parameters keep propagating)
        asleep |= sleep;
        const double f = z * exp_scaled_tab(q, -0.5, etab);
        const double r = exp_scaled_tab(fmin(ra, 700.0), 1.0, etab);
        fp[j] = sleep ? 0.0 : f;
        rr[j] = sleep ? 0.0 : r;
    }
    return asleep;
}

__device__ __forceinline__ int warp_max_int(int v) {
    double d = (double)v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    return (int)d;
}

// pointers and geometry of one image and of the active source's patch in it (shared memory)
struct MarchImg {
    const float* pixels;
    const float* sky;
    const double* pixconst;
    const float* iota;
    const uint8_t* bitmap;
    const double* coefs;
    double* bg;                 // (E_bg, V_bg) planes of this (sub, image), or null
    int H2, W2, off_h, off_w, imgH, n1, n2, band0;
};

// unit tables of a block (shared memory): walks are numbered image by image, rows fastest
struct MarchUnits {
    int ubeg[MARCH_NIMG + 1];
    int nseg[MARCH_NIMG];
    MarchBox box[MARCH_NIMG];
};

// Everything a block needs to find its work; built on the host from the patch geometry (build_march_blocks).
struct MarchHdr {
    int aslot, slot0, slot1;     // slot of the active source; slot range of its task
    int field, sub, task;
    int n0, n1;                  // image range of this block
    int pidx;                    // this block's NT_ACC-vector in plan.partials
    int nseg;                    // column segments per row
    unsigned hasbg;              // bit k: some other source of the task reaches image n0 + k (background buffer in use)
    int ubeg[MARCH_NIMG + 1];    // first walk of each image (walks are numbered image by image, rows fastest)
};

// One block per (active source of a task, group of <= MARCH_NIMG images).
template <int MODE>
__global__ void __launch_bounds__(MARCH_THREADS, CELESTE_MARCH_MINB)
    march_kernel(PlanDev plan, const MarchHdr* __restrict__ blocks, const double* __restrict__ vp) {
    static_assert(MODE <= 1, "the Hessian mode uses pixel_kernel");
    constexpr int NUA = MODE == 0 ? 1 : NACC_MODE1;   // (c, y)-space accumulators of the current walk
    constexpr int NS = MODE == 0 ? 2 : 7;             // mixture sums per pixel: F_dev F_exp | AX1 AX2 AS1 AS2 AS3
    CEL_DYNAMIC_SMEM(smem);
    double* tacc = smem;                                   // NT_ACC x MARCH_THREADS: task-space sums of this thread
    double* s_rec = tacc + NT_ACC * MARCH_THREADS;         // MARCH_NIMG x NC2 x MREC: the source being walked
    double* s_si = s_rec + MARCH_NIMG * NC2 * MREC;        // MARCH_NIMG x SI_STRIDE
    __shared__ double s_exptab[8];
    __shared__ double s_logtab[256];
    __shared__ MarchUnits s_nu;                            // walks of the current neighbour
    __shared__ MarchImg s_img[MARCH_NIMG];
    __shared__ MarchHdr s_hdr;
    __shared__ double s_xixi[2][4];                        // XiXi of the active source / of the current neighbour

    const int tid = threadIdx.x, lane = tid & 31, kk = tid & 1;
    const int pair0 = (tid >> 5) * NPW;                    // first walk slot of this warp
    if (tid < (int)(sizeof(MarchHdr) / sizeof(int))) reinterpret_cast<int*>(&s_hdr)[tid] = reinterpret_cast<const int*>(blocks + blockIdx.x)[tid];
    if (tid == MARCH_THREADS - 1) {
        const double* vs = vp + (size_t)NPARAM * blocks[blockIdx.x].aslot;
        galaxy_xixi(vs[3], vs[4], vs[5], s_xixi[0][0], s_xixi[0][1], s_xixi[0][2]);
    }
    __syncthreads();
    const MarchHdr& th = s_hdr;
    if (plan.task_mask && !plan.task_mask[th.task]) return;
    const int slot0 = th.slot0, slot1 = th.slot1, aslot = th.aslot;
    const FieldDev field = plan.fields[th.field];
    const int nimg = th.n1 - th.n0;
    const PatchDev* apatch = field.patches + plan.src_row[aslot];   // + n * S_tot
    const long long* bgp = plan.bg_ptr + (size_t)th.sub * plan.N;

#pragma unroll
    for (int a = 0; a < NT_ACC; ++a) tacc[a * MARCH_THREADS + tid] = 0.0;
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
    for (int i = tid; i < 256; i += MARCH_THREADS) s_logtab[i] = h_logtab[i];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
    for (int i = tid; i < 256; i += MARCH_THREADS) s_logtab[i] = g_logtab[i];
#endif
    const unsigned hasbg = th.hasbg;

    // ---- neighbours, one at a time in slot order: E_bg += E_s, V_bg += E2_s - E_s^2 over the shared pixels -------
    if (hasbg) {
        for (int k = 0; k < nimg; ++k)
            if ((hasbg >> k) & 1u) {
                const int n = th.n0 + k;
                const PatchDev& pa = apatch[(size_t)n * field.S_tot];
                double* bg = plan.bg + bgp[n];
                const int tot = 2 * pa.H2 * pa.W2;
                for (int i = tid; i < tot; i += MARCH_THREADS) bg[i] = 0.0;
            }
        for (int s = slot0; s < slot1; ++s) {
            if (s == aslot) continue;
            __syncthreads();      // the previous neighbour's sums (or the zeros) are in place; s_nu / s_nrec are free
            if (tid == (MARCH_THREADS > 32 ? 32 : 1)) {
                const double* vs = vp + (size_t)NPARAM * s;
                galaxy_xixi(vs[3], vs[4], vs[5], s_xixi[1][0], s_xixi[1][1], s_xixi[1][2]);
            }
            if (tid == 0) {
                s_nu.ubeg[0] = 0;
                for (int k = 0; k < MARCH_NIMG; ++k) {
                    int units = 0;
                    MarchBox bx{0, 0, 0, 0};
                    int nsg = 1;
                    if (k < nimg && ((hasbg >> k) & 1u)) {
                        const int n = th.n0 + k;
                        const PatchDev& pa = apatch[(size_t)n * field.S_tot];
                        const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
                        // active pixels: rows off+1..off+H2, columns off+1..off+W2; the neighbour covers columns
                        // off+1..off+W2-1 only (strict `w2 < W2`, elbo_objective.jl:349)
                        const int h_lo = max(pa.off_h, p.off_h) + 1, h_hi = min(pa.off_h + pa.H2, p.off_h + p.H2);
                        const int w_lo = max(pa.off_w, p.off_w) + 1, w_hi = min(pa.off_w + pa.W2, p.off_w + p.W2 - 1);
                        if (h_hi >= h_lo && w_hi >= w_lo) {
                            bx = MarchBox{h_lo, w_lo, h_hi - h_lo + 1, w_hi - w_lo + 1};
                            nsg = (bx.nw + MARCH_MAXSEG - 1) / MARCH_MAXSEG;
                            units = bx.nh * nsg;
                        }
                    }
                    s_nu.box[k] = bx;
                    s_nu.nseg[k] = nsg;
                    s_nu.ubeg[k + 1] = s_nu.ubeg[k] + units;
                }
            }
            __syncthreads();
            const int total = s_nu.ubeg[MARCH_NIMG];
            if (total == 0) continue;                      // uniform over the block
            for (int i = tid; i < nimg * NC2; i += MARCH_THREADS) {
                const int k = i / NC2, c = i - k * NC2;
                if (s_nu.ubeg[k + 1] > s_nu.ubeg[k])
                    march_make_record(field.patches[plan.src_row[s] + (size_t)(th.n0 + k) * field.S_tot], vp + (size_t)NPARAM * s,
                                      s_xixi[1], s_rec + k * NC2 * MREC, c);
            }
            if (tid < nimg && s_nu.ubeg[tid + 1] > s_nu.ubeg[tid]) {
                const int k = tid, n = th.n0 + k;
                march_stage_srcimg<0>(field.patches[plan.src_row[s] + (size_t)n * field.S_tot], vp + (size_t)NPARAM * s,
                                      field.images[n].band - 1, s_si + k * SI_STRIDE);
            }
            __syncthreads();
            double cnt_inactive = 0.0;
            for (int ub = pair0; ub < total; ub += NPAIR) {        // warp-uniform
                const int u = ub + (lane >> 1);
                const bool has = u < total;
                int k = 0;
#pragma unroll
                for (int q = 1; q < MARCH_NIMG; ++q) k += (has && u >= s_nu.ubeg[q]) ? 1 : 0;
                const int n = th.n0 + k;
                const MarchBox bx = s_nu.box[k];
                const int nsg = s_nu.nseg[k];
                const int ul = has ? u - s_nu.ubeg[k] : 0;
                const int nh = max(bx.nh, 1);
                const int seg = ul / nh, row = ul - seg * nh;
                const int segw = (bx.nw + nsg - 1) / nsg;
                const int c0 = seg * segw;
                const int len = has ? max(min(segw, bx.nw - c0), 0) : 0;
                const int nit = warp_max_int((len + 1) >> 1);
                if (nit == 0) continue;
                const int h = bx.h0 + row, w0 = bx.w0 + c0;     // 1-based image coordinates
                const PatchDev& pa = apatch[(size_t)n * field.S_tot];
                const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
                const int aH2 = pa.H2, aW2 = pa.W2, nH2 = p.H2, imgH = field.images[n].H, n1 = p.n1, n2 = p.n2;
                const double* coefs = p.coefs;
                const double* si = s_si + k * SI_STRIDE;
                const double* recs = s_rec + (k * NC2 + kk) * MREC;      // this lane's PSF component
                double fp[NPROTO], rr[NPROTO];
                // star: sliding window of row-interpolated columns, own pixels are columns kk, kk + 2, ...
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
                const uint8_t* abit = pa.bitmap + ah2 + (size_t)acol * aH2;
                const uint8_t* nbit = p.bitmap + nh2 + (size_t)ncol * nH2;
                const float* px = field.images[n].pixels + (size_t)(h - 1) + (size_t)(w0 - 1 + kk) * imgH;
                double* bgE = plan.bg + (has ? bgp[n] : 0) + ah2 + (size_t)acol * aH2;
                const size_t bgplane = (size_t)aH2 * aW2;
                const double theta = si[SI_THETA];
                int t = 0;
                while (t < nit) {
                // (re)start the recurrence exactly at column 2t; a walk with sleeping components restarts every few columns
                // (the decision is taken for the whole warp, so that every loop around the shuffles stays warp-uniform)
                const bool asleep = march_start(recs, s_exptab, (double)h, (double)(w0 + 2 * t), fp, rr);
                const bool careful = __ballot_sync(0xffffffffu, asleep && len > 0) != 0u;
                const int tend = careful ? min(nit, t + MARCH_CAREFUL_COLS / 2) : nit;
                for (; t < tend; ++t) {
                    const bool own = 2 * t + kk < len;
                    unsigned char ab = 0, nb = 0;
                    float xv = 0.f;
                    if (own) {
                        ab = *abit;
                        nb = *nbit;
                        xv = *px;
                    }
                    double R2 = 0.0, R3 = 0.0;
                    if (fast && own) {
                        const double wx0 = si[SI_WX], wx1 = si[SI_WX + 1], wx2 = si[SI_WX + 2], wx3 = si[SI_WX + 3];
                        R2 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                        ccol += n1;
                        R3 = wx0 * __ldg(ccol) + wx1 * __ldg(ccol + 1) + wx2 * __ldg(ccol + 2) + wx3 * __ldg(ccol + 3);
                        ccol += n1;
                    }
                    // this lane's half of the mixture value of both pixels
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
                    // the partner's half of MY pixel
                    double Fd = kk == 0 ? S[0][0] : S[1][0], Fe = kk == 0 ? S[0][1] : S[1][1];
                    Fd += __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][0] : S[0][0], 1);
                    Fe += __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][1] : S[0][1], 1);
                    if (ab && nb && !isnan(xv)) {
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
                        bgE[0] += Es;
                        bgE[bgplane] += E2s - Es * Es;
                        cnt_inactive += 1.0;                                          // elbo_objective.jl:353-357
                    }
                    abit += 2 * aH2;
                    nbit += 2 * nH2;
                    px += 2 * imgH;
                    bgE += 2 * aH2;
                    R0 = R2;
                    R1 = R3;
                }
                }
            }
            tacc[TA_CNT_INACTIVE * MARCH_THREADS + tid] += cnt_inactive;
        }
        __syncthreads();          // every neighbour's sums are visible to the walk of the active source; s_rec is free
    }
    for (int i = tid; i < nimg * NC2; i += MARCH_THREADS) {
        const int k = i / NC2, c = i - k * NC2;
        march_make_record(apatch[(size_t)(th.n0 + k) * field.S_tot], vp + (size_t)NPARAM * aslot, s_xixi[0], s_rec + k * NC2 * MREC, c);
    }
    if (tid < nimg) {
        const int k = tid, n = th.n0 + k;
        const PatchDev& pa = apatch[(size_t)n * field.S_tot];
        const ImageDev& img = field.images[n];
        march_stage_srcimg<MODE>(pa, vp + (size_t)NPARAM * aslot, img.band - 1, s_si + k * SI_STRIDE);
        MarchImg mi;
        mi.pixels = img.pixels;
        mi.sky = img.sky;
        mi.pixconst = img.pixconst;
        mi.iota = img.iota;
        mi.bitmap = pa.bitmap;
        mi.coefs = pa.coefs;
        mi.bg = ((hasbg >> k) & 1u) ? plan.bg + bgp[n] : nullptr;
        mi.H2 = pa.H2;
        mi.W2 = pa.W2;
        mi.off_h = pa.off_h;
        mi.off_w = pa.off_w;
        mi.imgH = img.H;
        mi.n1 = pa.n1;
        mi.n2 = pa.n2;
        mi.band0 = img.band - 1;
        s_img[k] = mi;
    }
    __syncthreads();

    // ---- the active source -----------------------------------------------------------------------------------
    // (pointers and geometry of the images live in shared memory and are re-read where used: the walk keeps its
    //  registers for the 28 doubles of component state and the 14 mixture sums)
    const int total = th.ubeg[MARCH_NIMG];
    for (int ub = pair0; ub < total; ub += NPAIR) {                // warp-uniform
        const int u = ub + (lane >> 1);
        const bool has = u < total;
        int k = 0;
#pragma unroll
        for (int q = 1; q < MARCH_NIMG; ++q) k += (has && u >= th.ubeg[q]) ? 1 : 0;
        const MarchImg& mi = s_img[k];
        const int H2 = max(mi.H2, 1), W2 = mi.W2;
        const int nseg = th.nseg;
        const int ul = has ? u - th.ubeg[k] : 0;
        const int seg = ul / H2, h2 = ul - seg * H2;
        const int segw = (W2 + nseg - 1) / nseg;
        const int c0 = seg * segw;
        const int len = has ? max(min(segw, W2 - c0), 0) : 0;
        const int nit = warp_max_int((len + 1) >> 1);
        if (nit == 0) continue;
        const int ncov = min(len, W2 - 1 - c0);                   // pixels before the (uncovered) last column, :349
        const int h = mi.off_h + h2 + 1, w0 = mi.off_w + c0 + 1;     // 1-based image coordinates
        const double* si = s_si + k * SI_STRIDE;
        const double* recs = s_rec + (k * NC2 + kk) * MREC;          // this lane's PSF component
        double fp[NPROTO], rr[NPROTO];
        const double d1 = (double)h - recs[4];                        // x1 - mu1 of this PSF component
        double d2 = (double)w0 - recs[5];
        // star: own pixels are columns kk, kk + 2, ... of the segment
        bool fast;
        int coff;                                                     // offset of the next spline column in coefs
        {
            const int ixf = (int)floor((double)h - si[SI_M] + 26.0), iy0 = (int)floor((double)w0 - si[SI_M + 1] + 26.0);
            fast = len > 0 && ixf >= 1 && ixf <= mi.n1 - 3 && iy0 >= 1 && iy0 + len - 1 <= mi.n2 - 3;
            coff = fast ? (iy0 - 1 + kk) * mi.n1 + (ixf - 1) : 0;
        }
        double R0 = 0.0, R1 = 0.0, D0 = 0.0, D1 = 0.0;
        if (fast && kk < len) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double* ccol = mi.coefs + coff;
                const double q0 = __ldg(ccol), q1 = __ldg(ccol + 1), q2 = __ldg(ccol + 2), q3 = __ldg(ccol + 3);
                const double r = si[SI_WX] * q0 + si[SI_WX + 1] * q1 + si[SI_WX + 2] * q2 + si[SI_WX + 3] * q3;
                double d = 0.0;
                if (MODE >= 1) d = si[SI_DWX] * q0 + si[SI_DWX + 1] * q1 + si[SI_DWX + 2] * q2 + si[SI_DWX + 3] * q3;
                coff += mi.n1;
                if (b == 0) {
                    R0 = r;
                    D0 = d;
                } else {
                    R1 = r;
                    D1 = d;
                }
            }
        }
        int pix = h2 + (c0 + kk) * H2;                                // own pixel inside the patch
        int ipix = (h - 1) + (w0 - 1 + kk) * mi.imgH;                 // ... and inside the image
        double* ta = tacc + tid;
        double cnt_active = 0.0, val = 0.0;

        int t = 0;
        while (t < nit) {
        // (re)start the recurrence exactly at column 2t; a walk with sleeping components restarts every few columns
        // (the decision is taken for the whole warp, so that every loop around the shuffles stays warp-uniform)
        const bool asleep = march_start(recs, s_exptab, (double)h, (double)(w0 + 2 * t), fp, rr);
        const bool careful = __ballot_sync(0xffffffffu, asleep && len > 0) != 0u;
        const int tend = careful ? min(nit, t + MARCH_CAREFUL_COLS / 2) : nit;
        for (; t < tend; ++t) {
            const int iown = 2 * t + kk;
            const bool own = iown < len;
            // this lane's pixel (column 2t + kk): its inputs are requested now and read after the mixture sums, whose
            // ~1000 instructions hide the latency without holding a register
            if (own) {
                CEL_PREFETCH_L1(mi.pixels + ipix);
                CEL_PREFETCH_L1(mi.sky + ipix);
                CEL_PREFETCH_L1(mi.pixconst + ipix);
                CEL_PREFETCH_L1(mi.bitmap + pix);
                if (mi.bg) {
                    CEL_PREFETCH_L1(mi.bg + pix);
                    CEL_PREFETCH_L1(mi.bg + pix + H2 * W2);
                }
                if (fast) {
                    CEL_PREFETCH_L1(mi.coefs + coff);
                    CEL_PREFETCH_L1(mi.coefs + coff + mi.n1 + 3);
                }
            }
            // this lane's half (PSF component kk) of the mixture sums of both pixels (columns 2t and 2t + 1)
            // (populate_gal_fsm!, fsm_util.jl:194-219); S = F_dev F_exp | AX1 AX2 AS1 AS2 AS3 (theta-weighted)
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
            // the partner's half of MY pixel (lane kk finishes the pixel of column 2 t + kk)
            double T[NS];
#pragma unroll
            for (int q = 0; q < NS; ++q)
                T[q] = (kk == 0 ? S[0][q] : S[1][q]) + __shfl_xor_sync(0xffffffffu, kk == 0 ? S[1][q] : S[0][q], 1);

            unsigned char bit = 0;
            float xf = 0.f, skyf = 0.f;
            double pconst = 0.0, bE = 0.0, bV = 0.0;
            double f0 = 0.0, g0[2] = {0.0, 0.0}, h0[3] = {0.0, 0.0, 0.0};
            if (own) {
                bit = mi.bitmap[pix];
                xf = mi.pixels[ipix];
                skyf = mi.sky[ipix];
                pconst = mi.pixconst[ipix];
                if (mi.bg) {
                    bE = mi.bg[pix];
                    bV = mi.bg[pix + H2 * W2];
                }
                // ... and star (star_light_density!, fsm_util.jl:225-248): two new columns of the sliding window
                if (fast) {
                    const double* ccol = mi.coefs + coff;
                    const int n1 = mi.n1;
                    const double q0 = __ldg(ccol), q1 = __ldg(ccol + 1), q2 = __ldg(ccol + 2), q3 = __ldg(ccol + 3);
                    const double q4 = __ldg(ccol + n1), q5 = __ldg(ccol + n1 + 1), q6 = __ldg(ccol + n1 + 2), q7 = __ldg(ccol + n1 + 3);
                    const double R2 = si[SI_WX] * q0 + si[SI_WX + 1] * q1 + si[SI_WX + 2] * q2 + si[SI_WX + 3] * q3;
                    const double R3 = si[SI_WX] * q4 + si[SI_WX + 1] * q5 + si[SI_WX + 2] * q6 + si[SI_WX + 3] * q7;
                    const double wy0 = si[SI_WY], wy1 = si[SI_WY + 1], wy2 = si[SI_WY + 2], wy3 = si[SI_WY + 3];
                    const double v = wy0 * R0 + wy1 * R1 + wy2 * R2 + wy3 * R3;
                    double gx = 0.0, gy = 0.0;
                    if (MODE >= 1) {
                        const double D2 = si[SI_DWX] * q0 + si[SI_DWX + 1] * q1 + si[SI_DWX + 2] * q2 + si[SI_DWX + 3] * q3;
                        const double D3 = si[SI_DWX] * q4 + si[SI_DWX + 1] * q5 + si[SI_DWX + 2] * q6 + si[SI_DWX + 3] * q7;
                        gx = wy0 * D0 + wy1 * D1 + wy2 * D2 + wy3 * D3;
                        gy = si[SI_DWY] * R0 + si[SI_DWY + 1] * R1 + si[SI_DWY + 2] * R2 + si[SI_DWY + 3] * R3;
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
                    } else {
                        f0 = 1e-3 * (v + 1.0);
                        g0[0] = 1e-3 * gx;
                        g0[1] = 1e-3 * gy;
                    }
                }
            }
            if (own && bit && !isnan(xf)) {                          // elbo_objective.jl:445, :459
                PixelConsts pc;
                pc.x = (double)xf;
                pc.iota = (double)mi.iota[h - 1];
                pc.pixconst = pconst;
                GalRaw gal;
                const bool covered = iown < ncov;      // the last column of the patch is not covered by its own source (:349)
                gal.f = theta * T[0] + (1.0 - theta) * T[1];
                if (MODE >= 1) {
                    gal.r[0] = -T[2];
                    gal.r[1] = -T[3];
                    gal.r[2] = 0.5 * T[4];
                    gal.r[3] = T[5];
                    gal.r[4] = 0.5 * T[6];
                    gal.r[5] = T[0] - T[1];                               // gal_frac_dev, fsm_util.jl:277-291
                }
                if (covered) {
                    if (!fast)
                        star_eval<MODE>(LdGlobal(), mi.coefs, mi.n1, mi.n2, (double)h - si[SI_M] + 26.0,
                                        (double)(w0 + iown) - si[SI_M + 1] + 26.0, f0, g0, h0);
                    cnt_active += 1.0;
                }
                const double cb[4] = {si[SI_CB], si[SI_CB + 1], si[SI_CB + 2], si[SI_CB + 3]};
                double la[NUA];
#pragma unroll
                for (int q = 0; q < NUA; ++q) la[q] = 0.0;
                pixel_accumulate<MODE>(la, 1, pc, (double)skyf + bE, bV, covered, true, cb, f0, g0, h0, gal, s_logtab);
                val += la[ACC_VAL];
                if (MODE >= 1 && covered) {
                    // (c, y) space of this image -> task space: dx_a/dpos_b = -J[a + 2 b]; the c-scalars go to their band
                    ta[(TA_POS + 0) * MARCH_THREADS] -= si[SI_J + 0] * la[ACC_G] + si[SI_J + 1] * la[ACC_G + 1];
                    ta[(TA_POS + 1) * MARCH_THREADS] -= si[SI_J + 2] * la[ACC_G] + si[SI_J + 3] * la[ACC_G + 1];
#pragma unroll
                    for (int q = 0; q < 3; ++q) ta[(TA_SIG + q) * MARCH_THREADS] += la[ACC_G + 2 + q];
                    ta[TA_THETA * MARCH_THREADS] += la[ACC_G + 5];
                    double* tb = ta + (TA_BAND + 4 * mi.band0) * MARCH_THREADS;
#pragma unroll
                    for (int q = 0; q < 4; ++q) tb[q * MARCH_THREADS] += la[ACC_C1 + q];
                }
            }
            pix += 2 * H2;
            ipix += 2 * mi.imgH;
            coff += 2 * mi.n1;
        }
        }
        ta[TA_VAL * MARCH_THREADS] += val;
        ta[TA_CNT_ACTIVE * MARCH_THREADS] += cnt_active;
    }
    __syncthreads();

    // fixed-order block reduction, all of a warp's accumulators in flight
    const int warp = tid >> 5;
    constexpr int NA = MODE == 0 ? TA_POS : NT_ACC;
    constexpr int NW = MARCH_THREADS / 32, PER = (NA + NW - 1) / NW;
    double* out = plan.partials + (size_t)th.pidx * NT_ACC;
    double sred[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int a = warp + NW * i;
        sred[i] = 0.0;
        if (a < NA) {
#pragma unroll
            for (int q = 0; q < NW; ++q) sred[i] += tacc[a * MARCH_THREADS + lane + 32 * q];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < PER; ++i) sred[i] += __shfl_xor_sync(0xffffffffu, sred[i], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (warp + NW * i < NA) out[warp + NW * i] = sred[i];
    }
}

// Host side: the blocks of a plan.  One block per (sub, group of <= MARCH_NIMG images); a source whose patches hold
// more than `split_pixels` pixels is cut into smaller image groups (down to one block per image) so that no single
// block is a long tail of the launch.
// Each block owns one NT_ACC-vector of plan.partials (pidx); part_ptr[sub] .. part_ptr[sub + 1] are the blocks of a
// sub, in image order.  geo(slot, n, off_h, off_w, H2, W2) -> the patch box of a slot in image n.
// nseg (column segments per row) minimises  block-rounds x (iterations of two columns + an exact start, ~0.6
// iterations): all warps of a block then leave the walk loop together instead of idling at its final barrier.
template <typename Geo>
inline void build_march_blocks(int n_subs, int N, const int* sub_task, const int* sub_slot, const int* task_ptr,
                               const int* task_field, Geo geo, long split_pixels, std::vector<MarchHdr>& blocks,
                               std::vector<int>& part_ptr) {
    blocks.clear();
    part_ptr.assign((size_t)n_subs + 1, 0);
    std::vector<long> cost;
    for (int u = 0; u < n_subs; ++u) {
        const int t = sub_task[u];
        const int aslot = sub_slot[u], slot0 = task_ptr[t], slot1 = task_ptr[t + 1];
        long tot = 0;
        for (int n = 0; n < N; ++n) {
            int oh, ow, H2, W2;
            geo(aslot, n, oh, ow, H2, W2);
            tot += (long)std::max(H2, 0) * std::max(W2, 0);
        }
        // images per block: all of them (<= MARCH_NIMG) unless the source is heavier than split_pixels; then as many
        // as keep a block under that limit (at least one)
        int step = MARCH_NIMG;
        if (tot > split_pixels && N > 0) {
            const long per_image = std::max(1L, tot / N);
            step = (int)std::max(1L, std::min((long)MARCH_NIMG, split_pixels / per_image));
        }
        for (int n0 = 0; n0 < N; n0 += step) {
            MarchHdr th{};
            th.aslot = aslot;
            th.slot0 = slot0;
            th.slot1 = slot1;
            th.field = task_field ? task_field[t] : 0;
            th.sub = u;
            th.task = t;
            th.n0 = n0;
            th.n1 = std::min(N, n0 + step);
            th.pidx = (int)blocks.size();
            int rows = 0, maxw = 0;
            long c = 0;
            th.hasbg = 0;
            for (int n = th.n0; n < th.n1; ++n) {
                int oh, ow, H2, W2;
                geo(aslot, n, oh, ow, H2, W2);
                if (H2 <= 0 || W2 <= 0) continue;
                rows += H2;
                maxw = std::max(maxw, W2);
                c += (long)H2 * W2 * (1 + (slot1 - slot0 - 1) / 4);
                for (int s = slot0; s < slot1; ++s) {
                    if (s == aslot) continue;
                    int ph, pw, pH2, pW2;
                    geo(s, n, ph, pw, pH2, pW2);
                    // rows off+1..off+H2 of both; columns off+1..off+W2 of the active patch, off+1..off+W2-1 of the
                    // neighbour (strict `w2 < W2`, elbo_objective.jl:349)
                    if (pH2 > 0 && pW2 > 1 && (ph + 1 <= oh + H2) && (ph + pH2 >= oh + 1) && (pw + 1 <= ow + W2) &&
                        (pw + pW2 - 1 >= ow + 1))
                        th.hasbg |= 1u << (n - th.n0);
                }
            }
            const int nmin = std::max(1, (maxw + MARCH_MAXSEG - 1) / MARCH_MAXSEG);
            long best = -1;
            th.nseg = nmin;
            for (int cand = nmin; cand < nmin + 6; ++cand) {
                const int L = (maxw + cand - 1) / cand;
                const long cst = (long)((rows * cand + NPAIR - 1) / NPAIR) * (10 * ((L + 1) / 2) + 6);
                if (best < 0 || cst < best) {
                    best = cst;
                    th.nseg = cand;
                }
            }
            th.ubeg[0] = 0;
            for (int k = 0; k < MARCH_NIMG; ++k) {
                int units = 0;
                if (th.n0 + k < th.n1) {
                    int oh, ow, H2, W2;
                    geo(aslot, th.n0 + k, oh, ow, H2, W2);
                    if (H2 > 0 && W2 > 0) units = H2 * th.nseg;
                }
                th.ubeg[k + 1] = th.ubeg[k] + units;
            }
            blocks.push_back(th);
            cost.push_back(c);
        }
        part_ptr[u + 1] = (int)blocks.size();
    }
    // heaviest first, so the launch tail is short; pidx keeps the per-sub order of the partial vectors
    std::vector<int> order(blocks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<MarchHdr> sorted(blocks.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = blocks[order[i]];
    blocks.swap(sorted);
}

constexpr size_t march_smem_bytes() {
    return ((size_t)NT_ACC * MARCH_THREADS + (size_t)MARCH_NIMG * NC2 * MREC + (size_t)MARCH_NIMG * SI_STRIDE) * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// The rest of the chain rule, once per source: task space -> the 28 live canonical parameters
// (what epilogue_kernel's Jy does per image).  One warp per task.
constexpr int MEPI_THREADS = 64;

template <int MODE>
__global__ void __launch_bounds__(MEPI_THREADS) march_epilogue_kernel(PlanDev plan, const double* __restrict__ vp,
                                                                      const int* __restrict__ part_ptr,
                                                                      double* __restrict__ out_v, double* __restrict__ out_d,
                                                                      long long* __restrict__ out_counters,
                                                                      int* __restrict__ out_flags) {
    __shared__ double ysum[NT_ACC];
    __shared__ double J0[3][3], T0[3][3][3];
    __shared__ double s_El[2][5], s_Ell[2][5];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int t = blockIdx.x;
    if (plan.task_mask && !plan.task_mask[t]) return;
    const int sub = plan.sub_ptr[t];                 // Sa == 1
    const int aslot = plan.sub_slot[sub];
    const double* vs = vp + (size_t)NPARAM * aslot;
    constexpr int NA = MODE == 0 ? TA_POS : NT_ACC;
    if (tid == 33 && MODE >= 1) brightness_values(vs, s_El, s_Ell);
    if (tid < NA) {
        double s = 0.0;
        for (int g = part_ptr[sub]; g < part_ptr[sub + 1]; ++g) s += plan.partials[(size_t)g * NT_ACC + tid];   // fixed order
        ysum[tid] = s;
    }
    if (tid == 32 && MODE >= 1) sigma_derivs(vs[3], vs[4], vs[5], J0, T0);
    if (tid == 0) s_bad = 0;
    __syncthreads();
    int bad = 0;
    if (MODE >= 1 && tid < NPARAM) {
        const int q = tid;
        double g = 0.0;
        int i, k;
        if (q < 2) {
            g = ysum[TA_POS + q];
        } else if (q == 2) {
            g = ysum[TA_THETA];
        } else if (q < 6) {
            for (int r = 0; r < 3; ++r) g += J0[r][q - 3] * ysum[TA_SIG + r];
        } else if (bright_of(q, i, k)) {
            // c = (a1 E_l1, a2 E_l2, a1 E_ll1, a2 E_ll2) of each band; every derivative is E * kappa (source_brightness.jl:45-202)
            for (int b = 0; b < 5; ++b) {
                double ka[10], la[10];
                band_coefs(b, ka, la);
                g += vs[26 + i] * (s_El[i][b] * ka[k] * ysum[TA_BAND + 4 * b + i] +
                                   s_Ell[i][b] * la[k] * ysum[TA_BAND + 4 * b + 2 + i]);
            }
        } else if (q == 26 || q == 27) {
            i = q - 26;
            for (int b = 0; b < 5; ++b)
                g += s_El[i][b] * ysum[TA_BAND + 4 * b + i] + s_Ell[i][b] * ysum[TA_BAND + 4 * b + 2 + i];
        }
        out_d[(size_t)NPARAM * sub + q] = g;
        bad |= !isfinite(g);
    }
    if (tid == 0) {
        out_v[t] = ysum[TA_VAL];
        out_counters[2 * t] = (long long)(ysum[TA_CNT_ACTIVE] + 0.5);
        out_counters[2 * t + 1] = (long long)(ysum[TA_CNT_INACTIVE] + 0.5);
        bad |= !isfinite(ysum[TA_VAL]);
    }
    if (bad) atomicOr(&s_bad, 1);
    __syncthreads();
    if (tid == 0) out_flags[t] = s_bad ? 1 : 0;
}

}  // namespace celeste
#endif
