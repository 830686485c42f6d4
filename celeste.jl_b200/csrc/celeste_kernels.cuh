// celeste_kernels.cuh -- sm_100a kernels of the ELBO hot path (see DESIGN.md "Kernels").
//
//   prep_image_kernel   once per image: pixconst = x log(iota) - lgamma(x + 1)   (elbo_objective.jl:292,391)
//   setup_kernel        per evaluation: mixtures of every (task source, image)   (fsm_util.jl:111-169)
//                       + per-source brightness moments                          (source_brightness.jl:27-202)
//   task_kernel<MODE>   THE hot loop, value / gradient modes: one block per (active source, <= 5 images),
//                       warps never synchronise                                  (elbo_objective.jl:330-470)
//   pixel_kernel<2>     THE hot loop, Hessian mode: one block per (active source, image, pixel chunk)
//   pair_kernel         Sa > 1 only: cross-source Hessian blocks                 (SensitiveFloats.jl:114-126)
//   epilogue_kernel<MODE> per task: fixed-order reduction of the chunk partials, raw -> parameter
//                       chain rule, 44 x 44 SensitiveFloat output                (SensitiveFloats.jl:23-47)
//
// FP64-pipe bound by design: the pixel kernel touches ~20 B of HBM per pixel-visit
// and executes ~1.5k (grad) / ~4k (Hessian) FP64 instructions for it; everything a thread needs
// per component comes from shared memory (active source) or L1-resident broadcasts (neighbours).
#ifndef CELESTE_KERNELS_CUH
#define CELESTE_KERNELS_CUH

#ifndef CELESTE_HOST_EMULATION   // tests/host_emul compiles this file with a host emulation layer instead
#include <cuda_runtime.h>
#define CEL_DYNAMIC_SMEM(name) extern __shared__ __align__(16) double name[]
#else
#define CEL_DYNAMIC_SMEM(name) double* name = ::cuda_emul::dynamic_smem()
#endif

#include <algorithm>
#include <vector>

#include "elbo_math.cuh"

namespace celeste {

struct ImageDev {
    int H, W, band;
    const float* pixels;
    const float* sky;
    const float* iota;
    const double* pixconst;   // x*log_iota[h] - lgamma(x+1)
};

struct PatchDev {
    int off_h, off_w, H2, W2;   // bitmap_offset, size(active_pixel_bitmap)
    const uint8_t* bitmap;
    double J[4];                // wcs_jacobian, col-major
    double wc[2], pc[2];
    int K;
    const double* psf;          // K x 7
    const double* coefs;        // padded spline coefficients
    int n1, n2;
};

// per (task source slot, image) scratch written by setup_kernel
constexpr int SLOTIMG_STRIDE = MAX_COMPS * COMP_STRIDE + 2;   // comps + m_pos
// per slot: El[2][5], Ell[2][5], a[2], theta
constexpr int SLOTBR_STRIDE = 24;

struct FieldDev {
    const ImageDev* images;
    const PatchDev* patches;  // s + n * S_tot
    int S_tot;
    int pad;
};

// everything a pixel-kernel block needs to find its work, in one 32-byte load
struct BlockHdr {
    int tn;        // sub * N + n   (sub = one active source of one task)
    int chunk;     // chunk index inside the active patch
    int aslot;     // slot of this active source
    int slot0, slot1;   // slot range of the task
    int patch;     // index of the active source's patch in FieldDev::patches
    int n;         // image
    int field;     // index into PlanDev::fields
    int sub0, sub, sub1;   // this task's active sources are subs [sub0, sub1); this block works for `sub`
    int task;
};

// one block of pair_kernel: the cross-source Hessian block of two active sources of one task in one image
struct PairHdr {
    int sub_a, sub_b;      // sub_a < sub_b (order of ea.active_sources)
    int slot_a, slot_b;
    int slot0, slot1;
    int n, field;
};
constexpr int NPAIR_ACC = 100;   // (c, y)_a x (c, y)_b

struct PixRec;
struct PlanDev {
    int n_tasks, N, n_fields, n_slots, n_subs, n_pairs;
    const FieldDev* fields;  // n_fields inference boxes share one plan (same N)
    const int* task_field;   // n_tasks: field of each task
    const int* slot_field;   // n_slots: field of each slot
    const int* task_ptr;     // n_tasks + 1 (slot ranges)
    const int* src_row;      // n_slots: 0-based patch row of each slot (inside its field)
    const int* sub_ptr;      // n_tasks + 1: active sources ("subs") of each task, ea.active_sources order
    const int* sub_slot;     // n_subs: slot of each active source
    const long long* h_ptr;  // n_tasks + 1: offset of each task's (44 Sa)^2 Hessian in the output
    const BlockHdr* blockmap; // n_blocks
    const int* chunk_ptr;    // n_subs * N + 1: first block of (sub, n)
    const PairHdr* pairmap;  // n_pairs * N
    const int* pair_ptr;     // n_tasks + 1: first pair of each task
    double* pair_partials;   // n_pairs * N * NPAIR_ACC
    const unsigned char* task_mask;   // optional (may be null): tasks with mask 0 are skipped, their outputs left untouched
    double* slotimg;         // n_slots * N * SLOTIMG_STRIDE
    double* slotbr;          // n_slots * SLOTBR_STRIDE
    double* partials;        // n_blocks * NACC
    const long long* bg_ptr; // n_subs * N: offset of the (E_bg, V_bg) planes of (sub, image) in bg, -1 if the task has no neighbour
    double* bg;              // march_kernel's / unit kernels' per-(sub, image) background planes (E_bg, V_bg)
    double* bg_cnt;          // neighbour pixel-visits counted by each piece of unit_bg_kernel
    const long long* l5_ptr; // n_subs * N: first pixel of (sub, image) in pix / l5 (unit kernels; walk order)
    double* l5;              // L5 = dL/df1 of every pixel of every unit (Hessian mode: phase A -> phase B)
    const struct PixRec* pix; // packed pixel records of every unit (unit_pack_kernel)
};


__constant__ double c_proto_eta[NPROTO];
__constant__ double c_proto_nu[NPROTO];

struct LdGlobal {
    __device__ __forceinline__ double operator()(const double* p) const { return __ldg(p); }
};
struct LdShared {
    __device__ __forceinline__ double operator()(const double* p) const { return *p; }
};

// ------------------------------------------------------------------------------------------------
__global__ void prep_image_kernel(int H, int W, const float* __restrict__ pixels, const float* __restrict__ iota,
                                  const double* __restrict__ log_iota, double* __restrict__ pixconst) {
    const size_t n = (size_t)H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % H);
        const double x = (double)pixels[i];
        const double li = log_iota ? log_iota[h] : (double)logf(iota[h]);   // Float32 log, elbo_objective.jl:292
        pixconst[i] = x * li - lgamma(x + 1.0);
    }
}

// one thread per (slot, image, PSF component): the 14 prototype components of that PSF component share the
// source's XiXi (one sin / cos per thread instead of one per component)
__global__ void setup_kernel(PlanDev plan, const double* __restrict__ vp) {
    const long total = (long)plan.n_slots * plan.N * MAX_K;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % MAX_K);
        const long sn = idx / MAX_K;
        const int n = (int)(sn % plan.N);
        const int slot = (int)(sn / plan.N);
        const double* vs = vp + (size_t)NPARAM * slot;
        const FieldDev& field = plan.fields[plan.slot_field[slot]];
        const PatchDev& p = field.patches[plan.src_row[slot] + (size_t)n * field.S_tot];
        double* rec = plan.slotimg + ((size_t)slot * plan.N + n) * SLOTIMG_STRIDE;
        // linear_world_to_pix, wcs_utils.jl:14-18
        const double d0 = vs[0] - p.wc[0], d1 = vs[1] - p.wc[1];
        const double m1 = (p.J[0] * d0 + p.J[2] * d1) + p.pc[0];
        const double m2 = (p.J[1] * d0 + p.J[3] * d1) + p.pc[1];
        if (k < p.K) {
            double x11, off, x22;
            galaxy_xixi(vs[3], vs[4], vs[5], x11, off, x22);
            const double* psf7 = p.psf + 7 * k;
            for (int j = 0; j < NPROTO; ++j) {
                const int c = j * p.K + k;
                make_component_xi(psf7, c_proto_eta[j], c_proto_nu[j], m1, m2, x11, off, x22, rec + c * COMP_STRIDE);
            }
        }
        if (k == 0) {
            rec[MAX_COMPS * COMP_STRIDE + 0] = m1;
            rec[MAX_COMPS * COMP_STRIDE + 1] = m2;
            if (n == 0) {
                double El[2][5], Ell[2][5];
                brightness_values(vs, El, Ell);
                double* br = plan.slotbr + (size_t)slot * SLOTBR_STRIDE;
                for (int i = 0; i < 2; ++i)
                    for (int b = 0; b < 5; ++b) {
                        br[i * 5 + b] = El[i][b];
                        br[10 + i * 5 + b] = Ell[i][b];
                    }
                br[20] = vs[26];
                br[21] = vs[27];
                br[22] = vs[2];
                br[23] = 0.0;
            }
        }
    }
}

// Brightness moments (source_brightness.jl:27-202) of EVERY slot of the plan, one thread per slot: what epilogue_kernel
// reads for the active sources and what the unit kernels read for every source they walk (active or neighbour).  The
// kernels that build their own mixtures (unit kernels) need nothing else from setup_kernel.
__global__ void slotbr_kernel(PlanDev plan, const double* __restrict__ vp, int* __restrict__ queue) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < 4 && queue) queue[slot] = 0;              // the unit kernels' work-queue counters of this evaluation
    if (slot >= plan.n_slots) return;
    const double* vs = vp + (size_t)NPARAM * slot;
    double El[2][5], Ell[2][5];
    brightness_values(vs, El, Ell);
    double* br = plan.slotbr + (size_t)slot * SLOTBR_STRIDE;
    for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 5; ++b) {
            br[i * 5 + b] = El[i][b];
            br[10 + i * 5 + b] = Ell[i][b];
        }
    br[20] = vs[26];
    br[21] = vs[27];
    br[22] = vs[2];
    br[23] = 0.0;
}

// ------------------------------------------------------------------------------------------------
#ifndef CELESTE_PIX_THREADS
#define CELESTE_PIX_THREADS 128
#endif
constexpr int PIX_THREADS = CELESTE_PIX_THREADS;   // threads per pixel-kernel block (multiple of 32)
constexpr int MAX_NB_LIST = 64;

#ifndef CELESTE_PIX_MINB_GRAD
#define CELESTE_PIX_MINB_GRAD 6
#endif
#ifndef CELESTE_PIX_MINB_HESS
#define CELESTE_PIX_MINB_HESS 3
#endif
// KT: PSF components per patch fixed at compile time (2 = the reference default psf_K, elbo_args.jl:197);
// KT == 0 reads K from each patch at run time.
// MULTI: some task of the plan has Sa > 1 (unit tests); false compiles the `already_visited` logic away.
template <int MODE, int KT, bool MULTI>
__global__ void __launch_bounds__(PIX_THREADS, MODE == 2 ? CELESTE_PIX_MINB_HESS : CELESTE_PIX_MINB_GRAD)
    pixel_kernel(PlanDev plan, int chunk_pixels) {
    constexpr int NACC = NAcc<MODE>::value;
    CEL_DYNAMIC_SMEM(smem);
    double* acc = smem;                                    // NACC x PIX_THREADS
    double* s_comps = acc + NACC * PIX_THREADS;           // MAX_COMPS * 6
    __shared__ double s_exptab[8];
    __shared__ int s_nb[MAX_NB_LIST];
    __shared__ int s_nb_count;
    __shared__ int s_nb_overflow;

    const int tid = threadIdx.x;
    const BlockHdr bm = plan.blockmap[blockIdx.x];
    if (plan.task_mask && !plan.task_mask[bm.task]) return;   // e.g. a source whose Newton iteration has converged
    const int chunk = bm.chunk, n = bm.n;
    const int slot0 = bm.slot0, slot1 = bm.slot1;
    const int aslot = bm.aslot;
    const int sub0 = bm.sub0, sub = bm.sub, sub1 = bm.sub1;
    const bool multi = MULTI && (sub1 - sub0) > 1;         // Sa > 1: unit tests only (elbo_objective.jl:429-434)
    const FieldDev field = plan.fields[bm.field];
    const ImageDev img = field.images[n];
    const PatchDev pa = field.patches[bm.patch];
    const int npix = pa.H2 * pa.W2;
    const int first = chunk * chunk_pixels;
    const int last = min(first + chunk_pixels, npix);      // exclusive

    // value / gradient modes keep their 3 / 13 accumulators in registers; the 68 of the Hessian mode live in
    // per-thread shared-memory slots
#ifndef CELESTE_REG_ACC
#define CELESTE_REG_ACC 0   // measured: register accumulators spill at 80 regs and lose 4% (profiles/tuning_r01.md)
#endif
    constexpr bool REG_ACC = (MODE <= 1) && (CELESTE_REG_ACC != 0);
    double racc[REG_ACC ? NACC : 1];
    if (REG_ACC) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) racc[a] = 0.0;
    } else {
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a * PIX_THREADS + tid] = 0.0;
    }
    double* const my_acc = REG_ACC ? racc : acc + tid;
    constexpr int ACC_STRIDE = REG_ACC ? 1 : PIX_THREADS;
    const double* arec = plan.slotimg + ((size_t)aslot * plan.N + n) * SLOTIMG_STRIDE;
    for (int i = tid; i < NPROTO * pa.K * COMP_STRIDE; i += PIX_THREADS) s_comps[i] = arec[i];
    if (tid == 0) {
        s_nb_count = 0;
        s_nb_overflow = 0;
    }
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    __syncthreads();

    // neighbours whose patch in this image can touch this chunk's pixels (ordered compaction by warp 0,
    // so the summation order over neighbours -- hence the result -- is deterministic)
    if (tid < 32 && last > first) {
        const int w2_lo = first / pa.H2, w2_hi = (last - 1) / pa.H2;
        int h2_lo = 0, h2_hi = pa.H2 - 1;
        if (w2_lo == w2_hi) {
            h2_lo = first % pa.H2;
            h2_hi = (last - 1) % pa.H2;
        }
        const int bh_lo = pa.off_h + h2_lo + 1, bh_hi = pa.off_h + h2_hi + 1;   // 1-based image rows
        const int bw_lo = pa.off_w + w2_lo + 1, bw_hi = pa.off_w + w2_hi + 1;
        for (int base = slot0; base < slot1; base += 32) {
            const int s = base + tid;
            bool hit = false;
            if (s < slot1 && s != aslot) {
                const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
                // rows off_h+1 .. off_h+H2, columns off_w+1 .. off_w+W2-1 (strict `w2 < W2`, elbo_objective.jl:349)
                hit = (p.off_h + 1 <= bh_hi) && (p.off_h + p.H2 >= bh_lo) && (p.off_w + 1 <= bw_hi) &&
                      (p.off_w + p.W2 - 1 >= bw_lo);
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            const int pos = s_nb_count + __popc(ballot & ((1u << tid) - 1u));
            if (hit) {
                if (pos < MAX_NB_LIST)
                    s_nb[pos] = s;
                else
                    s_nb_overflow = 1;
            }
            __syncwarp();
            if (tid == 0) s_nb_count = min(s_nb_count + __popc(ballot), MAX_NB_LIST);
            __syncwarp();
        }
    }
    __syncthreads();

    const double* abr = plan.slotbr + (size_t)aslot * SLOTBR_STRIDE;
    const int b = img.band - 1;
    const double a1 = abr[20], a2 = abr[21], theta = abr[22];
    const double cb[4] = {a1 * abr[b], a2 * abr[5 + b], a1 * abr[10 + b], a2 * abr[15 + b]};
    const double am1 = arec[MAX_COMPS * COMP_STRIDE], am2 = arec[MAX_COMPS * COMP_STRIDE + 1];

    // value-only contribution of neighbour slot s at image pixel (h, w) (elbo_objective.jl:342-372, inactive branch)
    double cnt_other_active = 0.0;
    auto neighbour = [&](int s, int h, int w, double& Ebg, double& Vbg, double& cnt) {
        const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
        const int h2 = h - p.off_h, w2 = w - p.off_w;
        if (h2 < 1 || h2 > p.H2 || w2 < 1 || w2 >= p.W2) return;
        if (!p.bitmap[(h2 - 1) + (size_t)(w2 - 1) * p.H2]) return;
        bool other_active = false;
        if (multi)
            for (int j = sub0; j < sub1; ++j) other_active |= (plan.sub_slot[j] == s);
        if (other_active)
            cnt_other_active += 1.0;     // counted as an ACTIVE pixel-visit (elbo_objective.jl:353-357)
        else
            cnt += 1.0;
        const double* rec = plan.slotimg + ((size_t)s * plan.N + n) * SLOTIMG_STRIDE;
        const double* br = plan.slotbr + (size_t)s * SLOTBR_STRIDE;
        const double m1 = __ldg(rec + MAX_COMPS * COMP_STRIDE), m2 = __ldg(rec + MAX_COMPS * COMP_STRIDE + 1);
        double f0, gd[2], hd[3];
        star_eval<0>(LdGlobal(), p.coefs, p.n1, p.n2, (double)h - m1 + 26.0, (double)w - m2 + 26.0, f0, gd, hd);
        const double f1 = gal_value<KT>(LdGlobal(), rec, p.K, s_exptab, __ldg(br + 22), (double)h, (double)w);
        const double na1 = __ldg(br + 20), na2 = __ldg(br + 21);
        const double Es = na1 * __ldg(br + b) * f0 + na2 * __ldg(br + 5 + b) * f1;
        const double E2s = na1 * __ldg(br + 10 + b) * f0 * f0 + na2 * __ldg(br + 15 + b) * f1 * f1;
        Ebg += Es;
        Vbg += E2s - Es * Es;
    };

    // per-pixel inputs are fetched one iteration ahead so their DRAM latency hides behind the FP64 work
    // (all five loads are issued unconditionally -- the patch box is clamped to the image -- and nothing
    // is tested until the values are consumed one iteration later)
    struct PixIn {
        unsigned char active;
        float x, sky, iota;
        double pixconst;
    };
    auto fetch = [&](int pix) {
        PixIn in;
        in.active = 0;
        in.x = in.sky = in.iota = 0.f;
        in.pixconst = 0.0;
        if (pix < last) {
            const int h2 = pix % pa.H2, w2 = pix / pa.H2;
            const size_t ipix = (size_t)(pa.off_h + h2) + (size_t)(pa.off_w + w2) * img.H;
            in.active = pa.bitmap[pix];
            in.x = img.pixels[ipix];
            in.sky = img.sky[ipix];
            in.iota = img.iota[pa.off_h + h2];
            in.pixconst = img.pixconst[ipix];
        }
        return in;
    };
    PixIn cur = fetch(first + tid);
    for (int pix = first + tid; pix < last; pix += PIX_THREADS) {
        const PixIn nxt = fetch(pix + PIX_THREADS);
        if (cur.active && !isnan(cur.x)) {                       // elbo_objective.jl:445, :459
            const int h2 = pix % pa.H2, w2 = pix / pa.H2;        // 0-based local
            const int h = pa.off_h + h2 + 1, w = pa.off_w + w2 + 1;   // 1-based image coordinates
            PixelConsts pc;
            pc.x = (double)cur.x;
            pc.iota = (double)cur.iota;
            pc.pixconst = cur.pixconst;
            double Ebg = (double)cur.sky;                        // :374
            double Vbg = 0.0;
            double cnt_inactive = 0.0;
            cnt_other_active = 0.0;
            // Sa > 1: a pixel is visited once, by the first active source whose patch holds it
            // (`already_visited`, elbo_objective.jl:450-455); later actives still take its derivatives
            bool first_visit = true;
            if (multi)
                for (int j = sub0; j < sub; ++j) {
                    const PatchDev& pj = field.patches[plan.src_row[plan.sub_slot[j]] + (size_t)n * field.S_tot];
                    const int hj = h - pj.off_h, wj = w - pj.off_w;
                    if (hj >= 1 && hj <= pj.H2 && wj >= 1 && wj <= pj.W2 && pj.bitmap[(hj - 1) + (size_t)(wj - 1) * pj.H2])
                        first_visit = false;
                }
            const int nnb = s_nb_count;
            for (int i = 0; i < nnb; ++i) neighbour(s_nb[i], h, w, Ebg, Vbg, cnt_inactive);
            if (s_nb_overflow) {
                // rare: more overlapping neighbours than the list holds; scan the remaining slots in order
                const int last_listed = s_nb[MAX_NB_LIST - 1];
                for (int s = last_listed + 1; s < slot1; ++s)
                    if (s != aslot) neighbour(s, h, w, Ebg, Vbg, cnt_inactive);
            }
            const bool covered = (w2 + 1) < pa.W2;               // strict last column, :349
            double f0 = 0.0, g0[2] = {0.0, 0.0}, h0[3] = {0.0, 0.0, 0.0};
            GalRaw gal;
            gal.f = 0.0;
            if (covered) {
                star_eval<MODE>(LdGlobal(), pa.coefs, pa.n1, pa.n2, (double)h - am1 + 26.0, (double)w - am2 + 26.0, f0,
                                g0, h0);
                gal_eval<MODE, KT>(LdShared(), s_comps, pa.K, c_proto_nu, s_exptab, theta, (double)h, (double)w, gal);
            }
            if (first_visit) {
                my_acc[ACC_CNT_ACTIVE * ACC_STRIDE] += (covered ? 1.0 : 0.0) + cnt_other_active;
                my_acc[ACC_CNT_INACTIVE * ACC_STRIDE] += cnt_inactive;
            }
            pixel_accumulate<MODE>(my_acc, ACC_STRIDE, pc, Ebg, Vbg, covered, first_visit, cb, f0, g0, h0, gal);
        }
        cur = nxt;
    }
    if (REG_ACC) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a * PIX_THREADS + tid] = racc[a];
    }
    __syncthreads();

    // fixed-order block reduction: warp w reduces accumulators a = w, w + 4, ...
    const int warp = tid >> 5, lane = tid & 31;
    double* out = plan.partials + (size_t)blockIdx.x * NACC;
    // (all of a warp's accumulators in flight at once: the block is short, so this tail must not be a chain of
    //  dependent shared-memory loads and shuffles per accumulator)
    constexpr int NW = PIX_THREADS / 32, PER = (NACC + NW - 1) / NW;
    double s[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int a = warp + NW * i;
        s[i] = 0.0;
        if (a < NACC) {
#pragma unroll
            for (int k = 0; k < NW; ++k) s[i] += acc[a * PIX_THREADS + lane + 32 * k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < PER; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (warp + NW * i < NACC) out[warp + NW * i] = s[i];
    }
}

// ------------------------------------------------------------------------------------------------
// task_kernel<MODE <= 1>: value / gradient modes.  One block per (active source of a task, group of <= TASK_NIMG
// images).  A typical patch is ~400 pixels = 12.5 warp-iterations, which a 4-warp block cannot split evenly
// (3.25 per warp: a quarter of the block idles at the final barrier in pixel_kernel).  Here the warps of a block
// walk through ALL images of the task without ever synchronising: the warp-iterations of image n are dealt
// round-robin starting where image n-1 stopped, each warp reduces its own accumulators by shuffles when it leaves
// an image and writes one partial per (sub, image, warp).  The per-block prologue (header, component records,
// neighbour lists) is paid once per task instead of once per image.  The Hessian mode keeps pixel_kernel: its 68
// accumulator slots per thread leave no shared memory for five images' component records.
constexpr int TASK_NIMG = 5;
constexpr int TASK_WARPS = PIX_THREADS / 32;
constexpr int TASK_NB = 32;          // neighbour-list capacity per image (overflow falls back to a slot scan)

struct TaskHdr {
    int tn0;             // sub * N  (partials of (sub, n, warp) live at ((tn0 + n) * TASK_WARPS + warp))
    int aslot, slot0, slot1;
    int field;
    int sub0, sub, sub1;
    int n0, n1;          // image range of this block
    int task, pad1;
};

template <int MODE, int KT, bool MULTI>
__global__ void __launch_bounds__(PIX_THREADS, CELESTE_PIX_MINB_GRAD) task_kernel(PlanDev plan, const TaskHdr* __restrict__ taskmap) {
    static_assert(MODE <= 1, "the Hessian mode uses pixel_kernel");
    constexpr int NACC = NAcc<MODE>::value;
    CEL_DYNAMIC_SMEM(smem);
    double* acc = smem;                                     // NACC x PIX_THREADS
    double* s_comps = acc + NACC * PIX_THREADS;             // TASK_NIMG x MAX_COMPS x COMP_STRIDE
    __shared__ double s_exptab[8];
    __shared__ int s_nb[TASK_NIMG][TASK_NB];
    __shared__ int s_nb_count[TASK_NIMG];
    __shared__ int s_nb_overflow[TASK_NIMG];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TaskHdr th = taskmap[blockIdx.x];
    if (plan.task_mask && !plan.task_mask[th.task]) return;
    const int slot0 = th.slot0, slot1 = th.slot1, aslot = th.aslot;
    const int sub0 = th.sub0, sub = th.sub, sub1 = th.sub1;
    const bool multi = MULTI && (sub1 - sub0) > 1;
    const FieldDev field = plan.fields[th.field];
    const int nimg = th.n1 - th.n0;

#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a * PIX_THREADS + tid] = 0.0;
    {
        // component records of all images of the group: one flat, coalesced copy (only the 14 K live records when K is known)
        constexpr int per = (KT > 0 ? NPROTO * KT : MAX_COMPS) * COMP_STRIDE;   // mixed K: whole records
        for (int i = tid; i < nimg * per; i += PIX_THREADS) {
            const int k = i / per, r = i - k * per;
            s_comps[k * MAX_COMPS * COMP_STRIDE + r] =
                plan.slotimg[((size_t)aslot * plan.N + th.n0 + k) * SLOTIMG_STRIDE + r];
        }
    }
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    // neighbour lists: warp w compacts (in slot order) the sources whose patch in image n0 + k meets the active patch
    for (int k = warp; k < nimg; k += TASK_WARPS) {
        const int n = th.n0 + k;
        const PatchDev& pa = field.patches[plan.src_row[aslot] + (size_t)n * field.S_tot];
        int count = 0, overflow = 0;
        for (int base = slot0; base < slot1; base += 32) {
            const int s = base + lane;
            bool hit = false;
            if (s < slot1 && s != aslot && pa.H2 > 0 && pa.W2 > 0) {
                const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
                hit = (p.off_h + 1 <= pa.off_h + pa.H2) && (p.off_h + p.H2 >= pa.off_h + 1) &&
                      (p.off_w + 1 <= pa.off_w + pa.W2) && (p.off_w + p.W2 - 1 >= pa.off_w + 1);
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            const int pos = count + __popc(ballot & ((1u << lane) - 1u));
            if (hit && pos < TASK_NB) s_nb[k][pos] = s;
            if (count + __popc(ballot) > TASK_NB) overflow = 1;
            count = min(count + __popc(ballot), TASK_NB);
        }
        if (lane == 0) {
            s_nb_count[k] = count;
            s_nb_overflow[k] = overflow;
        }
    }
    __syncthreads();

    const double* abr = plan.slotbr + (size_t)aslot * SLOTBR_STRIDE;
    const double a1 = abr[20], a2 = abr[21], theta = abr[22];
    int rot = 0;     // warp-iterations dealt so far (mod TASK_WARPS): where the next image starts

    for (int k = 0; k < nimg; ++k) {
        const int n = th.n0 + k;
        const ImageDev img = field.images[n];
        const PatchDev pa = field.patches[plan.src_row[aslot] + (size_t)n * field.S_tot];
        const int npix = pa.H2 * pa.W2;
        const int nit = (npix + 31) >> 5;
        const double* comps_k = s_comps + k * MAX_COMPS * COMP_STRIDE;
        const double* arec = plan.slotimg + ((size_t)aslot * plan.N + n) * SLOTIMG_STRIDE;
        const double am1 = arec[MAX_COMPS * COMP_STRIDE], am2 = arec[MAX_COMPS * COMP_STRIDE + 1];
        const int b = img.band - 1;
        const double cb[4] = {a1 * abr[b], a2 * abr[5 + b], a1 * abr[10 + b], a2 * abr[15 + b]};

        double cnt_other_active = 0.0;
        auto neighbour = [&](int s, int h, int w, double& Ebg, double& Vbg, double& cnt) {
            const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
            const int h2 = h - p.off_h, w2 = w - p.off_w;
            if (h2 < 1 || h2 > p.H2 || w2 < 1 || w2 >= p.W2) return;
            if (!p.bitmap[(h2 - 1) + (size_t)(w2 - 1) * p.H2]) return;
            bool other_active = false;
            if (multi)
                for (int j = sub0; j < sub1; ++j) other_active |= (plan.sub_slot[j] == s);
            if (other_active)
                cnt_other_active += 1.0;
            else
                cnt += 1.0;
            const double* rec = plan.slotimg + ((size_t)s * plan.N + n) * SLOTIMG_STRIDE;
            const double* br = plan.slotbr + (size_t)s * SLOTBR_STRIDE;
            const double m1 = __ldg(rec + MAX_COMPS * COMP_STRIDE), m2 = __ldg(rec + MAX_COMPS * COMP_STRIDE + 1);
            double f0, gd[2], hd[3];
            star_eval<0>(LdGlobal(), p.coefs, p.n1, p.n2, (double)h - m1 + 26.0, (double)w - m2 + 26.0, f0, gd, hd);
            const double f1 = gal_value<KT>(LdGlobal(), rec, p.K, s_exptab, __ldg(br + 22), (double)h, (double)w);
            const double na1 = __ldg(br + 20), na2 = __ldg(br + 21);
            const double Es = na1 * __ldg(br + b) * f0 + na2 * __ldg(br + 5 + b) * f1;
            const double E2s = na1 * __ldg(br + 10 + b) * f0 * f0 + na2 * __ldg(br + 15 + b) * f1 * f1;
            Ebg += Es;
            Vbg += E2s - Es * Es;
        };
        struct PixIn {
            unsigned char active;
            float x, sky, iota;
            double pixconst;
        };
        auto fetch = [&](int it) {
            PixIn in;
            in.active = 0;
            in.x = in.sky = in.iota = 0.f;
            in.pixconst = 0.0;
            const int pix = it * 32 + lane;
            if (it < nit && pix < npix) {
                const int h2 = pix % pa.H2, w2 = pix / pa.H2;
                const size_t ipix = (size_t)(pa.off_h + h2) + (size_t)(pa.off_w + w2) * img.H;
                in.active = pa.bitmap[pix];
                in.x = img.pixels[ipix];
                in.sky = img.sky[ipix];
                in.iota = img.iota[pa.off_h + h2];
                in.pixconst = img.pixconst[ipix];
            }
            return in;
        };
        // this warp's iterations of image n: it = first, first + 4, ...  with (it + rot) % 4 == warp
        const int first = (warp - rot + TASK_WARPS) % TASK_WARPS;
        PixIn cur = fetch(first);
        for (int it = first; it < nit; it += TASK_WARPS) {
            const PixIn nxt = fetch(it + TASK_WARPS);
            const int pix = it * 32 + lane;
            if (cur.active && !isnan(cur.x)) {
                const int h2 = pix % pa.H2, w2 = pix / pa.H2;
                const int h = pa.off_h + h2 + 1, w = pa.off_w + w2 + 1;
                PixelConsts pc;
                pc.x = (double)cur.x;
                pc.iota = (double)cur.iota;
                pc.pixconst = cur.pixconst;
                double Ebg = (double)cur.sky, Vbg = 0.0, cnt_inactive = 0.0;
                cnt_other_active = 0.0;
                bool first_visit = true;
                if (multi)
                    for (int j = sub0; j < sub; ++j) {
                        const PatchDev& pj = field.patches[plan.src_row[plan.sub_slot[j]] + (size_t)n * field.S_tot];
                        const int hj = h - pj.off_h, wj = w - pj.off_w;
                        if (hj >= 1 && hj <= pj.H2 && wj >= 1 && wj <= pj.W2 && pj.bitmap[(hj - 1) + (size_t)(wj - 1) * pj.H2])
                            first_visit = false;
                    }
                const int nnb = s_nb_count[k];
                for (int i = 0; i < nnb; ++i) neighbour(s_nb[k][i], h, w, Ebg, Vbg, cnt_inactive);
                if (s_nb_overflow[k]) {
                    const int last_listed = s_nb[k][TASK_NB - 1];
                    for (int s = last_listed + 1; s < slot1; ++s)
                        if (s != aslot) neighbour(s, h, w, Ebg, Vbg, cnt_inactive);
                }
                const bool covered = (w2 + 1) < pa.W2;
                double f0 = 0.0, g0[2] = {0.0, 0.0}, h0[3] = {0.0, 0.0, 0.0};
                GalRaw gal;
                gal.f = 0.0;
                if (covered) {
                    star_eval<MODE>(LdGlobal(), pa.coefs, pa.n1, pa.n2, (double)h - am1 + 26.0, (double)w - am2 + 26.0, f0,
                                    g0, h0);
                    gal_eval<MODE, KT>(LdShared(), comps_k, pa.K, c_proto_nu, s_exptab, theta, (double)h, (double)w, gal);
                }
                if (first_visit) {
                    acc[ACC_CNT_ACTIVE * PIX_THREADS + tid] += (covered ? 1.0 : 0.0) + cnt_other_active;
                    acc[ACC_CNT_INACTIVE * PIX_THREADS + tid] += cnt_inactive;
                }
                pixel_accumulate<MODE>(acc + tid, PIX_THREADS, pc, Ebg, Vbg, covered, first_visit, cb, f0, g0, h0, gal);
            }
            cur = nxt;
        }
        rot = (rot + nit) % TASK_WARPS;
        // leave image n: this warp's partial (fixed shuffle order), then clear the slots
        double* out = plan.partials + ((size_t)(th.tn0 + n) * TASK_WARPS + warp) * NACC;
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            double v = acc[a * PIX_THREADS + tid];
            acc[a * PIX_THREADS + tid] = 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) out[a] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sa > 1 only (unit tests of the reference, test/test_elbo.jl:64-130,223-301): the Hessian block that couples
// two active sources a, b of one task.  E_G and var_G are sums over sources, so the only coupling is through
// combine_sfs_hessian! (SensitiveFloats.jl:114-126):
//     H_ab += x * ( h12 (dV_a dE_b' + dE_a dV_b') + h22 dE_a dE_b' )
// accumulated here in (c, y)_a x (c, y)_b space (10 x 10) over the pixels both sources cover.
constexpr int PAIR_THREADS = 128;

template <int KT>
__global__ void __launch_bounds__(PAIR_THREADS) pair_kernel(PlanDev plan) {
    CEL_DYNAMIC_SMEM(smem);
    double* acc = smem;                                         // NPAIR_ACC x PAIR_THREADS
    double* s_comps_a = acc + NPAIR_ACC * PAIR_THREADS;
    double* s_comps_b = s_comps_a + MAX_COMPS * COMP_STRIDE;
    __shared__ double s_exptab[8];
    const int tid = threadIdx.x;
    const PairHdr ph = plan.pairmap[blockIdx.x];
    const int n = ph.n;
    const FieldDev field = plan.fields[ph.field];
    const ImageDev img = field.images[n];
    const PatchDev pa = field.patches[plan.src_row[ph.slot_a] + (size_t)n * field.S_tot];
    const PatchDev pb = field.patches[plan.src_row[ph.slot_b] + (size_t)n * field.S_tot];
    for (int a = 0; a < NPAIR_ACC; ++a) acc[a * PAIR_THREADS + tid] = 0.0;
    const double* reca = plan.slotimg + ((size_t)ph.slot_a * plan.N + n) * SLOTIMG_STRIDE;
    const double* recb = plan.slotimg + ((size_t)ph.slot_b * plan.N + n) * SLOTIMG_STRIDE;
    for (int i = tid; i < NPROTO * pa.K * COMP_STRIDE; i += PAIR_THREADS) s_comps_a[i] = reca[i];
    for (int i = tid; i < NPROTO * pb.K * COMP_STRIDE; i += PAIR_THREADS) s_comps_b[i] = recb[i];
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    __syncthreads();
    const int b = img.band - 1;
    const double* bra = plan.slotbr + (size_t)ph.slot_a * SLOTBR_STRIDE;
    const double* brb = plan.slotbr + (size_t)ph.slot_b * SLOTBR_STRIDE;
    const double cba[4] = {bra[20] * bra[b], bra[21] * bra[5 + b], bra[20] * bra[10 + b], bra[21] * bra[15 + b]};
    const double cbb[4] = {brb[20] * brb[b], brb[21] * brb[5 + b], brb[20] * brb[10 + b], brb[21] * brb[15 + b]};
    // pixels covered by both: rows off+1..off+H2, columns off+1..off+W2-1 (strict, elbo_objective.jl:349)
    const int h_lo = max(pa.off_h, pb.off_h) + 1, h_hi = min(pa.off_h + pa.H2, pb.off_h + pb.H2);
    const int w_lo = max(pa.off_w, pb.off_w) + 1, w_hi = min(pa.off_w + pa.W2 - 1, pb.off_w + pb.W2 - 1);
    const int nh = h_hi - h_lo + 1, nw = w_hi - w_lo + 1;
    const int npix = (nh > 0 && nw > 0) ? nh * nw : 0;
    for (int pix = tid; pix < npix; pix += PAIR_THREADS) {
        const int h = h_lo + pix % nh, w = w_lo + pix / nh;
        if (!pa.bitmap[(h - pa.off_h - 1) + (size_t)(w - pa.off_w - 1) * pa.H2]) continue;
        if (!pb.bitmap[(h - pb.off_h - 1) + (size_t)(w - pb.off_w - 1) * pb.H2]) continue;
        const size_t ipix = (size_t)(h - 1) + (size_t)(w - 1) * img.H;
        const float xf = img.pixels[ipix];
        if (isnan(xf)) continue;
        double E = (double)img.sky[ipix], V = 0.0;
        double ea[10], va[10], eb[10], vb[10];
        for (int which = 0; which < 2; ++which) {
            const PatchDev& pp = which == 0 ? pa : pb;
            const double* rec = which == 0 ? reca : recb;
            const double* br = which == 0 ? bra : brb;
            const double* cb = which == 0 ? cba : cbb;
            double f0, g0[2], h0[3];
            GalRaw gal;
            star_eval<1>(LdGlobal(), pp.coefs, pp.n1, pp.n2, (double)h - rec[MAX_COMPS * COMP_STRIDE] + 26.0,
                         (double)w - rec[MAX_COMPS * COMP_STRIDE + 1] + 26.0, f0, g0, h0);
            gal_eval<1, KT>(LdShared(), which == 0 ? s_comps_a : s_comps_b, pp.K, c_proto_nu, s_exptab, br[22], (double)h,
                            (double)w, gal);
            const double A1 = cb[0], A2 = cb[1], B1 = cb[2], B2 = cb[3], f1 = gal.f;
            const double m = A1 * f0 + A2 * f1;
            E += m;
            V += B1 * f0 * f0 + B2 * f1 * f1 - m * m;
            double* e = which == 0 ? ea : eb;
            double* v = which == 0 ? va : vb;
            e[0] = f0;
            e[1] = f1;
            e[2] = e[3] = 0.0;
            v[0] = -2.0 * m * f0;
            v[1] = -2.0 * m * f1;
            v[2] = f0 * f0;
            v[3] = f1 * f1;
            const double vf0 = 2.0 * (B1 * f0 - m * A1), vf1 = 2.0 * (B2 * f1 - m * A2);
            for (int k = 0; k < 6; ++k) {
                const double gk = k < 2 ? g0[k] : 0.0;
                e[4 + k] = A1 * gk + A2 * gal.r[k];
                v[4 + k] = vf0 * gk + vf1 * gal.r[k];
            }
        }
        // every other source of the task: values only
        for (int s = ph.slot0; s < ph.slot1; ++s) {
            if (s == ph.slot_a || s == ph.slot_b) continue;
            const PatchDev& p = field.patches[plan.src_row[s] + (size_t)n * field.S_tot];
            const int h2 = h - p.off_h, w2 = w - p.off_w;
            if (h2 < 1 || h2 > p.H2 || w2 < 1 || w2 >= p.W2) continue;
            if (!p.bitmap[(h2 - 1) + (size_t)(w2 - 1) * p.H2]) continue;
            const double* rec = plan.slotimg + ((size_t)s * plan.N + n) * SLOTIMG_STRIDE;
            const double* br = plan.slotbr + (size_t)s * SLOTBR_STRIDE;
            double f0, gd[2], hd[3];
            star_eval<0>(LdGlobal(), p.coefs, p.n1, p.n2, (double)h - rec[MAX_COMPS * COMP_STRIDE] + 26.0,
                         (double)w - rec[MAX_COMPS * COMP_STRIDE + 1] + 26.0, f0, gd, hd);
            const double f1 = gal_value<KT>(LdGlobal(), rec, p.K, s_exptab, br[22], (double)h, (double)w);
            const double Es = br[20] * br[b] * f0 + br[21] * br[5 + b] * f1;
            E += Es;
            V += br[20] * br[10 + b] * f0 * f0 + br[21] * br[15 + b] * f1 * f1 - Es * Es;
        }
        const double x = (double)xf;
        const double iE = 1.0 / E, iE2 = iE * iE;
        const double LEE = -x * (iE2 + 3.0 * V * iE2 * iE2);
        const double LEV = x * iE2 * iE;
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j)
                acc[(i * 10 + j) * PAIR_THREADS + tid] += LEE * ea[i] * eb[j] + LEV * (ea[i] * vb[j] + va[i] * eb[j]);
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    double* out = plan.pair_partials + (size_t)blockIdx.x * NPAIR_ACC;
    for (int a = warp; a < NPAIR_ACC; a += PAIR_THREADS / 32) {
        double t = 0.0;
        for (int k = 0; k < PAIR_THREADS / 32; ++k) t += acc[a * PAIR_THREADS + lane + 32 * k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) out[a] = t;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int EPI_THREADS = 128;
constexpr int NY = 10;   // intermediate variables: c (4) then y (6)

// d(c, y)/d(28 live parameters) for one (source, image): rows 0..3 c = (a1 E_l1, a2 E_l2, a1 E_ll1, a2 E_ll2),
// rows 4..9 y = (x1 x2 S11 S12 S22 theta)
__device__ inline void build_Jy(double (*Jy)[NLIVE], const PatchDev& p, int b, const double* br, const double J0[3][3]) {
    for (int r = 0; r < NY; ++r)
        for (int q = 0; q < NLIVE; ++q) Jy[r][q] = 0.0;
    Jy[4][0] = -p.J[0];      // dx_a/dpos_b = -J[a][b]
    Jy[4][1] = -p.J[2];
    Jy[5][0] = -p.J[1];
    Jy[5][1] = -p.J[3];
    for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 3; ++j) Jy[6 + k][3 + j] = J0[k][j];
    Jy[9][2] = 1.0;
    double ka[10], la[10];
    band_coefs(b, ka, la);
    for (int i = 0; i < 2; ++i) {
        const double ai = br[20 + i], El = br[i * 5 + b], Ell = br[10 + i * 5 + b];
        Jy[i][26 + i] = El;
        Jy[2 + i][26 + i] = Ell;
        for (int k = 0; k < 10; ++k) {
            Jy[i][bright_id(i, k)] = ai * El * ka[k];
            Jy[2 + i][bright_id(i, k)] = ai * Ell * la[k];
        }
    }
}

// inverse of bright_id: canonical (0-based) parameter q -> (type i, brightness parameter k); false if q is none
__device__ inline bool bright_of(int q, int& i, int& k) {
    if (q < 6 || q >= 26) return false;
    if (q < 8) {
        i = q - 6;
        k = 0;
    } else if (q < 10) {
        i = q - 8;
        k = 1;
    } else if (q < 18) {
        i = (q - 10) >> 2;
        k = 2 + ((q - 10) & 3);
    } else {
        i = (q - 18) >> 2;
        k = 6 + ((q - 18) & 3);
    }
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(EPI_THREADS) epilogue_kernel(PlanDev plan,
                                                               const double* __restrict__ vp, double* __restrict__ out_v,
                                                               double* __restrict__ out_d, double* __restrict__ out_h,
                                                               long long* __restrict__ out_counters,
                                                               int* __restrict__ out_flags, int hess_packed = 0) {
    constexpr int NACC = NAcc<MODE>::value;
    __shared__ double ysum[NACC_MODE2 > NPAIR_ACC ? NACC_MODE2 : NPAIR_ACC];
    __shared__ double Jy[NY][NLIVE];
    __shared__ double Jy2[NY][NLIVE];
    __shared__ double Hyy[NY][NY];
    __shared__ double Wm[NY][NLIVE];
    __shared__ double Hacc[NLIVE][NLIVE];
    __shared__ double gacc[NLIVE];
    __shared__ double s_val, s_cnt[2];
    __shared__ double J0[3][3], T0[3][3][3], J0b[3][3];
    __shared__ double s_kap[10], s_lam[10];
    __shared__ int s_bad;

    const int tid = threadIdx.x;
    const int t = blockIdx.x;
    if (plan.task_mask && !plan.task_mask[t]) return;
    const int sub0 = plan.sub_ptr[t], sub1 = plan.sub_ptr[t + 1];
    const int Sa = sub1 - sub0;
    const int P = NPARAM * Sa;
    const FieldDev field = plan.fields[plan.task_field[t]];
    // hess_packed (Sa = 1 plans): 406 doubles per task, the upper triangle of the 28 x 28 live block, row-major
    double* Hout = MODE >= 2 ? out_h + (hess_packed ? (long long)HESS_PACKED_LEN * t : plan.h_ptr[t]) : nullptr;

    if (tid == 0) {
        s_val = 0.0;
        s_cnt[0] = s_cnt[1] = 0.0;
        s_bad = 0;
    }
    if (MODE >= 2 && Sa > 1 && !hess_packed)
        for (int i = tid; i < P * P; i += EPI_THREADS) Hout[i] = 0.0;   // cross blocks are filled below
    __syncthreads();
    int bad = 0;

    for (int sub = sub0; sub < sub1; ++sub) {
        const int aslot = plan.sub_slot[sub];
        const int ka = sub - sub0;
        const double* vs = vp + (size_t)NPARAM * aslot;
        const double* br = plan.slotbr + (size_t)aslot * SLOTBR_STRIDE;
        for (int i = tid; i < NLIVE * NLIVE; i += EPI_THREADS) (&Hacc[0][0])[i] = 0.0;
        if (tid < NLIVE) gacc[tid] = 0.0;
        if (tid == 0 && MODE >= 1) sigma_derivs(vs[3], vs[4], vs[5], J0, T0);
        __syncthreads();
        // this thread's entries of the 28 x 28 block and what extra curvature term each takes: the same for every image
        constexpr int EPI_ENT = (NLIVE * NLIVE + EPI_THREADS - 1) / EPI_THREADS;
        int e_pp[EPI_ENT], e_q[EPI_ENT], e_kind[EPI_ENT], e_i[EPI_ENT], e_k1[EPI_ENT], e_k2[EPI_ENT];
        if (MODE >= 2) {
#pragma unroll
            for (int j = 0; j < EPI_ENT; ++j) {
                const int i = min(tid + j * EPI_THREADS, NLIVE * NLIVE - 1);
                const int pp = i / NLIVE, q = i % NLIVE;
                int ip = 0, kp = 0, iq = 0, kq = 0;
                const bool bp = bright_of(pp, ip, kp), bq = bright_of(q, iq, kq);
                e_pp[j] = pp;
                e_q[j] = q;
                e_kind[j] = 0;
                e_i[j] = e_k1[j] = e_k2[j] = 0;
                if (pp >= 3 && pp < 6 && q >= 3 && q < 6) {
                    e_kind[j] = 1;
                } else if (bp && bq && ip == iq) {
                    e_kind[j] = 2;
                    e_i[j] = ip;
                    e_k1[j] = kp;
                    e_k2[j] = kq;
                } else if ((bp && q == 26 + ip) || (bq && pp == 26 + iq)) {
                    e_kind[j] = 3;
                    e_i[j] = bp ? ip : iq;
                    e_k1[j] = bp ? kp : kq;
                }
            }
        }

        for (int n = 0; n < plan.N; ++n) {
            const int tn = sub * plan.N + n;
            const int c0 = plan.chunk_ptr[tn], c1 = plan.chunk_ptr[tn + 1];
            const PatchDev& p = field.patches[plan.src_row[aslot] + (size_t)n * field.S_tot];
            const int b = field.images[n].band - 1;
            // phase 1: fixed-order sum of the chunk partials; clear Jy; band coefficients
            for (int a = tid; a < NACC; a += EPI_THREADS) {
                double s = 0.0;
                for (int c = c0; c < c1; ++c) s += plan.partials[(size_t)c * NACC + a];
                ysum[a] = s;
            }
            if (MODE >= 1) {
                for (int i = tid; i < NY * NLIVE; i += EPI_THREADS) (&Jy[0][0])[i] = 0.0;
                if (tid == EPI_THREADS - 1) band_coefs(b, s_kap, s_lam);
            }
            __syncthreads();
            if (tid == 0) {
                s_val += ysum[ACC_VAL];
                s_cnt[0] += ysum[ACC_CNT_ACTIVE];
                s_cnt[1] += ysum[ACC_CNT_INACTIVE];
            }
            if (MODE >= 1) {
                // phase 2: the non-zero entries of Jy (same values as build_Jy), one thread each; Hyy from ysum
                if (tid < 4) {
                    Jy[4 + (tid >> 1)][tid & 1] = -p.J[(tid & 1) * 2 + (tid >> 1)];     // dx_a/dpos_b = -J[a][b]
                } else if (tid < 13) {
                    const int k = (tid - 4) / 3, j = (tid - 4) % 3;
                    Jy[6 + k][3 + j] = J0[k][j];
                } else if (tid == 13) {
                    Jy[9][2] = 1.0;
                } else if (tid >= 32 && tid < 52) {
                    const int i = (tid - 32) / 10, k = (tid - 32) % 10;
                    const double ai = br[20 + i], El = br[i * 5 + b], Ell = br[10 + i * 5 + b];
                    Jy[i][bright_id(i, k)] = ai * El * s_kap[k];
                    Jy[2 + i][bright_id(i, k)] = ai * Ell * s_lam[k];
                } else if (tid == 52 || tid == 53) {
                    const int i = tid - 52;
                    Jy[i][26 + i] = br[i * 5 + b];
                    Jy[2 + i][26 + i] = br[10 + i * 5 + b];
                }
                if (MODE >= 2 && tid >= 64) {
                    for (int e = tid - 64; e < NY * NY; e += EPI_THREADS - 64) {
                        const int r = e / NY, c = e % NY;
                        const int lo = r < c ? r : c, hi = r < c ? c : r;
                        Hyy[r][c] = hi < 4 ? ysum[ACC_CC + tri4(lo, hi)]
                                           : (lo < 4 ? ysum[ACC_CR + lo * 6 + (hi - 4)] : ysum[ACC_HH + tri6(lo - 4, hi - 4)]);
                    }
                }
                __syncthreads();
                // phase 3: gradient; W = Hyy Jy
                if (tid < NLIVE) {
                    double g = 0.0;
                    for (int c = 0; c < 4; ++c) g += Jy[c][tid] * ysum[ACC_C1 + c];
                    for (int k = 0; k < 6; ++k) g += Jy[4 + k][tid] * ysum[ACC_G + k];
                    gacc[tid] += g;
                }
                if (MODE >= 2) {
                    for (int i = tid; i < NY * NLIVE; i += EPI_THREADS) {
                        const int r = i / NLIVE, q = i % NLIVE;
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < NY; ++k) s += Hyy[r][k] * Jy[k][q];
                        Wm[r][q] = s;
                    }
                    __syncthreads();
                    // phase 4: Hacc += Jy' W + the curvature of Sigma(shape) + the curvature of c(a, beta), per entry
#pragma unroll
                    for (int j = 0; j < EPI_ENT; ++j) {
                        const int i = tid + j * EPI_THREADS;
                        if (i >= NLIVE * NLIVE) break;
                        const int pp = e_pp[j], q = e_q[j];
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < NY; ++r) s += Jy[r][pp] * Wm[r][q];
                        const int kind = e_kind[j];
                        if (kind == 1) {
                            // sum_k dL/dS_k * T0[k]   (transform_bvn_derivs_hessian!:481-488)
                            for (int k = 0; k < 3; ++k) s += ysum[ACC_G + 2 + k] * T0[k][pp - 3][q - 3];
                        } else if (kind == 2) {
                            // E * kappa kappa'
                            const int ip = e_i[j], kp = e_k1[j], kq = e_k2[j];
                            const double ai = br[20 + ip], El = br[ip * 5 + b], Ell = br[10 + ip * 5 + b];
                            s += ai * (ysum[ACC_C1 + ip] * El * s_kap[kp] * s_kap[kq] +
                                       ysum[ACC_C1 + 2 + ip] * Ell * s_lam[kp] * s_lam[kq]);
                        } else if (kind == 3) {
                            // the (a, beta) cross terms
                            const int i2 = e_i[j], k2 = e_k1[j];
                            s += ysum[ACC_C1 + i2] * br[i2 * 5 + b] * s_kap[k2] + ysum[ACC_C1 + 2 + i2] * br[10 + i2 * 5 + b] * s_lam[k2];
                        }
                        Hacc[pp][q] += s;
                    }
                }
            }
            __syncthreads();
        }

        // this source's blocks, SensitiveFloat layout (p fastest); rows/cols 29..44 (ids.k) stay zero
        if (MODE >= 1) {
            for (int i = tid; i < NPARAM; i += EPI_THREADS) {
                const double g = i < NLIVE ? gacc[i] : 0.0;
                out_d[(size_t)NPARAM * sub + i] = g;
                bad |= !isfinite(g);
            }
        }
        if (MODE >= 2 && hess_packed) {
            for (int i = tid; i < NLIVE * NLIVE; i += EPI_THREADS) {
                const int r = i / NLIVE, c = i % NLIVE;
                if (c < r) continue;
                const double v = 0.5 * (Hacc[r][c] + Hacc[c][r]);
                Hout[hess_packed_index(r, c)] = v;
                bad |= !isfinite(v);
            }
        } else if (MODE >= 2) {
            for (int i = tid; i < NPARAM * NPARAM; i += EPI_THREADS) {
                const int r = i % NPARAM, c = i / NPARAM;
                double v = 0.0;
                if (r < NLIVE && c < NLIVE) v = 0.5 * (Hacc[r][c] + Hacc[c][r]);   // exactly symmetric output
                Hout[(size_t)(NPARAM * ka + r) + (size_t)(NPARAM * ka + c) * P] = v;
                bad |= !isfinite(v);
            }
        }
        __syncthreads();
    }

    // cross-source blocks (Sa > 1): H_ab = sum_n Jy_a' M_n Jy_b
    if (MODE >= 2 && Sa > 1) {
        int pair = plan.pair_ptr[t];
        for (int ka = 0; ka < Sa; ++ka)
            for (int kb = ka + 1; kb < Sa; ++kb, ++pair) {
                const int sa = plan.sub_slot[sub0 + ka], sb = plan.sub_slot[sub0 + kb];
                const double* vsa = vp + (size_t)NPARAM * sa;
                const double* vsb = vp + (size_t)NPARAM * sb;
                for (int i = tid; i < NLIVE * NLIVE; i += EPI_THREADS) (&Hacc[0][0])[i] = 0.0;
                if (tid == 0) {
                    sigma_derivs(vsa[3], vsa[4], vsa[5], J0, T0);
                    double Ttmp[3][3][3];
                    sigma_derivs(vsb[3], vsb[4], vsb[5], J0b, Ttmp);
                }
                __syncthreads();
                for (int n = 0; n < plan.N; ++n) {
                    const double* M = plan.pair_partials + ((size_t)pair * plan.N + n) * NPAIR_ACC;
                    if (tid < NPAIR_ACC) ysum[tid] = M[tid];
                    const int b = field.images[n].band - 1;
                    if (tid == 0)
                        build_Jy(Jy, field.patches[plan.src_row[sa] + (size_t)n * field.S_tot], b,
                                 plan.slotbr + (size_t)sa * SLOTBR_STRIDE, J0);
                    if (tid == 32)
                        build_Jy(Jy2, field.patches[plan.src_row[sb] + (size_t)n * field.S_tot], b,
                                 plan.slotbr + (size_t)sb * SLOTBR_STRIDE, J0b);
                    __syncthreads();
                    for (int i = tid; i < NY * NLIVE; i += EPI_THREADS) {
                        const int r = i / NLIVE, q = i % NLIVE;
                        double s = 0.0;
                        for (int k = 0; k < NY; ++k) s += ysum[r * 10 + k] * Jy2[k][q];
                        Wm[r][q] = s;
                    }
                    __syncthreads();
                    for (int i = tid; i < NLIVE * NLIVE; i += EPI_THREADS) {
                        const int pp = i / NLIVE, q = i % NLIVE;
                        double s = 0.0;
                        for (int r = 0; r < NY; ++r) s += Jy[r][pp] * Wm[r][q];
                        Hacc[pp][q] += s;
                    }
                    __syncthreads();
                }
                for (int i = tid; i < NLIVE * NLIVE; i += EPI_THREADS) {
                    const int r = i / NLIVE, c = i % NLIVE;
                    const double v = Hacc[r][c];
                    Hout[(size_t)(NPARAM * ka + r) + (size_t)(NPARAM * kb + c) * P] = v;
                    Hout[(size_t)(NPARAM * kb + c) + (size_t)(NPARAM * ka + r) * P] = v;
                    bad |= !isfinite(v);
                }
                __syncthreads();
            }
    }

    if (tid == 0) {
        out_v[t] = s_val;
        out_counters[2 * t] = (long long)(s_cnt[0] + 0.5);
        out_counters[2 * t + 1] = (long long)(s_cnt[1] + 0.5);
        bad |= !isfinite(s_val);
    }
    if (bad) atomicOr(&s_bad, 1);
    __syncthreads();
    if (tid == 0) out_flags[t] = s_bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// epilogue_hess_kernel: the Hessian-mode epilogue of plans whose tasks all have Sa = 1 (the production shape), ONE WARP
// per task.  Same chain rule as epilogue_kernel<2>, organised around the sparsity of Jy = d(c, y) / d(28 live parameters):
// a parameter reaches at most three (c, y) rows -- position -> x (2 rows); gal_frac_dev -> theta (1); axis ratio, angle,
// radius -> Sigma (3); the brightness parameters and is_star of type i -> (A_i, B_i) (2) -- so an entry of the 28 x 28
// block costs <= 3 x 3 terms instead of two dense 10-term products, the 406 upper-triangle entries live in registers
// (13 per lane) across the images, and nothing synchronises but the warp.  epilogue_kernel<2> (a block per task, dense
// products through shared memory, 20 block barriers per task) took 0.6 ms per 10 000 tasks -- 12 % of the Hessian step.
constexpr int EPH_WARPS = 4;
constexpr int EPH_ENT = (HESS_PACKED_LEN + 31) / 32;   // upper-triangle entries per lane

struct EphWarp {             // per-warp shared memory
    double ysum[NACC_MODE2];
    double Hyy[NY][NY];
    double jv[NLIVE][3];     // the <= 3 non-zero entries of each column of Jy (zero-padded) ...
    int jr[NLIVE][3];        // ... and their rows
    double J0[3][3], T0[3][3][3];
    double kap[10], lam[10];
    double El[2][5], Ell[2][5];
};

__global__ void __launch_bounds__(32 * EPH_WARPS) epilogue_hess_kernel(PlanDev plan, const double* __restrict__ vp,
                                                                       double* __restrict__ out_v, double* __restrict__ out_d,
                                                                       double* __restrict__ out_h,
                                                                       long long* __restrict__ out_counters,
                                                                       int* __restrict__ out_flags, int hess_packed) {
    __shared__ EphWarp sw[EPH_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * EPH_WARPS + warp;
    if (t >= plan.n_tasks) return;
    if (plan.task_mask && !plan.task_mask[t]) return;
    EphWarp& W = sw[warp];
    const int sub = plan.sub_ptr[t];                      // Sa == 1
    const int aslot = plan.sub_slot[sub];
    const double* vs = vp + (size_t)NPARAM * aslot;
    const double* br = plan.slotbr + (size_t)aslot * SLOTBR_STRIDE;
    const FieldDev field = plan.fields[plan.task_field[t]];
    if (lane == 0) sigma_derivs(vs[3], vs[4], vs[5], W.J0, W.T0);
    if (lane < 10) {
        W.El[lane / 5][lane % 5] = br[lane];
        W.Ell[lane / 5][lane % 5] = br[10 + lane];
    }
    // this lane's upper-triangle entries (pp <= q) and the extra curvature each takes -- the same for every image
    int e_code[EPH_ENT];
#pragma unroll
    for (int j = 0; j < EPH_ENT; ++j) {
        const int e = min(lane + 32 * j, HESS_PACKED_LEN - 1);
        int pp = 0, rem = e;                              // row-major upper triangle: row pp holds NLIVE - pp entries
        while (rem >= NLIVE - pp) {
            rem -= NLIVE - pp;
            ++pp;
        }
        const int q = pp + rem;
        int ip = 0, kp = 0, iq = 0, kq = 0, kind = 0, a = 0, k1 = 0, k2 = 0;
        const bool bp = bright_of(pp, ip, kp), bq = bright_of(q, iq, kq);
        if (pp >= 3 && pp < 6 && q >= 3 && q < 6) {
            kind = 1;
        } else if (bp && bq && ip == iq) {
            kind = 2;
            a = ip;
            k1 = kp;
            k2 = kq;
        } else if ((bp && q == 26 + ip) || (bq && pp == 26 + iq)) {
            kind = 3;
            a = bp ? ip : iq;
            k1 = bp ? kp : kq;
        }
        e_code[j] = pp | (q << 5) | (kind << 10) | (a << 12) | (k1 << 13) | (k2 << 17);
    }
    double hacc[EPH_ENT];
#pragma unroll
    for (int j = 0; j < EPH_ENT; ++j) hacc[j] = 0.0;
    double g = 0.0, val = 0.0, cnt0 = 0.0, cnt1 = 0.0;
    const double a0 = br[20], a1 = br[21];
    __syncwarp();

    for (int n = 0; n < plan.N; ++n) {
        const int tn = sub * plan.N + n;
        const int c0 = plan.chunk_ptr[tn], c1 = plan.chunk_ptr[tn + 1];
        const PatchDev& p = field.patches[plan.src_row[aslot] + (size_t)n * field.S_tot];
        const int b = field.images[n].band - 1;
        // 1. fixed-order sum of the unit partials; band coefficients
        for (int a = lane; a < NACC_MODE2; a += 32) {
            double s = 0.0;
            for (int c = c0; c < c1; ++c) s += plan.partials[(size_t)c * NACC_MODE2 + a];
            W.ysum[a] = s;
        }
        if (lane == 31) band_coefs(b, W.kap, W.lam);
        __syncwarp();
        // 2. Hyy from the packed sums; the non-zero entries of every column of Jy
        for (int e = lane; e < NY * NY; e += 32) {
            const int r = e / NY, c = e % NY;
            const int lo = r < c ? r : c, hi = r < c ? c : r;
            W.Hyy[r][c] = hi < 4 ? W.ysum[ACC_CC + tri4(lo, hi)]
                                 : (lo < 4 ? W.ysum[ACC_CR + lo * 6 + (hi - 4)] : W.ysum[ACC_HH + tri6(lo - 4, hi - 4)]);
        }
        if (lane < NLIVE) {
            const int q = lane;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            int r0 = 0, r1 = 0, r2 = 0;
            int i, k;
            if (q < 2) {                                  // dx_a / dpos_q = -J[a][q]
                r0 = 4;
                r1 = 5;
                v0 = -p.J[2 * q];
                v1 = -p.J[2 * q + 1];
            } else if (q == 2) {
                r0 = 9;
                v0 = 1.0;
            } else if (q < 6) {
                r0 = 6;
                r1 = 7;
                r2 = 8;
                v0 = W.J0[0][q - 3];
                v1 = W.J0[1][q - 3];
                v2 = W.J0[2][q - 3];
            } else if (bright_of(q, i, k)) {
                const double ai = i == 0 ? a0 : a1;
                r0 = i;
                r1 = 2 + i;
                v0 = ai * W.El[i][b] * W.kap[k];
                v1 = ai * W.Ell[i][b] * W.lam[k];
            } else {                                      // is_star
                i = q - 26;
                r0 = i;
                r1 = 2 + i;
                v0 = W.El[i][b];
                v1 = W.Ell[i][b];
            }
            W.jr[q][0] = r0;
            W.jr[q][1] = r1;
            W.jr[q][2] = r2;
            W.jv[q][0] = v0;
            W.jv[q][1] = v1;
            W.jv[q][2] = v2;
        }
        __syncwarp();
        if (lane == 0) {
            val += W.ysum[ACC_VAL];
            cnt0 += W.ysum[ACC_CNT_ACTIVE];
            cnt1 += W.ysum[ACC_CNT_INACTIVE];
        }
        // 3. gradient: row r of the first-order sums is C1 (r < 4) or G (r >= 4)
        if (lane < NLIVE) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int r = W.jr[lane][a];
                g = fma(W.jv[lane][a], r < 4 ? W.ysum[ACC_C1 + r] : W.ysum[ACC_G + r - 4], g);
            }
        }
        // 4. H[pp][q] += Jy[:, pp]' Hyy Jy[:, q] + the curvature of Sigma(shape) and of c(a, beta)
#pragma unroll
        for (int j = 0; j < EPH_ENT; ++j) {
            const int code = e_code[j];
            const int pp = code & 31, q = (code >> 5) & 31, kind = (code >> 10) & 3;
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int ra = W.jr[pp][a];
                double tq = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) tq = fma(W.Hyy[ra][W.jr[q][c]], W.jv[q][c], tq);
                s = fma(W.jv[pp][a], tq, s);
            }
            if (kind == 1) {
                // sum_k dL/dS_k * T0[k]   (transform_bvn_derivs_hessian!:481-488)
                for (int k = 0; k < 3; ++k) s += W.ysum[ACC_G + 2 + k] * W.T0[k][pp - 3][q - 3];
            } else if (kind == 2) {
                const int ip = (code >> 12) & 1, kp = (code >> 13) & 15, kq = (code >> 17) & 15;
                const double ai = ip == 0 ? a0 : a1;      // E * kappa kappa'
                s += ai * (W.ysum[ACC_C1 + ip] * W.El[ip][b] * W.kap[kp] * W.kap[kq] +
                           W.ysum[ACC_C1 + 2 + ip] * W.Ell[ip][b] * W.lam[kp] * W.lam[kq]);
            } else if (kind == 3) {
                const int i2 = (code >> 12) & 1, k2 = (code >> 13) & 15;      // the (a, beta) cross terms
                s += W.ysum[ACC_C1 + i2] * W.El[i2][b] * W.kap[k2] + W.ysum[ACC_C1 + 2 + i2] * W.Ell[i2][b] * W.lam[k2];
            }
            hacc[j] += s;
        }
        __syncwarp();                                     // the image's tables are consumed
    }

    int bad = 0;
    if (lane < NLIVE) bad |= !isfinite(g);
    for (int i = lane; i < NPARAM; i += 32) out_d[(size_t)NPARAM * sub + i] = 0.0;
    __syncwarp();
    if (lane < NLIVE) out_d[(size_t)NPARAM * sub + lane] = g;
    if (hess_packed) {
        double* Hout = out_h + (size_t)HESS_PACKED_LEN * t;
#pragma unroll
        for (int j = 0; j < EPH_ENT; ++j) {
            const int e = lane + 32 * j;
            if (e < HESS_PACKED_LEN) {
                Hout[e] = hacc[j];
                bad |= !isfinite(hacc[j]);
            }
        }
    } else {
        double* Hout = out_h + plan.h_ptr[t];
        for (int i = lane; i < NPARAM * NPARAM; i += 32) Hout[i] = 0.0;      // rows / cols 29..44 (ids.k) stay zero
        __syncwarp();
#pragma unroll
        for (int j = 0; j < EPH_ENT; ++j) {
            const int e = lane + 32 * j;
            if (e < HESS_PACKED_LEN) {
                const int pp = e_code[j] & 31, q = (e_code[j] >> 5) & 31;
                Hout[pp + (size_t)q * NPARAM] = hacc[j];                     // exactly symmetric by construction
                Hout[q + (size_t)pp * NPARAM] = hacc[j];
                bad |= !isfinite(hacc[j]);
            }
        }
    }
    if (lane == 0) {
        out_v[t] = val;
        out_counters[2 * t] = (long long)(cnt0 + 0.5);
        out_counters[2 * t + 1] = (long long)(cnt1 + 0.5);
        bad |= !isfinite(val);
    }
    const unsigned anybad = __ballot_sync(0xffffffffu, bad != 0);
    if (lane == 0) out_flags[t] = anybad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// render_kernel: the value-only full-image render -- what bin/write_celeste_expectation.jl:111-156
// (fill_celeste_expectation!) computes by calling add_pixel_term! in value mode on EVERY pixel of every image:
//     out[h, w] = E_G - sky = sum over the sources whose patch covers (h, w) of  a1 E_l1 f0 + a2 E_l2 f1
// with the same in-patch test as the ELBO (strict w2 < W2, bitmap; elbo_objective.jl:349) and the sources added in
// task order onto the sky (:374) exactly as there.  Pixel-centric: one block per non-empty 32 x 8 image tile
// (32 consecutive rows = one coalesced warp access of the column-major image), the tile's source list binned on
// the host from the patch boxes; the component records of one source at a time go through shared memory.
constexpr int RT_H = 32, RT_W = 8, RENDER_THREADS = RT_H * RT_W;

struct RenderTile {
    int n;            // image
    int h0, w0;       // 0-based origin of the tile
    int begin, end;   // range into tile_slots (ascending slot order)
};

// host side: bin the S slots' patches of image n into tiles (CSR).  geo(s) -> off_h, off_w, H2, W2.
template <typename Geo>
inline void build_render_tiles(int N, const int* imgH, const int* imgW, int S, Geo geo, std::vector<RenderTile>& tiles,
                               std::vector<int>& tile_slots, int full_box = 0) {
    tiles.clear();
    tile_slots.clear();
    for (int n = 0; n < N; ++n) {
        const int H = imgH[n], W = imgW[n];
        const int th_n = (H + RT_H - 1) / RT_H, tw_n = (W + RT_W - 1) / RT_W;
        std::vector<int> count((size_t)th_n * tw_n + 1, 0);
        auto range = [&](int s, int& t0, int& t1, int& u0, int& u1) {
            int oh, ow, H2, W2;
            geo(s, n, oh, ow, H2, W2);
            const int hlo = std::max(oh + 1, 1), hhi = std::min(oh + H2, H);
            const int wlo = std::max(ow + 1, 1), whi = std::min(ow + W2 - (full_box ? 0 : 1), W);     // strict w2 < W2
            if (hlo > hhi || wlo > whi) return false;
            t0 = (hlo - 1) / RT_H;
            t1 = (hhi - 1) / RT_H;
            u0 = (wlo - 1) / RT_W;
            u1 = (whi - 1) / RT_W;
            return true;
        };
        int t0, t1, u0, u1;
        for (int s = 0; s < S; ++s)
            if (range(s, t0, t1, u0, u1))
                for (int u = u0; u <= u1; ++u)
                    for (int t = t0; t <= t1; ++t) count[(size_t)t + (size_t)u * th_n + 1]++;
        for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
        const int base = (int)tile_slots.size();
        tile_slots.resize(base + count.back());
        std::vector<int> fill(count.begin(), count.end() - 1);
        for (int s = 0; s < S; ++s)
            if (range(s, t0, t1, u0, u1))
                for (int u = u0; u <= u1; ++u)
                    for (int t = t0; t <= t1; ++t) tile_slots[base + fill[(size_t)t + (size_t)u * th_n]++] = s;
        for (int u = 0; u < tw_n; ++u)
            for (int t = 0; t < th_n; ++t) {
                const size_t i = (size_t)t + (size_t)u * th_n;
                if (count[i + 1] > count[i]) tiles.push_back(RenderTile{n, t * RT_H, u * RT_W, base + count[i], base + count[i + 1]});
            }
    }
}

template <int KT>
__global__ void __launch_bounds__(RENDER_THREADS) render_kernel(PlanDev plan, const RenderTile* __restrict__ tiles,
                                                                const int* __restrict__ tile_slots,
                                                                double* const* __restrict__ out, int full_box = 0) {
    __shared__ double s_comps[MAX_COMPS * COMP_STRIDE];
    __shared__ double s_exptab[8];
    const int tid = threadIdx.x;
    const RenderTile t = tiles[blockIdx.x];
    const FieldDev field = plan.fields[0];
    const ImageDev img = field.images[t.n];
    const int h = t.h0 + (tid % RT_H) + 1, w = t.w0 + (tid / RT_H) + 1;      // 1-based image coordinates
    const bool inside = h <= img.H && w <= img.W;
    const size_t ipix = (size_t)(h - 1) + (size_t)(w - 1) * img.H;
    const double sky = inside ? (double)img.sky[ipix] : 0.0;
    double E = sky;                                                          // E_G.v += sky (:374); sources on top
    const int b = img.band - 1;
#ifdef CELESTE_HOST_EMULATION
    if (tid < 8) s_exptab[tid] = h_exptab[tid];
#else
    if (tid < 8) s_exptab[tid] = c_exptab[tid];
#endif
    for (int i = t.begin; i < t.end; ++i) {
        const int s = tile_slots[i];
        const PatchDev& p = field.patches[plan.src_row[s] + (size_t)t.n * field.S_tot];
        const double* rec = plan.slotimg + ((size_t)s * plan.N + t.n) * SLOTIMG_STRIDE;
        const int nrec = NPROTO * (KT > 0 ? KT : p.K) * COMP_STRIDE;
        __syncthreads();                                                     // the previous source's records are consumed
        for (int j = tid; j < nrec; j += RENDER_THREADS) s_comps[j] = rec[j];
        __syncthreads();
        const int h2 = h - p.off_h, w2 = w - p.off_w;
        // full_box: every column of the box (Synthetic.gen_image! renders a body on its whole box); else the ELBO's rule
        if (!inside || h2 < 1 || h2 > p.H2 || w2 < 1 || w2 > p.W2 - (full_box ? 0 : 1)) continue;
        if (!p.bitmap[(h2 - 1) + (size_t)(w2 - 1) * p.H2]) continue;
        const double* br = plan.slotbr + (size_t)s * SLOTBR_STRIDE;
        const double m1 = rec[MAX_COMPS * COMP_STRIDE], m2 = rec[MAX_COMPS * COMP_STRIDE + 1];
        double f0, gd[2], hd[3];
        star_eval<0>(LdGlobal(), p.coefs, p.n1, p.n2, (double)h - m1 + 26.0, (double)w - m2 + 26.0, f0, gd, hd);
        const double f1 = gal_value<KT>(LdShared(), s_comps, p.K, s_exptab, br[22], (double)h, (double)w);
        E += br[20] * br[b] * f0 + br[21] * br[5 + b] * f1;
    }
    if (inside) out[t.n][ipix] = E - sky;
}

// ------------------------------------------------------------------------------------------------
// Register-resident DFMA chains: the measured FP64 peak that bounds pixel_kernel.
__global__ void dfma_peak_kernel(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;   // never true; keeps the chain alive
}

}  // namespace celeste
#endif
