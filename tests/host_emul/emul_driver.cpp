// emul_driver.cpp -- TEST INFRASTRUCTURE.  Runs the product's kernel source
// (celeste.jl_b200/csrc/celeste_kernels.cuh) under the host emulation layer with the same
// launch sequence as celeste_abi.cu (prep -> setup -> pixel -> epilogue), so kernel logic can
// be compared with the oracle without a GPU.  Built by tests/host_emul/Makefile; never shipped.
#include "cuda_emul.h"

#include "../../include/celeste_cuda.h"
#include "../../celeste.jl_b200/csrc/celeste_kernels.cuh"
#include "../../celeste.jl_b200/csrc/march_kernels.cuh"
#include "../../celeste.jl_b200/csrc/unit_kernels.cuh"
#include "../../celeste.jl_b200/csrc/maximize_kernels.cuh"
#include "../../celeste.jl_b200/csrc/patch_kernels.cuh"

using namespace celeste;

namespace {
void galaxy_prototypes(double eta[NPROTO], double nu[NPROTO]) {
    const double dev_amp[8] = {4.26347652e-2, 2.40127183e-1, 6.85907632e-1, 1.51937350,
                               2.83627243,    4.46467501,    5.72440830,    5.60989349};
    const double dev_var[8] = {2.23759216e-4, 1.00220099e-3, 4.18731126e-3, 1.69432589e-2,
                               6.84850479e-2, 2.87207080e-1, 1.33320254,    8.40215071};
    const double exp_amp[6] = {2.34853813e-3, 3.07995260e-2, 2.23364214e-1, 1.17949102, 4.33873750, 5.99820770};
    const double exp_var[6] = {1.20078965e-3, 8.84526493e-3, 3.91463084e-2,
                               1.39976817e-1, 4.60962500e-1, 1.50159566};
    double sd = 0, se = 0;
    for (double a : dev_amp) sd += a;
    for (double a : exp_amp) se += a;
    for (int j = 0; j < 8; ++j) {
        eta[j] = dev_amp[j] / sd;
        nu[j] = dev_var[j] / (1.078031 * 1.078031);
    }
    for (int j = 0; j < 6; ++j) {
        eta[8 + j] = exp_amp[j] / se;
        nu[8 + j] = exp_var[j] / (0.928896 * 0.928896);
    }
}

long g_march_split = 6000;   // pixels above which a source gets one march block per image (emul_set_grad_kernel(2) lowers it)
int g_grad_kernel = 3;   // 3: unit kernels where the product uses them (Sa = 1, K = 2); 1 / 2: march_kernel; 0: always task_kernel

long g_unit_target = CELESTE_UNIT_ROWS;  // build_unit_list: rows per unit (0 = never cut); emul_set_unit_target overrides
int g_hess_kernel = 1;   // 1: unit_kernel<2> where the product would use it (Sa = 1, K = 2); 0: always pixel_kernel<2>

// same launch sequence as celeste_abi.cu's unit path: setup (brightness moments for the epilogue), unit_kernel, epilogue
template <int MODE>
void run_unit(PlanDev pd, const FieldDev& fd, const double* vp, double* v, double* d, double* h, long long* counters,
              int* flags) {
    std::vector<int> sub_task(pd.n_subs);
    for (int u = 0; u < pd.n_subs; ++u) sub_task[u] = u;
    std::vector<UnitHdr> units, ub;
    std::vector<int> cp;
    long long maxpix = 1;
    build_unit_list(pd.n_subs, pd.N, sub_task.data(), pd.sub_slot, pd.task_ptr, (const int*)nullptr,
                    [&](int slot, int n, int& oh, int& ow, int& H2, int& W2) {
                        const PatchDev& pa = fd.patches[(size_t)pd.src_row[slot] + (size_t)n * fd.S_tot];
                        oh = pa.off_h;
                        ow = pa.off_w;
                        H2 = pa.H2;
                        W2 = pa.W2;
                    },
                    g_unit_target, 0L, g_unit_target > 0 ? UNIT_BG_PIXELS_SMALL : UNIT_BG_PIXELS_BIG, units, ub, cp, maxpix);
    const int grid = 2;
    std::vector<double> part((size_t)cp.back() * NAcc<MODE>::value + 1, 1e300);   // poisoned
    std::vector<long long> bg_ptr((size_t)pd.n_subs * pd.N, -1), l5_ptr((size_t)pd.n_subs * pd.N, 0);
    long long bg_total = 0, l5_total = 0;
    for (int u = 0; u < pd.n_subs; ++u)
        for (int n = 0; n < pd.N; ++n) {
            const PatchDev& pa = fd.patches[(size_t)pd.src_row[pd.sub_slot[u]] + (size_t)n * fd.S_tot];
            const long long px = (long long)std::max(pa.H2, 0) * std::max(pa.W2, 0);
            bg_ptr[(size_t)u * pd.N + n] = bg_total;
            bg_total += 2 * px;
            l5_ptr[(size_t)u * pd.N + n] = l5_total;
            l5_total += px;
        }
    std::vector<double> bg((size_t)bg_total + 1, 1e300), l5((size_t)l5_total + 1, 1e300);   // poisoned
    pd.partials = part.data();
    pd.chunk_ptr = cp.data();
    pd.bg_ptr = bg_ptr.data();
    pd.bg = bg.data();
    std::vector<PixRec> pix((size_t)l5_total + 1);
    pd.l5_ptr = l5_ptr.data();
    pd.l5 = l5.data();
    pd.pix = pix.data();
    std::vector<double> bg_cnt(ub.size() + 1, 1e300);
    pd.bg_cnt = bg_cnt.data();
    if (!units.empty()) cuda_emul::launch(unit_pack_kernel, (int)units.size(), 32, 0, pd, (const UnitHdr*)units.data(), (int)units.size(), pix.data());
    int queue[4] = {0, 0, 0, 0};
    cuda_emul::launch(slotbr_kernel, 1, std::max(pd.n_slots, 4), 0, pd, vp, &queue[0]);
    if (!ub.empty())
        cuda_emul::launch(unit_bg_kernel<3>, grid, UNIT_THREADS, unit_bg_smem_bytes(), pd, (const UnitHdr*)ub.data(), (int)ub.size(),
                          &queue[0], vp);
    cuda_emul::launch(unit_walk_kernel<MODE>, grid, UNIT_THREADS, unit_smem_bytes<MODE>(), pd, (const UnitHdr*)units.data(),
                      (int)units.size(), &queue[1], vp);
    if (MODE == 2)
        cuda_emul::launch(unit_moment_kernel, grid, UNIT_THREADS, unit_moment_smem_bytes(), pd, (const UnitHdr*)units.data(),
                          (int)units.size(), &queue[2], vp);
    if (MODE == 2)
        cuda_emul::launch(epilogue_hess_kernel, (pd.n_tasks + EPH_WARPS - 1) / EPH_WARPS, 32 * EPH_WARPS, 0, pd, vp, v, d, h, counters,
                          flags, 0);
    else
        cuda_emul::launch(epilogue_kernel<MODE>, pd.n_tasks, EPI_THREADS, 0, pd, vp, v, d, h, counters, flags, 0);
}

template <int MODE>
void run_task(PlanDev pd, const FieldDev& fd, const std::vector<TaskHdr>& taskmap, const std::vector<int>& tcp,
              const double* vp, double* v, double* d, double* h, long long* counters, int* flags) {
    if constexpr (MODE <= 1) {
        bool all_k2 = true;
        for (int i = 0; i < fd.S_tot * pd.N; ++i) all_k2 = all_k2 && fd.patches[i].K == 2;
        if (g_grad_kernel == 3 && all_k2 && pd.n_subs == pd.n_tasks) {
            run_unit<MODE>(pd, fd, vp, v, d, h, counters, flags);
            return;
        }
        if (g_grad_kernel == 1 && all_k2 && pd.n_subs == pd.n_tasks) {
            // same launch sequence as celeste_abi.cu's march path
            std::vector<int> sub_task(pd.n_subs), part_ptr;
            for (int u = 0; u < pd.n_subs; ++u) sub_task[u] = u;
            std::vector<MarchHdr> mm;
            build_march_blocks(pd.n_subs, pd.N, sub_task.data(), pd.sub_slot, pd.task_ptr, (const int*)nullptr,
                               [&](int slot, int n, int& oh, int& ow, int& H2, int& W2) {
                                   const PatchDev& pa = fd.patches[(size_t)pd.src_row[slot] + (size_t)n * fd.S_tot];
                                   oh = pa.off_h;
                                   ow = pa.off_w;
                                   H2 = pa.H2;
                                   W2 = pa.W2;
                               },
                               g_march_split, mm, part_ptr);
            std::vector<double> mpart(mm.size() * NT_ACC + 1);
            pd.partials = mpart.data();
            std::vector<long long> bg_ptr((size_t)pd.n_subs * pd.N, -1);
            long long bg_total = 0;
            for (int u = 0; u < pd.n_subs; ++u) {
                const int t = u;
                if (pd.task_ptr[t + 1] - pd.task_ptr[t] < 2) continue;
                for (int n = 0; n < pd.N; ++n) {
                    const PatchDev& pa = fd.patches[(size_t)pd.src_row[pd.sub_slot[u]] + (size_t)n * fd.S_tot];
                    if (pa.H2 <= 0 || pa.W2 <= 0) continue;
                    bg_ptr[(size_t)u * pd.N + n] = bg_total;
                    bg_total += 2LL * pa.H2 * pa.W2;
                }
            }
            std::vector<double> bg((size_t)bg_total + 1, 1e300);     // poisoned: the kernel must zero what it reads
            pd.bg_ptr = bg_ptr.data();
            pd.bg = bg.data();
            if (!mm.empty())
                cuda_emul::launch(march_kernel<MODE>, (int)mm.size(), MARCH_THREADS, march_smem_bytes(), pd, (const MarchHdr*)mm.data(),
                                  vp);
            cuda_emul::launch(march_epilogue_kernel<MODE>, pd.n_tasks, MEPI_THREADS, 0, pd, vp, (const int*)part_ptr.data(), v, d,
                              counters, flags);
            return;
        }
        const size_t smem = ((size_t)NAcc<MODE>::value * PIX_THREADS + (size_t)TASK_NIMG * MAX_COMPS * COMP_STRIDE) * sizeof(double);
        cuda_emul::launch(setup_kernel, 2, 64, 0, pd, vp);
        pd.chunk_ptr = tcp.data();
        bool k2 = true;
        for (int i = 0; i < fd.S_tot * pd.N; ++i) k2 = k2 && fd.patches[i].K == 2;
        if (!taskmap.empty()) {
            if (k2)
                cuda_emul::launch(task_kernel<MODE, 2, true>, (int)taskmap.size(), PIX_THREADS, smem, pd, taskmap.data());
            else
                cuda_emul::launch(task_kernel<MODE, 0, true>, (int)taskmap.size(), PIX_THREADS, smem, pd, taskmap.data());
        }
        cuda_emul::launch(epilogue_kernel<MODE>, pd.n_tasks, EPI_THREADS, 0, pd, vp, v, d, h, counters, flags, 0);
    }
}

template <int MODE>
void run(const PlanDev& pd, const FieldDev& fd, int n_blocks, int chunk_pixels, const double* vp, double* v, double* d,
         double* h, long long* counters, int* flags) {
    const size_t smem = ((size_t)NAcc<MODE>::value * PIX_THREADS + (size_t)MAX_COMPS * COMP_STRIDE) * sizeof(double);
    cuda_emul::launch(setup_kernel, 2, 64, 0, pd, vp);
    bool k2 = true;
    for (int i = 0; i < fd.S_tot * pd.N; ++i) k2 = k2 && fd.patches[i].K == 2;
    if (n_blocks > 0) {
        if (k2)
            cuda_emul::launch(pixel_kernel<MODE, 2, true>, n_blocks, PIX_THREADS, smem, pd, chunk_pixels);
        else
            cuda_emul::launch(pixel_kernel<MODE, 0, true>, n_blocks, PIX_THREADS, smem, pd, chunk_pixels);
    }
    if (MODE == 2 && pd.n_pairs > 0) {
        const size_t psm = ((size_t)NPAIR_ACC * PAIR_THREADS + 2 * (size_t)MAX_COMPS * COMP_STRIDE) * sizeof(double);
        if (k2)
            cuda_emul::launch(pair_kernel<2>, pd.n_pairs * pd.N, PAIR_THREADS, psm, pd);
        else
            cuda_emul::launch(pair_kernel<0>, pd.n_pairs * pd.N, PAIR_THREADS, psm, pd);
    }
    cuda_emul::launch(epilogue_kernel<MODE>, pd.n_tasks, EPI_THREADS, 0, pd, vp, v, d, h, counters, flags, 0);
}
}  // namespace

extern "C" int emul_elbo_batch(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches,
                               int32_t n_tasks, const int32_t* task_ptr, const int32_t* source_ids,
                               const int32_t* active_ptr, const int32_t* active_idx, const double* vp, int32_t mode,
                               double* v, double* d, double* h, int64_t* counters, int32_t* flags,
                               int32_t chunk_pixels) {
    galaxy_prototypes(c_proto_eta, c_proto_nu);
    std::vector<ImageDev> images(N);
    std::vector<std::vector<double>> pixconst(N);
    for (int n = 0; n < N; ++n) {
        const celeste_image& im = imgs[n];
        pixconst[n].resize((size_t)im.H * im.W);
        cuda_emul::launch(prep_image_kernel, 2, 32, 0, im.H, im.W, im.pixels, im.nelec_per_nmgy, im.log_iota,
                          pixconst[n].data());
        images[n] = ImageDev{im.H, im.W, im.band, im.pixels, im.sky, im.nelec_per_nmgy, pixconst[n].data()};
    }
    std::vector<PatchDev> pdv((size_t)S_tot * N);
    for (size_t i = 0; i < pdv.size(); ++i) {
        const celeste_patch& q = patches[i];
        PatchDev p;
        p.off_h = (int)q.bitmap_offset[0];
        p.off_w = (int)q.bitmap_offset[1];
        p.H2 = q.H2;
        p.W2 = q.W2;
        p.bitmap = q.active_pixel_bitmap;
        for (int k = 0; k < 4; ++k) p.J[k] = q.wcs_jacobian[k];
        p.wc[0] = q.world_center[0];
        p.wc[1] = q.world_center[1];
        p.pc[0] = q.pixel_center[0];
        p.pc[1] = q.pixel_center[1];
        p.K = q.K;
        p.psf = q.psf;
        p.coefs = q.itp_coefs;
        p.n1 = q.itp_dims[0];
        p.n2 = q.itp_dims[1];
        pdv[i] = p;
    }
    const int n_slots = task_ptr[n_tasks];
    std::vector<int> src_row(n_slots), sub_ptr(n_tasks + 1, 0), sub_slot, sub_task, pair_ptr(n_tasks + 1, 0);
    std::vector<long long> h_ptr(n_tasks + 1, 0);
    std::vector<PairHdr> pairmap;
    std::vector<BlockHdr> blockmap;
    for (int t = 0; t < n_tasks; ++t) {
        const int Sa = active_ptr[t + 1] - active_ptr[t];
        if (Sa < 1 || Sa > 8) return CELESTE_ERR_UNSUPPORTED;
        sub_ptr[t + 1] = sub_ptr[t] + Sa;
        h_ptr[t + 1] = h_ptr[t] + (long long)(NPARAM * Sa) * (NPARAM * Sa);
        pair_ptr[t + 1] = pair_ptr[t] + Sa * (Sa - 1) / 2;
        for (int k = 0; k < Sa; ++k) {
            sub_slot.push_back(task_ptr[t] + active_idx[active_ptr[t] + k] - 1);
            sub_task.push_back(t);
        }
        for (int s = task_ptr[t]; s < task_ptr[t + 1]; ++s) src_row[s] = source_ids[s] - 1;
    }
    const int n_subs = (int)sub_slot.size();
    for (int t = 0; t < n_tasks; ++t) {
        const int Sa = sub_ptr[t + 1] - sub_ptr[t];
        for (int ka = 0; ka < Sa; ++ka)
            for (int kb = ka + 1; kb < Sa; ++kb)
                for (int n = 0; n < N; ++n)
                    pairmap.push_back(PairHdr{sub_ptr[t] + ka, sub_ptr[t] + kb, sub_slot[sub_ptr[t] + ka],
                                              sub_slot[sub_ptr[t] + kb], task_ptr[t], task_ptr[t + 1], n, 0});
    }
    std::vector<int> chunk_ptr((size_t)n_subs * N + 1, 0);
    for (int u = 0; u < n_subs; ++u)
        for (int n = 0; n < N; ++n) {
            const int t = sub_task[u];
            const size_t pidx = (size_t)src_row[sub_slot[u]] + (size_t)n * S_tot;
            const PatchDev& pa = pdv[pidx];
            const long npix = (long)pa.H2 * pa.W2;
            const int nchunk = (int)((npix + chunk_pixels - 1) / chunk_pixels);
            const int tn = u * N + n;
            chunk_ptr[tn + 1] = chunk_ptr[tn] + nchunk;
            for (int c = 0; c < nchunk; ++c)
                blockmap.push_back(BlockHdr{tn, c, sub_slot[u], task_ptr[t], task_ptr[t + 1], (int)pidx, n, 0, sub_ptr[t], u,
                                            sub_ptr[t + 1], t});
        }
    std::vector<int> tp(task_ptr, task_ptr + n_tasks + 1);
    std::vector<TaskHdr> taskmap;
    for (int u = 0; u < n_subs; ++u)
        for (int n0 = 0; n0 < N; n0 += TASK_NIMG) {
            const int t = sub_task[u];
            taskmap.push_back(TaskHdr{u * N, sub_slot[u], task_ptr[t], task_ptr[t + 1], 0, sub_ptr[t], u, sub_ptr[t + 1], n0,
                                      std::min(N, n0 + TASK_NIMG), t, 0});
        }
    std::vector<int> tcp((size_t)n_subs * N + 1);
    for (size_t i = 0; i < tcp.size(); ++i) tcp[i] = (int)(i * TASK_WARPS);
    std::vector<double> slotimg((size_t)n_slots * N * SLOTIMG_STRIDE), slotbr((size_t)n_slots * SLOTBR_STRIDE),
        partials(std::max({blockmap.size() * NACC_MODE2, (size_t)n_subs * N * TASK_WARPS * NACC_MODE1,
                           (size_t)1}) + 1),
        pair_partials(pairmap.size() * NPAIR_ACC + 1);
    PlanDev pd;
    FieldDev fd{images.data(), pdv.data(), S_tot, 0};
    std::vector<int> tfield(n_tasks, 0), sfield(n_slots, 0);
    pd.n_tasks = n_tasks;
    pd.N = N;
    pd.n_fields = 1;
    pd.n_slots = n_slots;
    pd.n_subs = n_subs;
    pd.n_pairs = (int)pairmap.size() / (N > 0 ? N : 1);
    pd.fields = &fd;
    pd.task_field = tfield.data();
    pd.slot_field = sfield.data();
    pd.task_ptr = tp.data();
    pd.src_row = src_row.data();
    pd.sub_ptr = sub_ptr.data();
    pd.sub_slot = sub_slot.data();
    pd.h_ptr = h_ptr.data();
    pd.blockmap = blockmap.data();
    pd.chunk_ptr = chunk_ptr.data();
    pd.pairmap = pairmap.data();
    pd.pair_ptr = pair_ptr.data();
    pd.slotimg = slotimg.data();
    pd.slotbr = slotbr.data();
    pd.partials = partials.data();
    pd.pair_partials = pair_partials.data();
    pd.task_mask = nullptr;
    pd.bg_ptr = nullptr;
    pd.bg = nullptr;
    std::vector<long long> cnt(2 * (size_t)n_tasks);
    const int nb = (int)blockmap.size();
    if (mode == 0)
        run_task<0>(pd, fd, taskmap, tcp, vp, v, d, h, cnt.data(), flags);
    else if (mode == 1)
        run_task<1>(pd, fd, taskmap, tcp, vp, v, d, h, cnt.data(), flags);
    else {
        bool all_k2 = true;
        for (size_t i = 0; i < pdv.size(); ++i) all_k2 = all_k2 && pdv[i].K == 2;
        if (g_hess_kernel == 1 && all_k2 && n_subs == n_tasks)
            run_unit<2>(pd, fd, vp, v, d, h, cnt.data(), flags);
        else
            run<2>(pd, fd, nb, chunk_pixels, vp, v, d, h, cnt.data(), flags);
    }
    for (size_t i = 0; i < cnt.size(); ++i) counters[i] = cnt[i];
    return 0;
}

// 0: task_kernel, 1: march_kernel, 2: march_kernel with every source split into one block per image, 3: unit_kernel
extern "C" int emul_set_grad_kernel(int32_t which) {
    g_grad_kernel = (which == 0 || which == 3) ? which : 1;
    g_march_split = which == 2 ? 1 : 6000;
    return 0;
}

extern "C" int emul_set_unit_target(int64_t target) {
    g_unit_target = (long)target;
    return 0;
}

// 0: pixel_kernel<2>, 1: unit_kernel<2> (where the product uses it)
extern "C" int emul_set_hess_kernel(int32_t which) {
    g_hess_kernel = which ? 1 : 0;
    return 0;
}

// build_march_blocks (the host half of march_kernels.cuh, shared with celeste_abi.cu) on its own: one task per
// source list, Sa = 1, active source first.  out: 12 ints per block (aslot, slot0, slot1, n0, n1, pidx, nseg, hasbg,
// and ubeg[1..4] are not needed by the checker: it gets ubeg[MARCH_NIMG] as `walks`) -- see tests/test_march_blocks.py.
extern "C" int emul_march_blocks(int32_t N, int32_t S_tot, const celeste_patch* patches, int32_t n_tasks,
                                 const int32_t* task_ptr, const int32_t* source_ids, int64_t split, int32_t capacity,
                                 int32_t* out, int32_t* part_ptr_out) {
    const int n_slots = task_ptr[n_tasks];
    std::vector<int> sub_task(n_tasks), sub_slot(n_tasks), src_row(n_slots);
    for (int t = 0; t < n_tasks; ++t) {
        sub_task[t] = t;
        sub_slot[t] = task_ptr[t];            // the active source is the first slot of its task
    }
    for (int s = 0; s < n_slots; ++s) src_row[s] = source_ids[s] - 1;
    std::vector<MarchHdr> mm;
    std::vector<int> part_ptr;
    build_march_blocks(n_tasks, N, sub_task.data(), sub_slot.data(), task_ptr, (const int*)nullptr,
                       [&](int slot, int n, int& oh, int& ow, int& H2, int& W2) {
                           const celeste_patch& q = patches[(size_t)src_row[slot] + (size_t)n * S_tot];
                           oh = (int)q.bitmap_offset[0];
                           ow = (int)q.bitmap_offset[1];
                           H2 = q.H2;
                           W2 = q.W2;
                       },
                       (long)split, mm, part_ptr);
    for (int t = 0; t <= n_tasks; ++t) part_ptr_out[t] = part_ptr[t];
    if ((int)mm.size() > capacity) return -(int)mm.size();
    for (size_t i = 0; i < mm.size(); ++i) {
        const MarchHdr& h = mm[i];
        int32_t* o = out + 12 * i;
        o[0] = h.aslot; o[1] = h.slot0; o[2] = h.slot1; o[3] = h.n0; o[4] = h.n1; o[5] = h.pidx; o[6] = h.nseg;
        o[7] = (int32_t)h.hasbg; o[8] = h.ubeg[MARCH_NIMG]; o[9] = h.sub; o[10] = h.task; o[11] = NPAIR;
    }
    return (int)mm.size();
}

extern "C" int emul_tr_subproblem(int32_t batch, int32_t n, const double* g, const double* H, const double* delta,
                                  double* s, double* m, int32_t* interior) {
    cuda_emul::launch(tr_subproblem_kernel, batch, TR_THREADS, 0, n, g, H, delta, (const unsigned char*)nullptr, s, m,
                      interior);
    return 0;
}

// newton_step_kernel under emulation; the buffers are host arrays
extern "C" int emul_newton_step(int32_t phase, int32_t batch, const celeste_newton_buffers* nb) {
    NewtonDev dev;
    dev.x = nb->x;
    dev.f = nb->f;
    dev.g = nb->g;
    dev.H = nb->H;
    dev.delta = nb->delta;
    dev.x_new = nb->x_new;
    dev.m_pred = nb->m_pred;
    dev.interior = nb->interior;
    dev.active = nb->active;
    dev.converged = nb->converged;
    dev.iters = nb->iters;
    dev.f_calls = nb->f_calls;
    dev.lo = nb->lo;
    dev.hi = nb->hi;
    dev.v = nb->v;
    dev.d = nb->d;
    dev.h = nb->h;
    dev.flags = nb->flags;
    dev.vp_all = nb->vp_all;
    dev.aslot = reinterpret_cast<const long long*>(nb->aslot);
    dev.prior = nb->prior;
    dev.h_layout = nb->h_layout;
    cuda_emul::launch(newton_step_kernel, batch, TR_THREADS, 0, dev, phase);
    return 0;
}

// render_kernel under emulation (one field, slots = the S sources in order)
static int emul_render_impl(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches, int32_t S,
                            const int32_t* source_ids, const double* vp, double* const* out, int full_box);
extern "C" int emul_render_expectation(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches,
                                       int32_t S, const int32_t* source_ids, const double* vp, double* const* out) {
    return emul_render_impl(N, imgs, S_tot, patches, S, source_ids, vp, out, 0);
}
extern "C" int emul_render_boxes(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches,
                                 int32_t S, const int32_t* source_ids, const double* vp, double* const* out) {
    return emul_render_impl(N, imgs, S_tot, patches, S, source_ids, vp, out, 1);
}
static int emul_render_impl(int32_t N, const celeste_image* imgs, int32_t S_tot, const celeste_patch* patches, int32_t S,
                            const int32_t* source_ids, const double* vp, double* const* out, int full_box) {
    galaxy_prototypes(c_proto_eta, c_proto_nu);
    std::vector<ImageDev> images(N);
    std::vector<int> imgH(N), imgW(N);
    for (int n = 0; n < N; ++n) {
        const celeste_image& im = imgs[n];
        images[n] = ImageDev{im.H, im.W, im.band, im.pixels, im.sky, im.nelec_per_nmgy, nullptr};
        imgH[n] = im.H;
        imgW[n] = im.W;
        std::fill(out[n], out[n] + (size_t)im.H * im.W, 0.0);
    }
    if (S == 0) return 0;
    std::vector<PatchDev> pdv((size_t)S_tot * N);
    for (size_t i = 0; i < pdv.size(); ++i) {
        const celeste_patch& q = patches[i];
        PatchDev p;
        p.off_h = (int)q.bitmap_offset[0];
        p.off_w = (int)q.bitmap_offset[1];
        p.H2 = q.H2;
        p.W2 = q.W2;
        p.bitmap = q.active_pixel_bitmap;
        for (int k = 0; k < 4; ++k) p.J[k] = q.wcs_jacobian[k];
        p.wc[0] = q.world_center[0];
        p.wc[1] = q.world_center[1];
        p.pc[0] = q.pixel_center[0];
        p.pc[1] = q.pixel_center[1];
        p.K = q.K;
        p.psf = q.psf;
        p.coefs = q.itp_coefs;
        p.n1 = q.itp_dims[0];
        p.n2 = q.itp_dims[1];
        pdv[i] = p;
    }
    std::vector<int> src_row(S), sfield(S, 0);
    for (int s = 0; s < S; ++s) src_row[s] = source_ids[s] - 1;
    std::vector<double> slotimg((size_t)S * N * SLOTIMG_STRIDE), slotbr((size_t)S * SLOTBR_STRIDE);
    FieldDev fd{images.data(), pdv.data(), S_tot, 0};
    PlanDev pd{};
    pd.n_tasks = 1;
    pd.N = N;
    pd.n_fields = 1;
    pd.n_slots = S;
    pd.fields = &fd;
    pd.slot_field = sfield.data();
    pd.src_row = src_row.data();
    pd.slotimg = slotimg.data();
    pd.slotbr = slotbr.data();
    cuda_emul::launch(setup_kernel, 2, 64, 0, pd, vp);
    std::vector<RenderTile> tiles;
    std::vector<int> tile_slots;
    build_render_tiles(N, imgH.data(), imgW.data(), S,
                       [&](int s, int n, int& oh, int& ow, int& H2, int& W2) {
                           const PatchDev& p = pdv[(size_t)src_row[s] + (size_t)n * S_tot];
                           oh = p.off_h;
                           ow = p.off_w;
                           H2 = p.H2;
                           W2 = p.W2;
                       },
                       tiles, tile_slots, full_box);
    bool k2 = true;
    for (size_t i = 0; i < pdv.size(); ++i) k2 = k2 && pdv[i].K == 2;
    if (!tiles.empty()) {
        if (k2)
            cuda_emul::launch(render_kernel<2>, (int)tiles.size(), RENDER_THREADS, 0, pd, tiles.data(), tile_slots.data(), out, full_box);
        else
            cuda_emul::launch(render_kernel<0>, (int)tiles.size(), RENDER_THREADS, 0, pd, tiles.data(), tile_slots.data(), out, full_box);
    }
    return 0;
}

// patch_kernels.cuh under emulation -----------------------------------------------------------------------------
extern "C" int emul_spline_build(int32_t grid_n, const double* raw, int32_t K, const double* psf, double* coefs) {
    SplineJob job{raw, psf, K, grid_n, coefs};
    const size_t sm = ((size_t)grid_n * grid_n + (size_t)(grid_n + 2) * grid_n) * sizeof(double);
    cuda_emul::launch(spline_build_kernel, 1, PB_THREADS, sm, (const SplineJob*)&job);
    return 0;
}

extern "C" int emul_bitmap_build(int32_t H, int32_t W, const float* pixels, int32_t off_h, int32_t off_w, int32_t H2,
                                 int32_t W2, uint8_t* bitmap) {
    ImageDev img{H, W, 1, pixels, nullptr, nullptr, nullptr};
    BitmapJob job{0, off_h, off_w, H2, W2, bitmap};
    cuda_emul::launch(bitmap_build_kernel, 1, 128, 0, (const ImageDev*)&img, (const BitmapJob*)&job);
    return 0;
}

// boxes: S x N x 4 ints (off_h, off_w, H2, W2), index (s + n * S) * 4.  Returns the CSR length; nbr may be null.
extern "C" int emul_find_neighbors(int32_t S, int32_t N, const int32_t* boxes, int32_t* nbr_ptr, int32_t* nbr) {
    std::vector<PatchDev> pdv((size_t)S * N);
    for (size_t i = 0; i < pdv.size(); ++i) {
        PatchDev p{};
        p.off_h = boxes[4 * i];
        p.off_w = boxes[4 * i + 1];
        p.H2 = boxes[4 * i + 2];
        p.W2 = boxes[4 * i + 3];
        pdv[i] = p;
    }
    std::vector<int> counts(S, 0), ptr(S + 1, 0);
    const int wpb = 4, blocks = (S + wpb - 1) / wpb;
    if (S > 0) cuda_emul::launch(neighbor_kernel, blocks, wpb * 32, 0, (const PatchDev*)pdv.data(), S, N, 0, counts.data(),
                                 (const int*)nullptr, (int*)nullptr);
    for (int t = 0; t < S; ++t) ptr[t + 1] = ptr[t] + counts[t];
    for (int t = 0; t <= S; ++t) nbr_ptr[t] = ptr[t];
    if (nbr && ptr[S] > 0)
        cuda_emul::launch(neighbor_kernel, blocks, wpb * 32, 0, (const PatchDev*)pdv.data(), S, N, 1, (int*)nullptr,
                          (const int*)ptr.data(), nbr);
    return ptr[S];
}
