// celeste_abi.cu -- C ABI (include/celeste_cuda.h) over the sm_100a kernels.
// Host-side plumbing only: handles, uploads, the task plan, kernel launches, status codes.
// No CPU implementation of the path exists in this library: without a CUDA device every
// compute entry point returns CELESTE_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/celeste_cuda.h"
#include "celeste_kernels.cuh"
#include "march_kernels.cuh"
#include "unit_kernels.cuh"
#include "maximize_kernels.cuh"
#include "patch_kernels.cuh"

using namespace celeste;

namespace {

thread_local char g_detail[512] = "";

void set_detail(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void set_detail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_detail, sizeof g_detail, fmt, ap);
    va_end(ap);
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            set_detail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
            return (_e == cudaErrorMemoryAllocation) ? CELESTE_ERR_ALLOC                           \
                   : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ||              \
                      _e == cudaErrorInitializationError)                                          \
                       ? CELESTE_ERR_NO_DEVICE                                                     \
                       : CELESTE_ERR_CUDA;                                                         \
        }                                                                                          \
    } while (0)

template <typename T>
struct DevBuf {                      // owns one cudaMalloc block: movable, not copyable
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) {
        o.p = nullptr;
        o.n = 0;
    }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& h) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

// light_source_model.jl:45-72
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
    const double er0 = 1.078031, er1 = 0.928896;
    for (int j = 0; j < 8; ++j) {
        eta[j] = dev_amp[j] / sd;
        nu[j] = dev_var[j] / (er0 * er0);
    }
    for (int j = 0; j < 6; ++j) {
        eta[8 + j] = exp_amp[j] / se;
        nu[8 + j] = exp_var[j] / (er1 * er1);
    }
}

int upload_constants() {
    double eta[NPROTO], nu[NPROTO];
    galaxy_prototypes(eta, nu);
    CUDA_TRY(cudaMemcpyToSymbol(c_proto_eta, eta, sizeof eta));
    CUDA_TRY(cudaMemcpyToSymbol(c_proto_nu, nu, sizeof nu));
    return CELESTE_OK;
}

int g_chunk_pixels = 0;

// cudaSetDevice for the duration of an ABI call; the caller's current device is restored on return
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;          // nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// multiprocessors of the current device (148 on a B200), queried once per device
int sm_count() {
    static std::mutex mu;
    static std::map<int, int> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
    return n;
}

template <int MODE>
size_t pixel_smem_bytes() {
    return ((size_t)NAcc<MODE>::value * PIX_THREADS + (size_t)MAX_COMPS * COMP_STRIDE) * sizeof(double);
}

template <int MODE>
size_t task_smem_bytes() {
    return ((size_t)NAcc<MODE>::value * PIX_THREADS + (size_t)TASK_NIMG * MAX_COMPS * COMP_STRIDE) * sizeof(double);
}

int configure_kernels() {
#define CEL_CFG(M, K)                                                                                                  \
    CUDA_TRY(cudaFuncSetAttribute(pixel_kernel<M, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                  (int)pixel_smem_bytes<M>()));                                                       \
    CUDA_TRY(cudaFuncSetAttribute(pixel_kernel<M, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                  (int)pixel_smem_bytes<M>()))
    CEL_CFG(2, 0);      // pixel_kernel serves the Hessian mode only; value / gradient use task_kernel
    CEL_CFG(2, 2);
#undef CEL_CFG
#define CEL_TCFG(M, K, X) \
    CUDA_TRY(cudaFuncSetAttribute(task_kernel<M, K, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)task_smem_bytes<M>()))
    CEL_TCFG(0, 0, false);
    CEL_TCFG(0, 0, true);
    CEL_TCFG(0, 2, false);
    CEL_TCFG(0, 2, true);
    CEL_TCFG(1, 0, false);
    CEL_TCFG(1, 0, true);
    CEL_TCFG(1, 2, false);
    CEL_TCFG(1, 2, true);
#undef CEL_TCFG
    CUDA_TRY(cudaFuncSetAttribute(march_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem_bytes()));
    CUDA_TRY(cudaFuncSetAttribute(march_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem_bytes()));
    CUDA_TRY(cudaFuncSetAttribute(unit_walk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)unit_smem_bytes<0>()));
    CUDA_TRY(cudaFuncSetAttribute(unit_walk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)unit_smem_bytes<1>()));
    CUDA_TRY(cudaFuncSetAttribute(unit_walk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)unit_smem_bytes<2>()));
    if (const char* env = std::getenv("CELESTE_MARCH_CARVEOUT")) {     // kernel-tuning knob: shared-memory share of L1, percent
        const int pct = std::atoi(env);
        if (pct >= 0 && pct <= 100) {
            CUDA_TRY(cudaFuncSetAttribute(march_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
            CUDA_TRY(cudaFuncSetAttribute(march_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        }
    }
    {
        const int psm = (int)(((size_t)NPAIR_ACC * PAIR_THREADS + 2 * (size_t)MAX_COMPS * COMP_STRIDE) * sizeof(double));
        CUDA_TRY(cudaFuncSetAttribute(pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
        CUDA_TRY(cudaFuncSetAttribute(pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm));
    }
    return CELESTE_OK;
}

int ensure_device_ready() {
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    static std::map<int, int> ready;
    auto it = ready.find(dev);
    if (it != ready.end()) return it->second;
    int st = upload_constants();
    if (st == CELESTE_OK) st = configure_kernels();
    ready[dev] = st;
    return st;
}

}  // namespace

struct celeste_field {
    int device = 0;
    int N = 0;
    int S_tot = 0;
    std::vector<ImageDev> h_images;
    std::vector<DevBuf<float>> pixels, sky, iota;
    std::vector<DevBuf<double>> pixconst;
    DevBuf<ImageDev> d_images;
    // patches
    std::vector<PatchDev> h_patches;
    DevBuf<PatchDev> d_patches;
    DevBuf<uint8_t> bitmap_pool;
    DevBuf<double> double_pool;   // psf records + spline coefficient arrays
    unsigned long long patch_generation = 0;
    int uniform_K = 0;   // K shared by every patch (0: mixed) -> selects the K-specialised pixel kernel
    // Plans of small celeste_elbo_batch / celeste_elbo_single calls, kept for the next call with the same task
    // structure: ElboMaximize.evaluate! (ElboMaximize.jl:161-172) calls elbo once per Newton iterate with the same
    // ElboArgs and a new vp, from every thread of the reference (one ElboArgs per thread, ParallelRun.jl:236-253).
    struct CachedPlan {
        std::vector<int32_t> key;
        celeste_plan* plan = nullptr;
        bool busy = false;
        unsigned long long last_use = 0;
    };
    std::mutex cache_mu;
    std::vector<CachedPlan> cache;
    unsigned long long cache_clock = 0;
    ~celeste_field();
};

struct celeste_plan {
    std::vector<celeste_field*> fields;
    std::vector<unsigned long long> patch_generation;
    int device = 0;
    int uniform_K = 0;
    DevBuf<FieldDev> d_fields;
    DevBuf<int> task_field, slot_field;
    int n_tasks = 0, n_slots = 0, N = 0;
    int n_blocks = 0, chunk_pixels = 0;
    std::vector<int> h_task_ptr;
    DevBuf<int> task_ptr, src_row, sub_ptr, sub_slot, chunk_ptr, pair_ptr;
    DevBuf<long long> h_ptr;
    DevBuf<BlockHdr> blockmap;
    DevBuf<TaskHdr> taskmap;          // value / gradient modes: one block per (sub, image group)
    DevBuf<int> task_chunk_ptr;       // partial ranges of task_kernel: TASK_WARPS per (sub, image)
    int n_taskblocks = 0;
    // march_kernel (value / gradient, Sa = 1, K = 2): one block per (sub, group of MARCH_NIMG images)
    DevBuf<MarchHdr> marchmap;
    DevBuf<int> march_part_ptr;      // n_subs + 1: the partial vectors (= blocks) of each sub
    DevBuf<long long> bg_ptr;
    DevBuf<double> bg;
    int n_marchblocks = 0;
    bool use_march = false;
    // unit kernels (every mode of the production shape; CELESTE_GRAD_KERNEL=march|task and CELESTE_HESS_KERNEL=pixel
    // select the round-1 kernels for A/B runs): one warp per unit -- rows of a (sub, image) -- pulled from a device-side queue
    DevBuf<UnitHdr> unitmap, unitmap_bg;   // every unit, heaviest first; the units with a neighbour, by shared pixels
    DevBuf<int> unit_chunk_ptr;      // the partial vectors (= units) of each (sub, image)
    DevBuf<int> unit_queue;          // one counter per kernel of the sequence
    DevBuf<double> bg_cnt;           // neighbour pixel-visits of each piece of unit_bg_kernel
    DevBuf<long long> l5_ptr;
    DevBuf<double> l5;               // L5 = dL/df1 of every active pixel (phase A -> phase B)
    DevBuf<PixRec> pix;              // pixel records of every unit in walk order (packed once, unit_pack_kernel)
    int n_units = 0, n_units_bg = 0, sms = 148;
    long long unit_maxpix = 1;
    bool use_unit_hess = false, use_unit_grad = false;
    bool small_plan = false;         // fewer than 8 whole (sub, image) units per resident warp: cut finer (see plan creation)
    bool block_epilogue = false;     // CELESTE_EPILOGUE=block: epilogue_kernel<2> instead of epilogue_hess_kernel (A/B knob)
    bool need_pack = false;          // pix is filled on the first evaluation (needs the uploaded plan arrays)
    DevBuf<PairHdr> pairmap;
    DevBuf<double> slotimg, slotbr, partials, pair_partials;
    int n_subs = 0, n_pairs = 0;
    size_t h_total = 0;      // doubles in the Hessian output: sum over tasks of (44 Sa)^2
    // staging for the host-buffer entry point
    DevBuf<double> vp_dev, d_dev, h_dev;
    cudaStream_t stream = nullptr;
    const unsigned char* task_mask = nullptr;   // device pointer owned by the caller
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_unit[2] = {nullptr, nullptr};   // after unit_bg_kernel, after unit_walk_kernel
    bool unit_timed = false;
    int hess_layout = CELESTE_HESS_DENSE;
    // small plans (what celeste_elbo_single makes): all outputs in ONE device block mirrored by one pinned host block
    // (a single D2H per call), and the kernel sequence of each mode captured once into a CUDA graph
    DevBuf<unsigned char> out_block;
    DevBuf<unsigned char> small_block;      // v | counters | flags of the large-plan host path
    unsigned char* small_pin = nullptr;
    unsigned char* out_pin = nullptr;
    double* vp_pin = nullptr;
    cudaGraphExec_t graph[3] = {nullptr, nullptr, nullptr};
    int graph_layout[3] = {-1, -1, -1};
    void drop_graphs() {
        for (auto& g : graph)
            if (g) {
                cudaGraphExecDestroy(g);
                g = nullptr;
            }
    }
    ~celeste_plan() {
        drop_graphs();
        if (out_pin) cudaFreeHost(out_pin);
        if (small_pin) cudaFreeHost(small_pin);
        if (vp_pin) cudaFreeHost(vp_pin);
        if (stream) cudaStreamDestroy(stream);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : ev_unit)
            if (e) cudaEventDestroy(e);
    }
};

celeste_field::~celeste_field() {
    for (auto& c : cache) delete c.plan;
}

extern "C" {

void celeste_get_errmsg(int status, char* buf) {
    const char* m;
    switch (status) {
        case CELESTE_OK: m = "ok"; break;
        case CELESTE_ERR_NO_DEVICE: m = "no usable CUDA device (no CPU fallback exists)"; break;
        case CELESTE_ERR_BAD_ARG: m = "invalid argument"; break;
        case CELESTE_ERR_ALLOC: m = "device memory allocation failed"; break;
        case CELESTE_ERR_CUDA: m = "CUDA runtime error"; break;
        case CELESTE_ERR_UNSUPPORTED: m = "unsupported configuration (keep the reference path)"; break;
        case CELESTE_ERR_NONFINITE: m = "ELBO value/gradient/Hessian contains Inf/NaNs"; break;
        case CELESTE_ERR_STATE: m = "handle in wrong state"; break;
        default: m = "unknown status"; break;
    }
    std::snprintf(buf, 61, "%s", m);
}

void celeste_get_errdetail(char* buf) { std::snprintf(buf, 512, "%s", g_detail); }

int celeste_version(void) { return 100; }

int celeste_init(int device, int* n_devices_out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (n_devices_out) *n_devices_out = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess || n == 0) {
        set_detail("cudaGetDeviceCount: %s (n=%d)", cudaGetErrorString(e), n);
        return CELESTE_ERR_NO_DEVICE;
    }
    if (device >= 0) {
        if (device >= n) {
            set_detail("device %d requested, %d present", device, n);
            return CELESTE_ERR_BAD_ARG;
        }
        CUDA_TRY(cudaSetDevice(device));
    }
    return ensure_device_ready();
}

int celeste_field_create(celeste_field** out, int32_t N, const celeste_image* imgs) {
    if (!out || N < 0 || (N > 0 && !imgs)) {
        set_detail("field_create: bad arguments");
        return CELESTE_ERR_BAD_ARG;
    }
    *out = nullptr;
    int st = ensure_device_ready();
    if (st != CELESTE_OK) return st;
    celeste_field* f = new (std::nothrow) celeste_field;
    if (!f) return CELESTE_ERR_ALLOC;
    std::unique_ptr<celeste_field> guard(f);
    CUDA_TRY(cudaGetDevice(&f->device));
    f->N = N;
    f->pixels.resize(N);
    f->sky.resize(N);
    f->iota.resize(N);
    f->pixconst.resize(N);
    f->h_images.resize(N);
    for (int n = 0; n < N; ++n) {
        const celeste_image& im = imgs[n];
        if (im.H <= 0 || im.W <= 0 || im.band < 1 || im.band > CELESTE_NUM_BANDS || !im.pixels || !im.sky ||
            !im.nelec_per_nmgy) {
            set_detail("field_create: image %d malformed (H=%d W=%d band=%d)", n, im.H, im.W, im.band);
            return CELESTE_ERR_BAD_ARG;
        }
        const size_t np = (size_t)im.H * im.W;
        CUDA_TRY(f->pixels[n].alloc(np));
        CUDA_TRY(f->sky[n].alloc(np));
        CUDA_TRY(f->iota[n].alloc(im.H));
        CUDA_TRY(f->pixconst[n].alloc(np));
        CUDA_TRY(cudaMemcpy(f->pixels[n].p, im.pixels, np * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(f->sky[n].p, im.sky, np * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(f->iota[n].p, im.nelec_per_nmgy, im.H * sizeof(float), cudaMemcpyHostToDevice));
        DevBuf<double> li;
        if (im.log_iota) {
            CUDA_TRY(li.alloc(im.H));
            CUDA_TRY(cudaMemcpy(li.p, im.log_iota, im.H * sizeof(double), cudaMemcpyHostToDevice));
        }
        prep_image_kernel<<<8 * sm_count(), 256>>>(im.H, im.W, f->pixels[n].p, f->iota[n].p, li.p, f->pixconst[n].p);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        ImageDev d;
        d.H = im.H;
        d.W = im.W;
        d.band = im.band;
        d.pixels = f->pixels[n].p;
        d.sky = f->sky[n].p;
        d.iota = f->iota[n].p;
        d.pixconst = f->pixconst[n].p;
        f->h_images[n] = d;
    }
    CUDA_TRY(f->d_images.upload(f->h_images));
    *out = guard.release();
    return CELESTE_OK;
}

int celeste_patches_set(celeste_field* f, int32_t S_tot, int32_t N, const celeste_patch* p) {
    if (!f || S_tot < 0 || N != f->N || (S_tot > 0 && N > 0 && !p)) {
        set_detail("patches_set: bad arguments (N=%d, field N=%d)", N, f ? f->N : -1);
        return CELESTE_ERR_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(f->device));
    const size_t np = (size_t)S_tot * N;
    // pool layout: bitmaps; doubles (psf records, de-duplicated spline coefficient arrays)
    std::vector<uint8_t> hb;
    std::vector<double> hd;
    std::vector<size_t> bm_off(np), psf_off(np), coef_off(np);
    std::map<std::pair<const double*, std::pair<int, int>>, size_t> coef_seen;
    for (size_t i = 0; i < np; ++i) {
        const celeste_patch& q = p[i];
        if (q.H2 < 0 || q.W2 < 0 || q.K < 1 || q.K > MAX_K || !q.psf || !q.itp_coefs || q.itp_dims[0] < 4 ||
            q.itp_dims[1] < 4 || ((size_t)q.H2 * q.W2 > 0 && !q.active_pixel_bitmap)) {
            set_detail("patches_set: patch %zu malformed (H2=%d W2=%d K=%d, K must be 1..%d)", i, q.H2, q.W2, q.K, MAX_K);
            return q.K > MAX_K ? CELESTE_ERR_UNSUPPORTED : CELESTE_ERR_BAD_ARG;
        }
        bm_off[i] = hb.size();
        const size_t nb = (size_t)q.H2 * q.W2;
        if (nb) hb.insert(hb.end(), q.active_pixel_bitmap, q.active_pixel_bitmap + nb);
        psf_off[i] = hd.size();
        hd.insert(hd.end(), q.psf, q.psf + 7 * (size_t)q.K);
        auto key = std::make_pair(q.itp_coefs, std::make_pair((int)q.itp_dims[0], (int)q.itp_dims[1]));
        auto it = coef_seen.find(key);
        if (it == coef_seen.end()) {
            coef_off[i] = hd.size();
            hd.insert(hd.end(), q.itp_coefs, q.itp_coefs + (size_t)q.itp_dims[0] * q.itp_dims[1]);
            coef_seen[key] = coef_off[i];
        } else {
            coef_off[i] = it->second;
        }
    }
    if (hb.empty()) hb.push_back(0);
    CUDA_TRY(f->bitmap_pool.upload(hb));
    CUDA_TRY(f->double_pool.upload(hd));
    f->h_patches.resize(np);
    for (size_t i = 0; i < np; ++i) {
        const celeste_patch& q = p[i];
        PatchDev d;
        d.off_h = (int)q.bitmap_offset[0];
        d.off_w = (int)q.bitmap_offset[1];
        d.H2 = q.H2;
        d.W2 = q.W2;
        d.bitmap = f->bitmap_pool.p + bm_off[i];
        for (int k = 0; k < 4; ++k) d.J[k] = q.wcs_jacobian[k];
        d.wc[0] = q.world_center[0];
        d.wc[1] = q.world_center[1];
        d.pc[0] = q.pixel_center[0];
        d.pc[1] = q.pixel_center[1];
        d.K = q.K;
        d.psf = f->double_pool.p + psf_off[i];
        d.coefs = f->double_pool.p + coef_off[i];
        d.n1 = q.itp_dims[0];
        d.n2 = q.itp_dims[1];
        f->h_patches[i] = d;
    }
    CUDA_TRY(f->d_patches.upload(f->h_patches));
    f->S_tot = S_tot;
    f->uniform_K = np ? p[0].K : 0;
    for (size_t i = 0; i < np; ++i)
        if (p[i].K != f->uniform_K) f->uniform_K = 0;
    f->patch_generation++;
    return CELESTE_OK;
}

int celeste_patches_build(celeste_field* f, int32_t S_tot, int32_t N, const celeste_patch_spec* specs) {
    if (!f || S_tot < 0 || N != f->N || (S_tot > 0 && N > 0 && !specs)) {
        set_detail("patches_build: bad arguments (N=%d, field N=%d)", N, f ? f->N : -1);
        return CELESTE_ERR_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(f->device));
    const size_t np = (size_t)S_tot * N;
    // pools: bitmaps (filled on the device); doubles = psf records + raw stamps (uploaded) + coefficient arrays (built)
    std::vector<double> hd;
    std::vector<size_t> bm_off(np), psf_off(np), coef_off(np);
    size_t bm_total = 0;
    struct JobKey {
        const double* raw;
        const double* psf;
        int K, n;
        bool operator<(const JobKey& o) const {
            return std::tie(raw, psf, K, n) < std::tie(o.raw, o.psf, o.K, o.n);
        }
    };
    std::map<JobKey, size_t> job_of;                 // unique stamps -> job index
    struct JobHost {
        size_t raw_off, psf_off, coef_off;
        bool has_raw;
        int K, n;
    };
    std::vector<JobHost> jobs;
    std::vector<size_t> patch_job(np);
    for (size_t i = 0; i < np; ++i) {
        const celeste_patch_spec& q = specs[i];
        const int n = i / (size_t)S_tot;
        const ImageDev& im = f->h_images[n];
        if (q.H2 < 0 || q.W2 < 0 || q.K < 1 || q.K > MAX_K || !q.psf || q.grid_n < 3 || q.grid_n > PB_MAX_GRID) {
            set_detail("patches_build: patch %zu malformed (H2=%d W2=%d K=%d grid_n=%d; K 1..%d, grid_n 3..%d)", i, q.H2, q.W2,
                       q.K, q.grid_n, MAX_K, PB_MAX_GRID);
            return (q.K > MAX_K || q.grid_n > PB_MAX_GRID) ? CELESTE_ERR_UNSUPPORTED : CELESTE_ERR_BAD_ARG;
        }
        if ((size_t)q.H2 * q.W2 > 0 && (q.bitmap_offset[0] < 0 || q.bitmap_offset[1] < 0 || q.bitmap_offset[0] + q.H2 > im.H ||
                                        q.bitmap_offset[1] + q.W2 > im.W)) {
            set_detail("patches_build: patch %zu: the box must be clamped to the image (clamp_box, imaged_sources.jl:10-14)", i);
            return CELESTE_ERR_BAD_ARG;
        }
        bm_off[i] = bm_total;
        bm_total += (size_t)q.H2 * q.W2;
        psf_off[i] = hd.size();
        hd.insert(hd.end(), q.psf, q.psf + 7 * (size_t)q.K);
        const JobKey key{q.grid_psf, q.grid_psf ? nullptr : q.psf, q.grid_psf ? 0 : q.K, q.grid_n};
        auto it = job_of.find(key);
        if (it == job_of.end()) {
            JobHost j;
            j.has_raw = q.grid_psf != nullptr;
            j.K = q.K;
            j.n = q.grid_n;
            j.psf_off = psf_off[i];
            j.raw_off = hd.size();
            if (j.has_raw) hd.insert(hd.end(), q.grid_psf, q.grid_psf + (size_t)q.grid_n * q.grid_n);
            j.coef_off = 0;
            job_of[key] = jobs.size();
            patch_job[i] = jobs.size();
            jobs.push_back(j);
        } else {
            patch_job[i] = it->second;
        }
    }
    for (auto& j : jobs) {                            // coefficient arrays behind everything uploaded
        j.coef_off = hd.size();
        hd.resize(hd.size() + (size_t)(j.n + 2) * (j.n + 2), 0.0);
    }
    CUDA_TRY(f->bitmap_pool.alloc(std::max<size_t>(bm_total, 1)));
    CUDA_TRY(f->double_pool.upload(hd));
    f->h_patches.resize(np);
    std::vector<BitmapJob> bjobs;
    for (size_t i = 0; i < np; ++i) {
        const celeste_patch_spec& q = specs[i];
        const JobHost& j = jobs[patch_job[i]];
        PatchDev d;
        d.off_h = (int)q.bitmap_offset[0];
        d.off_w = (int)q.bitmap_offset[1];
        d.H2 = q.H2;
        d.W2 = q.W2;
        d.bitmap = f->bitmap_pool.p + bm_off[i];
        for (int k = 0; k < 4; ++k) d.J[k] = q.wcs_jacobian[k];
        d.wc[0] = q.world_center[0];
        d.wc[1] = q.world_center[1];
        d.pc[0] = q.pixel_center[0];
        d.pc[1] = q.pixel_center[1];
        d.K = q.K;
        d.psf = f->double_pool.p + psf_off[i];
        d.coefs = f->double_pool.p + j.coef_off;
        d.n1 = d.n2 = j.n + 2;
        f->h_patches[i] = d;
        if ((size_t)q.H2 * q.W2 > 0)
            bjobs.push_back(BitmapJob{(int)(i / (size_t)std::max(S_tot, 1)), d.off_h, d.off_w, d.H2, d.W2,
                                      f->bitmap_pool.p + bm_off[i]});
    }
    std::vector<SplineJob> sjobs;
    int max_n = 3;
    for (const auto& j : jobs) {
        sjobs.push_back(SplineJob{j.has_raw ? f->double_pool.p + j.raw_off : nullptr, f->double_pool.p + j.psf_off, j.K, j.n,
                                  f->double_pool.p + j.coef_off});
        max_n = std::max(max_n, j.n);
    }
    DevBuf<SplineJob> d_sjobs;
    DevBuf<BitmapJob> d_bjobs;
    CUDA_TRY(d_sjobs.upload(sjobs));
    CUDA_TRY(d_bjobs.upload(bjobs));
    if (!sjobs.empty()) {
        const size_t sm = ((size_t)max_n * max_n + (size_t)(max_n + 2) * max_n) * sizeof(double);
        spline_build_kernel<<<(unsigned)sjobs.size(), PB_THREADS, sm>>>(d_sjobs.p);
        CUDA_TRY(cudaGetLastError());
    }
    if (!bjobs.empty()) {
        bitmap_build_kernel<<<(unsigned)bjobs.size(), 128>>>(f->d_images.p, d_bjobs.p);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(f->d_patches.upload(f->h_patches));
    CUDA_TRY(cudaDeviceSynchronize());
    f->S_tot = S_tot;
    f->uniform_K = np ? specs[0].K : 0;
    for (size_t i = 0; i < np; ++i)
        if (specs[i].K != f->uniform_K) f->uniform_K = 0;
    f->patch_generation++;
    return CELESTE_OK;
}

int celeste_patch_readback(celeste_field* f, int32_t s, int32_t n, int32_t* dims_out, uint8_t* bitmap, double* coefs) {
    if (!f || !dims_out || s < 0 || s >= f->S_tot || n < 0 || n >= f->N) {
        set_detail("patch_readback: bad arguments (s=%d n=%d)", s, n);
        return CELESTE_ERR_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(f->device));
    const PatchDev& p = f->h_patches[(size_t)s + (size_t)n * f->S_tot];
    dims_out[0] = p.H2;
    dims_out[1] = p.W2;
    dims_out[2] = p.n1;
    dims_out[3] = p.n2;
    if (bitmap && (size_t)p.H2 * p.W2 > 0)
        CUDA_TRY(cudaMemcpy(bitmap, p.bitmap, (size_t)p.H2 * p.W2, cudaMemcpyDeviceToHost));
    if (coefs) CUDA_TRY(cudaMemcpy(coefs, p.coefs, (size_t)p.n1 * p.n2 * sizeof(double), cudaMemcpyDeviceToHost));
    return CELESTE_OK;
}

int celeste_find_neighbors(celeste_field* f, int32_t* nbr_ptr, int32_t* nbr, int64_t capacity, int64_t* needed_out) {
    if (!f || !nbr_ptr || capacity < 0 || (capacity > 0 && !nbr)) {
        set_detail("find_neighbors: bad arguments");
        return CELESTE_ERR_BAD_ARG;
    }
    CUDA_TRY(cudaSetDevice(f->device));
    const int S = f->S_tot;
    nbr_ptr[0] = 0;
    if (needed_out) *needed_out = 0;
    if (S == 0) return CELESTE_OK;
    DevBuf<int> d_counts, d_ptr, d_out;
    CUDA_TRY(d_counts.alloc(S));
    const int wpb = 4, blocks = (S + wpb - 1) / wpb;
    neighbor_kernel<<<blocks, wpb * 32>>>(f->d_patches.p, S, f->N, 0, d_counts.p, nullptr, nullptr);
    CUDA_TRY(cudaGetLastError());
    std::vector<int> counts(S), ptr(S + 1, 0);
    CUDA_TRY(cudaMemcpy(counts.data(), d_counts.p, S * sizeof(int), cudaMemcpyDeviceToHost));
    for (int t = 0; t < S; ++t) ptr[t + 1] = ptr[t] + counts[t];
    for (int t = 0; t <= S; ++t) nbr_ptr[t] = ptr[t];
    if (needed_out) *needed_out = ptr[S];
    if (ptr[S] > capacity) {
        set_detail("find_neighbors: %d entries needed, capacity %lld", ptr[S], (long long)capacity);
        return CELESTE_ERR_BAD_ARG;
    }
    if (ptr[S] == 0) return CELESTE_OK;
    CUDA_TRY(d_ptr.upload(ptr));
    CUDA_TRY(d_out.alloc(ptr[S]));
    neighbor_kernel<<<blocks, wpb * 32>>>(f->d_patches.p, S, f->N, 1, nullptr, d_ptr.p, d_out.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(nbr, d_out.p, (size_t)ptr[S] * sizeof(int), cudaMemcpyDeviceToHost));
    return CELESTE_OK;
}

void celeste_field_destroy(celeste_field* f) { delete f; }

int celeste_plan_create_multi(int32_t n_fields, celeste_field* const* fields, celeste_plan** out, int32_t n_tasks,
                              const int32_t* task_field, const int32_t* task_ptr, const int32_t* source_ids,
                              const int32_t* active_ptr, const int32_t* active_idx) {
    if (n_fields < 1 || !fields || !out || n_tasks < 0 || !task_ptr || !active_ptr ||
        (n_tasks > 0 && (!source_ids || !active_idx || (n_fields > 1 && !task_field)))) {
        set_detail("plan_create: bad arguments");
        return CELESTE_ERR_BAD_ARG;
    }
    *out = nullptr;
    for (int i = 0; i < n_fields; ++i)
        if (!fields[i] || fields[i]->N != fields[0]->N || fields[i]->device != fields[0]->device) {
            set_detail("plan_create: field %d is null or differs in image count / device from field 0", i);
            return CELESTE_ERR_BAD_ARG;
        }
    CUDA_TRY(cudaSetDevice(fields[0]->device));
    std::unique_ptr<celeste_plan> pl(new (std::nothrow) celeste_plan);
    if (!pl) return CELESTE_ERR_ALLOC;
    pl->fields.assign(fields, fields + n_fields);
    pl->device = fields[0]->device;
    pl->uniform_K = fields[0]->uniform_K;
    std::vector<FieldDev> hf(n_fields);
    for (int i = 0; i < n_fields; ++i) {
        pl->patch_generation.push_back(fields[i]->patch_generation);
        if (fields[i]->uniform_K != pl->uniform_K) pl->uniform_K = 0;
        hf[i].images = fields[i]->d_images.p;
        hf[i].patches = fields[i]->d_patches.p;
        hf[i].S_tot = fields[i]->S_tot;
        hf[i].pad = 0;
    }
    pl->n_tasks = n_tasks;
    pl->N = fields[0]->N;
    const int n_slots = task_ptr[n_tasks];
    pl->n_slots = n_slots;
    pl->h_task_ptr.assign(task_ptr, task_ptr + n_tasks + 1);
    std::vector<int> src_row(n_slots), tfield(n_tasks), sfield(n_slots);
    std::vector<int> sub_ptr(n_tasks + 1, 0), sub_slot, sub_task, pair_ptr(n_tasks + 1, 0);
    std::vector<long long> h_ptr(n_tasks + 1, 0);
    std::vector<PairHdr> pairmap;
    for (int t = 0; t < n_tasks; ++t) {
        const int s0 = task_ptr[t], s1 = task_ptr[t + 1];
        if (s0 < 0 || s1 < s0 || (t == 0 && s0 != 0)) {
            set_detail("plan_create: task_ptr not a prefix array at task %d", t);
            return CELESTE_ERR_BAD_ARG;
        }
        const int fi = task_field ? task_field[t] : 0;
        if (fi < 0 || fi >= n_fields) {
            set_detail("plan_create: task %d names field %d of %d", t, fi, n_fields);
            return CELESTE_ERR_BAD_ARG;
        }
        tfield[t] = fi;
        const int Sa = active_ptr[t + 1] - active_ptr[t];
        if (Sa < 1 || Sa > 8) {
            set_detail("plan_create: task %d has Sa=%d active sources (supported: 1..8; production uses 1, "
                       "ParallelRun.jl:253,489)", t, Sa);
            return Sa < 1 ? CELESTE_ERR_BAD_ARG : CELESTE_ERR_UNSUPPORTED;
        }
        sub_ptr[t + 1] = sub_ptr[t] + Sa;
        h_ptr[t + 1] = h_ptr[t] + (long long)(NPARAM * Sa) * (NPARAM * Sa);
        for (int k = 0; k < Sa; ++k) {
            const int a = active_idx[active_ptr[t] + k];
            if (a < 1 || a > s1 - s0) {
                set_detail("plan_create: task %d active index %d outside 1..%d", t, a, s1 - s0);
                return CELESTE_ERR_BAD_ARG;
            }
            for (int k2 = 0; k2 < k; ++k2)
                if (active_idx[active_ptr[t] + k2] == a) {
                    set_detail("plan_create: task %d lists active source %d twice", t, a);
                    return CELESTE_ERR_BAD_ARG;
                }
            sub_slot.push_back(s0 + a - 1);
            sub_task.push_back(t);
        }
        for (int s = s0; s < s1; ++s) {
            const int row = source_ids[s];
            if (row < 1 || row > fields[fi]->S_tot) {
                set_detail("plan_create: task %d source id %d outside 1..%d", t, row, fields[fi]->S_tot);
                return CELESTE_ERR_BAD_ARG;
            }
            src_row[s] = row - 1;
            sfield[s] = fi;
        }
        pair_ptr[t + 1] = pair_ptr[t] + Sa * (Sa - 1) / 2;
    }
    const int n_subs = (int)sub_slot.size();
    pl->n_subs = n_subs;
    pl->n_pairs = pair_ptr[n_tasks];
    pl->h_total = (size_t)h_ptr[n_tasks];
    for (int t = 0; t < n_tasks; ++t) {
        const int Sa = sub_ptr[t + 1] - sub_ptr[t];
        for (int ka = 0; ka < Sa; ++ka)
            for (int kb = ka + 1; kb < Sa; ++kb)
                for (int n = 0; n < pl->N; ++n) {
                    PairHdr ph;
                    ph.sub_a = sub_ptr[t] + ka;
                    ph.sub_b = sub_ptr[t] + kb;
                    ph.slot_a = sub_slot[ph.sub_a];
                    ph.slot_b = sub_slot[ph.sub_b];
                    ph.slot0 = task_ptr[t];
                    ph.slot1 = task_ptr[t + 1];
                    ph.n = n;
                    ph.field = tfield[t];
                    pairmap.push_back(ph);
                }
    }
    // block map: one block per (active source of a task, image, chunk of its patch's pixels)
    int chunk_pixels = g_chunk_pixels > 0 ? g_chunk_pixels : 4 * PIX_THREADS;
    if (const char* env = std::getenv("CELESTE_CHUNK_PIXELS"))   // kernel-tuning knob
        if (std::atoi(env) > 0) chunk_pixels = std::atoi(env);
    pl->chunk_pixels = chunk_pixels;
    std::vector<int> chunk_ptr((size_t)n_subs * pl->N + 1, 0);
    std::vector<BlockHdr> blockmap;
    for (int u = 0; u < n_subs; ++u)
        for (int n = 0; n < pl->N; ++n) {
            const int t = sub_task[u];
            const celeste_field* f = fields[tfield[t]];
            const size_t pidx = (size_t)src_row[sub_slot[u]] + (size_t)n * f->S_tot;
            const PatchDev& pa = f->h_patches[pidx];
            const long npix = (long)pa.H2 * pa.W2;
            const int nchunk = (int)((npix + chunk_pixels - 1) / chunk_pixels);
            const int tn = u * pl->N + n;
            chunk_ptr[tn + 1] = chunk_ptr[tn] + nchunk;
            for (int c = 0; c < nchunk; ++c) {
                BlockHdr hd;
                hd.tn = tn;
                hd.chunk = c;
                hd.aslot = sub_slot[u];
                hd.slot0 = task_ptr[t];
                hd.slot1 = task_ptr[t + 1];
                hd.patch = (int)pidx;
                hd.n = n;
                hd.field = tfield[t];
                hd.sub0 = sub_ptr[t];
                hd.sub = u;
                hd.sub1 = sub_ptr[t + 1];
                hd.task = t;
                blockmap.push_back(hd);
            }
        }
    // value / gradient modes: task-level blocks (task_kernel), heaviest first so the launch tail is short
    std::vector<TaskHdr> taskmap;
    std::vector<long> taskcost;
    for (int u = 0; u < n_subs; ++u) {
        const int t = sub_task[u];
        const celeste_field* f = fields[tfield[t]];
        for (int n0 = 0; n0 < pl->N; n0 += TASK_NIMG) {
            TaskHdr th;
            th.tn0 = u * pl->N;
            th.aslot = sub_slot[u];
            th.slot0 = task_ptr[t];
            th.slot1 = task_ptr[t + 1];
            th.field = tfield[t];
            th.sub0 = sub_ptr[t];
            th.sub = u;
            th.sub1 = sub_ptr[t + 1];
            th.n0 = n0;
            th.n1 = std::min(pl->N, n0 + TASK_NIMG);
            th.task = t;
            th.pad1 = 0;
            long cost = 0;
            for (int n = th.n0; n < th.n1; ++n) {
                const PatchDev& pa = f->h_patches[(size_t)src_row[sub_slot[u]] + (size_t)n * f->S_tot];
                cost += (long)pa.H2 * pa.W2 * (1 + (th.slot1 - th.slot0 - 1) / 4);
            }
            taskmap.push_back(th);
            taskcost.push_back(cost);
        }
    }
    {
        std::vector<int> order(taskmap.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return taskcost[x] > taskcost[y]; });
        std::vector<TaskHdr> sorted(taskmap.size());
        for (size_t i = 0; i < order.size(); ++i) sorted[i] = taskmap[order[i]];
        taskmap.swap(sorted);
    }
    pl->n_taskblocks = (int)taskmap.size();
    // The production shape (Sa = 1 everywhere, K = 2: ParallelRun.jl:253,489, elbo_args.jl:197) is served by the unit
    // kernels in every mode.  A/B knobs: CELESTE_GRAD_KERNEL=march (march_kernel) / =task (task_kernel) for the value
    // and gradient modes, CELESTE_HESS_KERNEL=pixel (pixel_kernel<2>) for the Hessian mode.  Every other shape keeps
    // task_kernel / pixel_kernel.
    {
        bool shape_ok = pl->uniform_K == 2 && n_subs == n_tasks;
        // the walk kernels index pixels with 32-bit integers (registers are what limits them)
        for (int i = 0; i < n_fields && shape_ok; ++i) {
            for (const ImageDev& im : fields[i]->h_images)
                if ((long long)im.H * im.W >= (1LL << 31)) shape_ok = false;
            for (const PatchDev& pa : fields[i]->h_patches)
                if ((long long)pa.H2 * pa.W2 >= (1LL << 29)) shape_ok = false;
        }
        const char* genv = std::getenv("CELESTE_GRAD_KERNEL");
        const char* henv = std::getenv("CELESTE_HESS_KERNEL");
        const bool g_march = genv && std::strcmp(genv, "march") == 0, g_task = genv && std::strcmp(genv, "task") == 0;
        pl->use_march = shape_ok && g_march;
        pl->use_unit_grad = shape_ok && !g_march && !g_task;
        pl->use_unit_hess = shape_ok && !(henv && std::strcmp(henv, "pixel") == 0);
        const char* eenv = std::getenv("CELESTE_EPILOGUE");
        pl->block_epilogue = eenv && std::strcmp(eenv, "block") == 0;
    }
    if (pl->use_march) {
        // A source normally gets ONE block (all its images: best packing of its rows into the block's walk slots).
        // When the plan is small against the GPU, the heaviest sources would then be the whole tail of the launch
        // (a 51 x 51 x 5-pixel source alone on an SM runs for longer than the rest of a 1000-source plan), so sources
        // above `split` patch pixels are cut into smaller image groups: split = the plan's pixels per resident block
        // slot (SMs x 3; CELESTE_MARCH_SPLIT_PCT scales it), never below 1500.  CELESTE_MARCH_SPLIT=<pixels> overrides
        // (kernel-tuning knobs).
        long split;
        {
            long total_px = 0;
            for (int u = 0; u < n_subs; ++u) {
                const celeste_field* f = fields[tfield[sub_task[u]]];
                for (int n = 0; n < pl->N; ++n) {
                    const PatchDev& pa = f->h_patches[(size_t)src_row[sub_slot[u]] + (size_t)n * f->S_tot];
                    total_px += (long)std::max(pa.H2, 0) * std::max(pa.W2, 0);
                }
            }
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
            const long alpha_pct = std::getenv("CELESTE_MARCH_SPLIT_PCT") ? std::atol(std::getenv("CELESTE_MARCH_SPLIT_PCT")) : 100;
            split = std::max(1500L, total_px * std::max(1L, alpha_pct) / (100L * sms * CELESTE_MARCH_MINB));
        }
        if (const char* env = std::getenv("CELESTE_MARCH_SPLIT"))
            if (std::atol(env) > 0) split = std::atol(env);
        std::vector<MarchHdr> mm;
        std::vector<int> part_ptr;
        build_march_blocks(n_subs, pl->N, sub_task.data(), sub_slot.data(), task_ptr, tfield.data(),
                           [&](int slot, int n, int& oh, int& ow, int& H2, int& W2) {
                               const celeste_field* f = fields[sfield[slot]];
                               const PatchDev& pa = f->h_patches[(size_t)src_row[slot] + (size_t)n * f->S_tot];
                               oh = pa.off_h;
                               ow = pa.off_w;
                               H2 = pa.H2;
                               W2 = pa.W2;
                           },
                           split, mm, part_ptr);
        pl->n_marchblocks = (int)mm.size();
        CUDA_TRY(pl->marchmap.upload(mm));
        CUDA_TRY(pl->march_part_ptr.upload(part_ptr));
    }
    if (pl->use_march || pl->use_unit_grad || pl->use_unit_hess) {
        // background planes (E_bg, V_bg) of every (sub, image) whose task has a neighbour
        std::vector<long long> bg_ptr((size_t)n_subs * pl->N, -1);
        long long bg_total = 0;
        for (int u = 0; u < n_subs; ++u) {
            const int t = sub_task[u];
            if (task_ptr[t + 1] - task_ptr[t] < 2) continue;
            const celeste_field* f = fields[tfield[t]];
            for (int n = 0; n < pl->N; ++n) {
                const PatchDev& pa = f->h_patches[(size_t)src_row[sub_slot[u]] + (size_t)n * f->S_tot];
                if (pa.H2 <= 0 || pa.W2 <= 0) continue;
                bg_ptr[(size_t)u * pl->N + n] = bg_total;
                bg_total += 2LL * pa.H2 * pa.W2;
            }
        }
        CUDA_TRY(pl->bg_ptr.upload(bg_ptr));
        CUDA_TRY(pl->bg.alloc((size_t)bg_total));
    }
    if (pl->use_unit_hess || pl->use_unit_grad) {
        std::vector<UnitHdr> um, ub;
        std::vector<int> ucp;
        cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, pl->device);
        // Rows per unit.  A plan that keeps every resident warp busy with >= 8 whole (sub, image) units is not cut (the
        // per-unit prologue costs 5-20 % there: profiles/tuning_r02.md); a smaller plan -- a single celeste_elbo_single
        // call, one rank's share of an 8-GPU run -- is cut at CELESTE_UNIT_ROWS rows so that its sources spread over the
        // whole GPU and the launch has no tail.  Within either regime the cut depends on the patch only, so a task's
        // result is bit-for-bit independent of what else is in the plan.  CELESTE_UNIT_ROWS=<n> forces n (0 = never).
        pl->small_plan = (long)n_subs * pl->N < 8L * pl->sms * CELESTE_UNIT_MINB * UNIT_WARPS;
        long target = pl->small_plan ? CELESTE_UNIT_ROWS : 0;
        if (const char* env = std::getenv("CELESTE_UNIT_ROWS")) target = std::atol(env);
        build_unit_list(n_subs, pl->N, sub_task.data(), sub_slot.data(), task_ptr, tfield.data(),
                        [&](int slot, int n, int& oh, int& ow, int& H2, int& W2) {
                            const celeste_field* f = fields[sfield[slot]];
                            const PatchDev& pa = f->h_patches[(size_t)src_row[slot] + (size_t)n * f->S_tot];
                            oh = pa.off_h;
                            ow = pa.off_w;
                            H2 = pa.H2;
                            W2 = pa.W2;
                        },
                        target, std::getenv("CELESTE_UNIT_PIXELS") ? std::atol(std::getenv("CELESTE_UNIT_PIXELS")) : 0L,
                        pl->small_plan ? UNIT_BG_PIXELS_SMALL : UNIT_BG_PIXELS_BIG, um, ub, ucp, pl->unit_maxpix);
        pl->n_units = (int)um.size();
        pl->n_units_bg = (int)ub.size();
        std::vector<long long> l5_ptr((size_t)n_subs * pl->N, 0);
        long long l5_total = 0;
        for (int u = 0; u < n_subs; ++u) {
            const celeste_field* f = fields[tfield[sub_task[u]]];
            for (int n = 0; n < pl->N; ++n) {
                const PatchDev& pa = f->h_patches[(size_t)src_row[sub_slot[u]] + (size_t)n * f->S_tot];
                l5_ptr[(size_t)u * pl->N + n] = l5_total;
                l5_total += (long long)std::max(pa.H2, 0) * std::max(pa.W2, 0);
            }
        }
        CUDA_TRY(pl->unitmap.upload(um));
        CUDA_TRY(pl->unitmap_bg.upload(ub));
        CUDA_TRY(pl->unit_chunk_ptr.upload(ucp));
        CUDA_TRY(pl->unit_queue.alloc(4));
        CUDA_TRY(pl->bg_cnt.alloc(std::max<size_t>(ub.size(), 1)));
        CUDA_TRY(pl->l5_ptr.upload(l5_ptr));
        CUDA_TRY(pl->l5.alloc(pl->use_unit_hess ? (size_t)l5_total : 0));
        CUDA_TRY(pl->pix.alloc((size_t)l5_total));
        pl->need_pack = true;
    }
    std::vector<int> task_chunk_ptr((size_t)n_subs * pl->N + 1);
    for (size_t i = 0; i < task_chunk_ptr.size(); ++i) task_chunk_ptr[i] = (int)(i * TASK_WARPS);
    CUDA_TRY(pl->taskmap.upload(taskmap));
    CUDA_TRY(pl->task_chunk_ptr.upload(task_chunk_ptr));
    pl->n_blocks = (int)blockmap.size();
    std::vector<int> tp(task_ptr, task_ptr + n_tasks + 1);
    CUDA_TRY(pl->d_fields.upload(hf));
    CUDA_TRY(pl->task_field.upload(tfield));
    CUDA_TRY(pl->slot_field.upload(sfield));
    CUDA_TRY(pl->task_ptr.upload(tp));
    CUDA_TRY(pl->src_row.upload(src_row));
    CUDA_TRY(pl->sub_ptr.upload(sub_ptr));
    CUDA_TRY(pl->sub_slot.upload(sub_slot));
    CUDA_TRY(pl->h_ptr.upload(h_ptr));
    CUDA_TRY(pl->pair_ptr.upload(pair_ptr));
    CUDA_TRY(pl->pairmap.upload(pairmap));
    CUDA_TRY(pl->pair_partials.alloc(pairmap.size() * NPAIR_ACC));
    CUDA_TRY(pl->chunk_ptr.upload(chunk_ptr));
    CUDA_TRY(pl->blockmap.upload(blockmap));
    CUDA_TRY(pl->slotimg.alloc((size_t)n_slots * pl->N * SLOTIMG_STRIDE));
    CUDA_TRY(pl->slotbr.alloc((size_t)n_slots * SLOTBR_STRIDE));
    CUDA_TRY(pl->partials.alloc(std::max({(size_t)pl->n_blocks * NACC_MODE2, (size_t)n_subs * pl->N * TASK_WARPS * NACC_MODE1,
                                           (size_t)pl->n_marchblocks * NT_ACC, (size_t)pl->n_units * NACC_MODE2})));
    *out = pl.release();
    return CELESTE_OK;
}

int celeste_plan_create(celeste_field* f, celeste_plan** out, int32_t n_tasks, const int32_t* task_ptr,
                        const int32_t* source_ids, const int32_t* active_ptr, const int32_t* active_idx) {
    celeste_field* fields[1] = {f};
    return celeste_plan_create_multi(1, fields, out, n_tasks, nullptr, task_ptr, source_ids, active_ptr, active_idx);
}

void celeste_plan_destroy(celeste_plan* p) { delete p; }

int celeste_plan_set_task_mask(celeste_plan* p, const uint8_t* mask_dev) {
    if (!p) return CELESTE_ERR_BAD_ARG;
    p->task_mask = mask_dev;
    return CELESTE_OK;
}

int celeste_set_chunk_pixels(int32_t chunk_pixels) {
    if (chunk_pixels < 0) return CELESTE_ERR_BAD_ARG;
    g_chunk_pixels = chunk_pixels;
    return CELESTE_OK;
}

int celeste_plan_enable_timing(celeste_plan* p, int32_t on) {
    if (!p) return CELESTE_ERR_BAD_ARG;
    if (on) {
        for (auto& e : p->ev)
            if (!e) CUDA_TRY(cudaEventCreate(&e));
        for (auto& e : p->ev_unit)
            if (!e) CUDA_TRY(cudaEventCreate(&e));
    }
    p->timing = on != 0;
    return CELESTE_OK;
}

int celeste_plan_unit_times(celeste_plan* p, float ms[3]) {
    if (!p || !ms || !p->timing) {
        set_detail("plan_unit_times: timing not enabled");
        return CELESTE_ERR_STATE;
    }
    ms[0] = ms[1] = ms[2] = 0.f;
    if (!p->unit_timed) return CELESTE_OK;
    CUDA_TRY(cudaEventSynchronize(p->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms[0], p->ev[1], p->ev_unit[0]));
    CUDA_TRY(cudaEventElapsedTime(&ms[1], p->ev_unit[0], p->ev_unit[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms[2], p->ev_unit[1], p->ev[2]));
    return CELESTE_OK;
}

int celeste_plan_set_hessian_layout(celeste_plan* p, int32_t layout) {
    if (!p || (layout != CELESTE_HESS_DENSE && layout != CELESTE_HESS_PACKED28)) {
        set_detail("plan_set_hessian_layout: bad arguments (layout=%d)", layout);
        return CELESTE_ERR_BAD_ARG;
    }
    if (layout == CELESTE_HESS_PACKED28 && p->n_subs != p->n_tasks) {
        set_detail("plan_set_hessian_layout: the packed layout needs Sa = 1 in every task");
        return CELESTE_ERR_UNSUPPORTED;
    }
    p->hess_layout = layout;      // (captured graphs carry their layout and are re-captured when it differs)
    return CELESTE_OK;
}

int celeste_plan_kernel_times(celeste_plan* p, float ms[3]) {
    if (!p || !ms || !p->timing) {
        set_detail("plan_kernel_times: timing not enabled");
        return CELESTE_ERR_STATE;
    }
    CUDA_TRY(cudaEventSynchronize(p->ev[3]));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
    return CELESTE_OK;
}

int celeste_plan_launches(const celeste_plan* p, int32_t mode) {
    if (!p || p->n_tasks == 0) return 0;
    if ((mode <= 1 && p->use_unit_grad) || (mode == 2 && p->use_unit_hess))       // [slotbr,] [bg,] walk, [moment,] epilogue
        return 1 + (p->n_units_bg > 0 ? 1 : 0) + 1 + (mode == 2 ? 1 : 0) + 1;
    if (mode <= 1 && p->use_march) return 2;                                       // march, epilogue
    return (p->n_blocks > 0 ? 3 : 2) + ((mode == 2 && p->n_pairs > 0) ? 1 : 0);   // setup, pixel, [pair,] epilogue
}

int celeste_plan_kernel_name(const celeste_plan* p, int32_t mode, char* buf) {
    if (!p || !buf || mode < 0 || mode > 2) return CELESTE_ERR_BAD_ARG;
    const char* nm = mode == 2 ? (p->use_unit_hess ? "unit_kernel" : "pixel_kernel")
                               : (p->use_unit_grad ? "unit_kernel" : (p->use_march ? "march_kernel" : "task_kernel"));
    std::snprintf(buf, 32, "%s", nm);
    return CELESTE_OK;
}

}  // extern "C"

static PlanDev plan_dev(const celeste_plan* p) {
    PlanDev d;
    d.n_tasks = p->n_tasks;
    d.N = p->N;
    d.n_fields = (int)p->fields.size();
    d.n_slots = p->n_slots;
    d.fields = p->d_fields.p;
    d.task_field = p->task_field.p;
    d.slot_field = p->slot_field.p;
    d.task_ptr = p->task_ptr.p;
    d.src_row = p->src_row.p;
    d.n_subs = p->n_subs;
    d.n_pairs = p->n_pairs;
    d.sub_ptr = p->sub_ptr.p;
    d.sub_slot = p->sub_slot.p;
    d.h_ptr = p->h_ptr.p;
    d.blockmap = p->blockmap.p;
    d.chunk_ptr = p->chunk_ptr.p;
    d.pairmap = p->pairmap.p;
    d.pair_ptr = p->pair_ptr.p;
    d.pair_partials = p->pair_partials.p;
    d.task_mask = p->task_mask;
    d.slotimg = p->slotimg.p;
    d.slotbr = p->slotbr.p;
    d.partials = p->partials.p;
    d.bg_ptr = p->bg_ptr.p;
    d.bg = p->bg.p;
    d.bg_cnt = p->bg_cnt.p;
    d.l5_ptr = p->l5_ptr.p;
    d.l5 = p->l5.p;
    d.pix = p->pix.p;
    return d;
}

template <int MODE>
static int launch_mode(celeste_plan* p, const double* vp_dev, double* v, double* d, double* h, long long* counters,
                       int* flags, cudaStream_t st) {
    const PlanDev pd = plan_dev(p);
    const long total = (long)p->n_slots * p->N * MAX_K;
    const int sblocks = (int)std::max<long>(1, std::min<long>((total + 127) / 128, 16L * sm_count()));
    if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[0], st));
    if ((MODE <= 1 && p->use_unit_grad) || (MODE == 2 && p->use_unit_hess)) {
        // one warp per (sub, image) unit from a device-side queue (unit_kernels.cuh); the epilogue is the general one
        PlanDev pu = pd;
        pu.chunk_ptr = p->unit_chunk_ptr.p;
        if (p->need_pack) {
            if (p->n_units > 0) unit_pack_kernel<<<p->n_units, 128, 0, st>>>(pu, p->unitmap.p, p->n_units, p->pix.p);
            p->need_pack = false;
        }
        slotbr_kernel<<<(std::max(p->n_slots, 4) + 127) / 128, 128, 0, st>>>(pu, vp_dev, p->unit_queue.p);
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[1], st));
        auto grid_for = [&](int n_units, int minb) {
            return std::max(1, std::min(p->sms * minb, (n_units + UNIT_WARPS - 1) / UNIT_WARPS));
        };
        if (p->n_units_bg > 0) {
            if (p->small_plan)
                unit_bg_kernel<3><<<grid_for(p->n_units_bg, 3), UNIT_THREADS, unit_bg_smem_bytes(), st>>>(
                    pu, p->unitmap_bg.p, p->n_units_bg, p->unit_queue.p, vp_dev);
            else
                unit_bg_kernel<CELESTE_UNIT_BG_MINB><<<grid_for(p->n_units_bg, CELESTE_UNIT_BG_MINB), UNIT_THREADS, unit_bg_smem_bytes(), st>>>(
                    pu, p->unitmap_bg.p, p->n_units_bg, p->unit_queue.p, vp_dev);
        }
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev_unit[0], st));
        if (p->n_units > 0)
            unit_walk_kernel<MODE><<<grid_for(p->n_units, MODE == 2 ? CELESTE_UNIT_MINB : CELESTE_UNIT_MINB_GRAD), UNIT_THREADS,
                                     unit_smem_bytes<MODE>(), st>>>(
                pu, p->unitmap.p, p->n_units, p->unit_queue.p + 1, vp_dev);
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev_unit[1], st));
        p->unit_timed = p->timing;
        if (p->n_units > 0) {
            if (MODE == 2)
                unit_moment_kernel<<<grid_for(p->n_units, CELESTE_UNIT_MOM_MINB), UNIT_THREADS, unit_moment_smem_bytes(), st>>>(
                    pu, p->unitmap.p, p->n_units, p->unit_queue.p + 2, vp_dev);
        }
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[2], st));
        if (MODE == 2 && !p->block_epilogue)
            epilogue_hess_kernel<<<(p->n_tasks + EPH_WARPS - 1) / EPH_WARPS, 32 * EPH_WARPS, 0, st>>>(
                pu, vp_dev, v, d, h, counters, flags, p->hess_layout == CELESTE_HESS_PACKED28 ? 1 : 0);
        else
            epilogue_kernel<MODE><<<p->n_tasks, EPI_THREADS, 0, st>>>(pu, vp_dev, v, d, h, counters, flags,
                                                                      p->hess_layout == CELESTE_HESS_PACKED28 ? 1 : 0);
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[3], st));
        CUDA_TRY(cudaGetLastError());
        return CELESTE_OK;
    }
    p->unit_timed = false;
    if constexpr (MODE <= 1) {
        if (p->use_march) {
            // value / gradient, production shape: row walks with the exp recurrence (march_kernels.cuh).  The blocks
            // build their sources' mixtures themselves: no set-up launch.
            if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[1], st));
            if (p->n_marchblocks > 0)
                march_kernel<MODE><<<p->n_marchblocks, MARCH_THREADS, march_smem_bytes(), st>>>(pd, p->marchmap.p, vp_dev);
            if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[2], st));
            march_epilogue_kernel<MODE><<<p->n_tasks, MEPI_THREADS, 0, st>>>(pd, vp_dev, p->march_part_ptr.p, v, d, counters, flags);
            if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[3], st));
            CUDA_TRY(cudaGetLastError());
            return CELESTE_OK;
        }
    }
    setup_kernel<<<sblocks, 128, 0, st>>>(pd, vp_dev);
    if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[1], st));
    if constexpr (MODE <= 1) {
        // value / gradient: task-level blocks, one partial per (sub, image, warp)
        PlanDev pt = pd;
        pt.chunk_ptr = p->task_chunk_ptr.p;
        const bool multi = p->n_subs > p->n_tasks;
        const size_t sm = task_smem_bytes<MODE>();
        if (p->n_taskblocks > 0) {
            if (p->uniform_K == 2 && !multi)
                task_kernel<MODE, 2, false><<<p->n_taskblocks, PIX_THREADS, sm, st>>>(pt, p->taskmap.p);
            else if (p->uniform_K == 2)
                task_kernel<MODE, 2, true><<<p->n_taskblocks, PIX_THREADS, sm, st>>>(pt, p->taskmap.p);
            else if (!multi)
                task_kernel<MODE, 0, false><<<p->n_taskblocks, PIX_THREADS, sm, st>>>(pt, p->taskmap.p);
            else
                task_kernel<MODE, 0, true><<<p->n_taskblocks, PIX_THREADS, sm, st>>>(pt, p->taskmap.p);
        }
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[2], st));
        epilogue_kernel<MODE><<<p->n_tasks, EPI_THREADS, 0, st>>>(pt, vp_dev, v, d, h, counters, flags);
        if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[3], st));
        CUDA_TRY(cudaGetLastError());
        return CELESTE_OK;
    } else if (p->n_blocks > 0) {
        const bool multi = p->n_subs > p->n_tasks;
        const size_t sm = pixel_smem_bytes<MODE>();
        if (p->uniform_K == 2 && !multi)
            pixel_kernel<MODE, 2, false><<<p->n_blocks, PIX_THREADS, sm, st>>>(pd, p->chunk_pixels);
        else if (p->uniform_K == 2)
            pixel_kernel<MODE, 2, true><<<p->n_blocks, PIX_THREADS, sm, st>>>(pd, p->chunk_pixels);
        else if (!multi)
            pixel_kernel<MODE, 0, false><<<p->n_blocks, PIX_THREADS, sm, st>>>(pd, p->chunk_pixels);
        else
            pixel_kernel<MODE, 0, true><<<p->n_blocks, PIX_THREADS, sm, st>>>(pd, p->chunk_pixels);
    }
    if (MODE == 2 && p->n_pairs > 0) {     // Sa > 1 (unit tests): cross-source Hessian blocks
        const size_t psm = ((size_t)NPAIR_ACC * PAIR_THREADS + 2 * (size_t)MAX_COMPS * COMP_STRIDE) * sizeof(double);
        if (p->uniform_K == 2)
            pair_kernel<2><<<p->n_pairs * p->N, PAIR_THREADS, psm, st>>>(pd);
        else
            pair_kernel<0><<<p->n_pairs * p->N, PAIR_THREADS, psm, st>>>(pd);
    }
    if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[2], st));
    epilogue_kernel<MODE><<<p->n_tasks, EPI_THREADS, 0, st>>>(pd, vp_dev, v, d, h, counters, flags,
                                                              p->hess_layout == CELESTE_HESS_PACKED28 ? 1 : 0);
    if (p->timing) CUDA_TRY(cudaEventRecord(p->ev[3], st));
    CUDA_TRY(cudaGetLastError());
    return CELESTE_OK;
}

extern "C" {

int celeste_elbo_plan_device(celeste_plan* p, const double* vp_dev, int32_t mode, double* v_dev, double* d_dev,
                             double* h_dev, int64_t* counters_dev, int32_t* flags_dev, void* cuda_stream) {
    if (!p || !vp_dev || !v_dev || !counters_dev || !flags_dev || mode < 0 || mode > 2 || (mode >= 1 && !d_dev) ||
        (mode >= 2 && !h_dev)) {
        set_detail("elbo_plan_device: bad arguments (mode=%d)", mode);
        return CELESTE_ERR_BAD_ARG;
    }
    for (size_t i = 0; i < p->fields.size(); ++i)
        if (p->patch_generation[i] != p->fields[i]->patch_generation) {
            set_detail("plan is stale: celeste_patches_set was called after celeste_plan_create");
            return CELESTE_ERR_STATE;
        }
    if (p->n_tasks == 0) return CELESTE_OK;
    {
        // the plan's buffers live on p->device: launching from another current device would hand the kernels
        // foreign pointers (an opaque CUDA error later); refuse it here
        int cur = -1;
        CUDA_TRY(cudaGetDevice(&cur));
        if (cur != p->device) {
            set_detail("elbo_plan_device: current device %d, the plan lives on device %d (call cudaSetDevice first)", cur, p->device);
            return CELESTE_ERR_STATE;
        }
    }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    long long* c = reinterpret_cast<long long*>(counters_dev);
    switch (mode) {
        case 0: return launch_mode<0>(p, vp_dev, v_dev, d_dev, h_dev, c, flags_dev, st);
        case 1: return launch_mode<1>(p, vp_dev, v_dev, d_dev, h_dev, c, flags_dev, st);
        default: return launch_mode<2>(p, vp_dev, v_dev, d_dev, h_dev, c, flags_dev, st);
    }
}

// Small plans: one H2D, one graph launch (the kernel sequence of `mode`, captured on first use), one D2H, one sync.
static int plan_host_small(celeste_plan* p, const double* vp, int32_t mode, double* v, double* d, double* h,
                           int64_t* counters, int32_t* flags, size_t nd, size_t nh) {
    const size_t nt = p->n_tasks, nvp = (size_t)p->n_slots * NPARAM;
    // block layout (8-byte aligned pieces): v | d | h | counters | flags
    const size_t o_v = 0, o_d = o_v + nt * 8, o_h = o_d + nd * 8, o_c = o_h + nh * 8, o_f = o_c + 2 * nt * 8;
    const size_t total = o_f + ((nt * 4 + 7) / 8) * 8;
    if (p->out_block.n < total) {
        p->drop_graphs();
        CUDA_TRY(p->out_block.alloc(total));
        if (p->out_pin) cudaFreeHost(p->out_pin);
        p->out_pin = nullptr;
        CUDA_TRY(cudaMallocHost((void**)&p->out_pin, total));
    }
    if (!p->vp_pin) CUDA_TRY(cudaMallocHost((void**)&p->vp_pin, nvp * sizeof(double)));
    CUDA_TRY(p->vp_dev.ensure(nvp));
    cudaStream_t st = p->stream;
    unsigned char* ob = p->out_block.p;
    std::memcpy(p->vp_pin, vp, nvp * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(p->vp_dev.p, p->vp_pin, nvp * sizeof(double), cudaMemcpyHostToDevice, st));
    const bool graphable = !p->timing && !p->task_mask && !p->need_pack;
    if (graphable && p->graph[mode] && p->graph_layout[mode] == p->hess_layout) {
        CUDA_TRY(cudaGraphLaunch(p->graph[mode], st));
    } else {
        const bool capture = graphable;
        if (p->graph[mode]) {
            cudaGraphExecDestroy(p->graph[mode]);
            p->graph[mode] = nullptr;
        }
        if (capture) CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = celeste_elbo_plan_device(p, p->vp_dev.p, mode, (double*)(ob + o_v), (double*)(ob + o_d), (double*)(ob + o_h),
                                          (int64_t*)(ob + o_c), (int32_t*)(ob + o_f), st);
        if (capture) {
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &g);
            if (rc != CELESTE_OK || e != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                if (rc != CELESTE_OK) return rc;
                CUDA_TRY(e);
            }
            CUDA_TRY(cudaGraphInstantiate(&p->graph[mode], g, 0));
            cudaGraphDestroy(g);
            p->graph_layout[mode] = p->hess_layout;
            CUDA_TRY(cudaGraphLaunch(p->graph[mode], st));
        } else if (rc != CELESTE_OK) {
            return rc;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(p->out_pin, ob, total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned char* op = p->out_pin;
    std::memcpy(v, op + o_v, nt * 8);
    if (mode >= 1) std::memcpy(d, op + o_d, nd * 8);
    if (mode >= 2) std::memcpy(h, op + o_h, nh * 8);
    if (counters) std::memcpy(counters, op + o_c, 2 * nt * 8);
    const int* hf = reinterpret_cast<const int*>(op + o_f);
    bool bad = false;
    for (size_t t = 0; t < nt; ++t) {
        if (flags) flags[t] = hf[t];
        if (hf[t] & CELESTE_FLAG_NONFINITE) {
            if (!bad) set_detail("task %zu: non-finite ELBO (assert_all_finite, elbo_args.jl:145)", t);
            bad = true;
        }
    }
    return bad ? CELESTE_ERR_NONFINITE : CELESTE_OK;
}

int celeste_elbo_plan_host(celeste_plan* p, const double* vp, int32_t mode, double* v, double* d, double* h,
                           int64_t* counters, int32_t* flags) {
    if (!p || !vp || !v || mode < 0 || mode > 2 || (mode >= 1 && !d) || (mode >= 2 && !h)) {
        set_detail("elbo_plan_host: bad arguments (mode=%d)", mode);
        return CELESTE_ERR_BAD_ARG;
    }
    if (p->n_tasks == 0) return CELESTE_OK;
    DeviceGuard guard(p->device);           // run on the plan's device, give the caller's current device back
    if (!guard.ok) {
        set_detail("elbo_plan_host: cannot select device %d", p->device);
        return CELESTE_ERR_CUDA;
    }
    if (!p->stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    const size_t nt = p->n_tasks;
    const size_t nd = mode >= 1 ? (size_t)p->n_subs * NPARAM : 0;
    const size_t nh = mode >= 2 ? (p->hess_layout == CELESTE_HESS_PACKED28 ? (size_t)p->n_tasks * HESS_PACKED_LEN : p->h_total) : 0;
    if ((nt * 3 + nd + nh) * 8 + (size_t)p->n_slots * NPARAM * 8 <= (1u << 20))        // <= 1 MB moved per call
        return plan_host_small(p, vp, mode, v, d, h, counters, flags, nd, nh);
    // v | counters | flags live in ONE device block mirrored by a pinned host block: one D2H for the three small
    // outputs (every copy costs ~8 us of latency on the stream, which is what a 1250-source step has to spare), and no
    // pageable destination; the gradient and the Hessian go straight into the caller's buffers
    const size_t o_v = 0, o_c = o_v + nt * 8, o_f = o_c + 2 * nt * 8, small_total = o_f + ((nt * 4 + 7) / 8) * 8;
    if (p->small_block.n < small_total) {
        CUDA_TRY(p->small_block.alloc(small_total));
        if (p->small_pin) cudaFreeHost(p->small_pin);
        p->small_pin = nullptr;
        CUDA_TRY(cudaMallocHost((void**)&p->small_pin, small_total));
    }
    CUDA_TRY(p->vp_dev.ensure((size_t)p->n_slots * NPARAM));
    if (mode >= 1) CUDA_TRY(p->d_dev.ensure(nd));
    if (mode >= 2) CUDA_TRY(p->h_dev.ensure(nh));
    cudaStream_t st = p->stream;
    unsigned char* sb = p->small_block.p;
    CUDA_TRY(cudaMemcpyAsync(p->vp_dev.p, vp, (size_t)p->n_slots * NPARAM * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = celeste_elbo_plan_device(p, p->vp_dev.p, mode, (double*)(sb + o_v), p->d_dev.p, p->h_dev.p,
                                      (int64_t*)(sb + o_c), (int32_t*)(sb + o_f), st);
    if (rc != CELESTE_OK) return rc;
    if (mode >= 1) CUDA_TRY(cudaMemcpyAsync(d, p->d_dev.p, nd * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (mode >= 2) CUDA_TRY(cudaMemcpyAsync(h, p->h_dev.p, nh * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(p->small_pin, sb, small_total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::memcpy(v, p->small_pin + o_v, nt * 8);
    if (counters) std::memcpy(counters, p->small_pin + o_c, 2 * nt * 8);
    const int* hflags = reinterpret_cast<const int*>(p->small_pin + o_f);
    bool bad = false;
    for (size_t t = 0; t < nt; ++t) {
        if (flags) flags[t] = hflags[t];
        if (hflags[t] & CELESTE_FLAG_NONFINITE) {
            if (!bad) set_detail("task %zu: non-finite ELBO (assert_all_finite, elbo_args.jl:145)", t);
            bad = true;
        }
    }
    return bad ? CELESTE_ERR_NONFINITE : CELESTE_OK;
}

int celeste_elbo_batch(celeste_field* f, int32_t n_tasks, const int32_t* task_ptr, const int32_t* source_ids,
                       const int32_t* active_ptr, const int32_t* active_idx, const double* vp, int32_t mode, double* v,
                       double* d, double* h, int64_t* counters, int32_t* flags) {
    if (n_tasks == 0) return CELESTE_OK;
    if (!f || n_tasks < 0 || !task_ptr || !active_ptr) {
        set_detail("elbo_batch: bad arguments");
        return CELESTE_ERR_BAD_ARG;
    }
    // Small task lists keep their plan in the field (see celeste_field::CachedPlan): a repeated call with the same
    // structure costs one H2D, one graph launch and one D2H.  CELESTE_PLAN_CACHE=0 disables it (A/B knob).
    static const bool cache_on = !(std::getenv("CELESTE_PLAN_CACHE") && std::atoi(std::getenv("CELESTE_PLAN_CACHE")) == 0);
    constexpr int CACHE_MAX_SLOTS = 512, CACHE_CAPACITY = 256;
    const int n_slots = task_ptr[n_tasks], n_act = active_ptr[n_tasks];
    if (!cache_on || n_slots < 0 || n_slots > CACHE_MAX_SLOTS || n_act < 0 || !source_ids || !active_idx) {
        celeste_plan* pl = nullptr;
        int rc = celeste_plan_create(f, &pl, n_tasks, task_ptr, source_ids, active_ptr, active_idx);
        if (rc != CELESTE_OK) return rc;
        rc = celeste_elbo_plan_host(pl, vp, mode, v, d, h, counters, flags);
        celeste_plan_destroy(pl);
        return rc;
    }
    std::vector<int32_t> key;
    key.reserve(2 + 2 * (size_t)n_tasks + 2 + n_slots + n_act);
    key.push_back(n_tasks);
    {
        // the kernel-selection knobs are read when a plan is built: a plan built under other knobs is another plan
        unsigned hsh = 2166136261u;
        for (const char* name : {"CELESTE_GRAD_KERNEL", "CELESTE_HESS_KERNEL", "CELESTE_MARCH_SPLIT", "CELESTE_MARCH_SPLIT_PCT",
                                 "CELESTE_CHUNK_PIXELS", "CELESTE_UNIT_ROWS", "CELESTE_UNIT_PIXELS", "CELESTE_EPILOGUE"}) {
            const char* e = std::getenv(name);
            for (const char* c = e ? e : ""; *c; ++c) hsh = (hsh ^ (unsigned char)*c) * 16777619u;
            hsh = (hsh ^ 0xffu) * 16777619u;
        }
        key.push_back((int32_t)hsh);
    }
    key.insert(key.end(), task_ptr, task_ptr + n_tasks + 1);
    key.insert(key.end(), active_ptr, active_ptr + n_tasks + 1);
    key.insert(key.end(), source_ids, source_ids + n_slots);
    key.insert(key.end(), active_idx, active_idx + n_act);
    celeste_plan* pl = nullptr;
    {
        std::lock_guard<std::mutex> lk(f->cache_mu);
        for (auto& c : f->cache)
            if (!c.busy && c.key == key) {
                if (c.plan->patch_generation[0] != f->patch_generation) {      // celeste_patches_set since: rebuild
                    delete c.plan;
                    c.plan = nullptr;
                    c.key.clear();
                    continue;
                }
                c.busy = true;
                c.last_use = ++f->cache_clock;
                pl = c.plan;
                break;
            }
    }
    bool cached = pl != nullptr;
    if (!pl) {
        int rc = celeste_plan_create(f, &pl, n_tasks, task_ptr, source_ids, active_ptr, active_idx);
        if (rc != CELESTE_OK) return rc;
    }
    int rc = celeste_elbo_plan_host(pl, vp, mode, v, d, h, counters, flags);
    {
        std::lock_guard<std::mutex> lk(f->cache_mu);
        if (cached) {
            for (auto& c : f->cache)
                if (c.plan == pl) c.busy = false;
        } else {
            // insert: reuse an emptied entry, else grow, else evict the least recently used idle entry
            celeste_field::CachedPlan* slot = nullptr;
            for (auto& c : f->cache)
                if (!c.plan) slot = &c;
            if (!slot && (int)f->cache.size() < CACHE_CAPACITY) {
                f->cache.emplace_back();
                slot = &f->cache.back();
            }
            if (!slot) {
                for (auto& c : f->cache)
                    if (!c.busy && (!slot || c.last_use < slot->last_use)) slot = &c;
                if (slot) delete slot->plan;
            }
            if (slot) {
                slot->key = std::move(key);
                slot->plan = pl;
                slot->busy = false;
                slot->last_use = ++f->cache_clock;
            } else {
                delete pl;       // every entry busy: do not keep this one
            }
        }
    }
    return rc;
}

int celeste_elbo_single(celeste_field* f, int32_t S, const int32_t* source_ids, int32_t Sa, const int32_t* active_idx,
                        const double* vp, int32_t mode, double* v, double* d, double* h, int64_t* counters,
                        int32_t* flags) {
    const int32_t task_ptr[2] = {0, S};
    const int32_t active_ptr[2] = {0, Sa};
    return celeste_elbo_batch(f, 1, task_ptr, source_ids, active_ptr, active_idx, vp, mode, v, d, h, counters, flags);
}

static int render_impl(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp, double* const* out,
                       int full_box);

int celeste_render_expectation(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp,
                               double* const* out) {
    return render_impl(f, S, source_ids, vp, out, 0);
}

int celeste_render_boxes(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp, double* const* out) {
    return render_impl(f, S, source_ids, vp, out, 1);
}

static int render_impl(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp, double* const* out,
                       int full_box) {
    if (!f || S < 0 || !out || (S > 0 && (!source_ids || !vp))) {
        set_detail("render_expectation: bad arguments (S=%d)", S);
        return CELESTE_ERR_BAD_ARG;
    }
    const int N = f->N;
    for (int n = 0; n < N; ++n)
        if (!out[n]) {
            set_detail("render_expectation: out[%d] is null", n);
            return CELESTE_ERR_BAD_ARG;
        }
    if (S == 0) {
        for (int n = 0; n < N; ++n)
            std::memset(out[n], 0, (size_t)f->h_images[n].H * f->h_images[n].W * sizeof(double));
        return CELESTE_OK;
    }
    // the S sources as the slots of a one-task plan (reuses setup_kernel: load_bvn_mixtures! + load_source_brightnesses)
    const int32_t task_ptr[2] = {0, S}, active_ptr[2] = {0, 1}, act[1] = {1};
    celeste_plan* raw = nullptr;
    int rc = celeste_plan_create(f, &raw, 1, task_ptr, source_ids, active_ptr, act);
    if (rc != CELESTE_OK) return rc;
    std::unique_ptr<celeste_plan, void (*)(celeste_plan*)> pl(raw, celeste_plan_destroy);
    CUDA_TRY(cudaSetDevice(f->device));
    if (!pl->stream) CUDA_TRY(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    cudaStream_t st = pl->stream;
    CUDA_TRY(pl->vp_dev.ensure((size_t)S * NPARAM));
    CUDA_TRY(cudaMemcpyAsync(pl->vp_dev.p, vp, (size_t)S * NPARAM * sizeof(double), cudaMemcpyHostToDevice, st));
    const PlanDev pd = plan_dev(pl.get());
    {
        const long total = (long)pl->n_slots * N * MAX_K;
        const int sblocks = (int)std::max<long>(1, std::min<long>((total + 127) / 128, 16L * sm_count()));
        setup_kernel<<<sblocks, 128, 0, st>>>(pd, pl->vp_dev.p);
    }
    std::vector<int> imgH(N), imgW(N);
    for (int n = 0; n < N; ++n) {
        imgH[n] = f->h_images[n].H;
        imgW[n] = f->h_images[n].W;
    }
    std::vector<RenderTile> tiles;
    std::vector<int> tile_slots;
    build_render_tiles(N, imgH.data(), imgW.data(), S,
                       [&](int s, int n, int& oh, int& ow, int& H2, int& W2) {
                           const PatchDev& p = f->h_patches[(size_t)(source_ids[s] - 1) + (size_t)n * f->S_tot];
                           oh = p.off_h;
                           ow = p.off_w;
                           H2 = p.H2;
                           W2 = p.W2;
                       },
                       tiles, tile_slots, full_box);
    DevBuf<RenderTile> d_tiles;
    DevBuf<int> d_slots;
    CUDA_TRY(d_tiles.upload(tiles));
    CUDA_TRY(d_slots.upload(tile_slots));
    std::vector<DevBuf<double>> d_out(N);
    std::vector<double*> h_ptrs(N);
    for (int n = 0; n < N; ++n) {
        const size_t cnt = (size_t)imgH[n] * imgW[n];
        CUDA_TRY(d_out[n].alloc(cnt));
        CUDA_TRY(cudaMemsetAsync(d_out[n].p, 0, cnt * sizeof(double), st));     // pixels no source covers: E_G - sky = 0
        h_ptrs[n] = d_out[n].p;
    }
    DevBuf<double*> d_ptrs;
    CUDA_TRY(d_ptrs.upload(h_ptrs));
    if (!tiles.empty()) {
        if (pl->uniform_K == 2)
            render_kernel<2><<<(unsigned)tiles.size(), RENDER_THREADS, 0, st>>>(pd, d_tiles.p, d_slots.p, d_ptrs.p, full_box);
        else
            render_kernel<0><<<(unsigned)tiles.size(), RENDER_THREADS, 0, st>>>(pd, d_tiles.p, d_slots.p, d_ptrs.p, full_box);
        CUDA_TRY(cudaGetLastError());
    }
    for (int n = 0; n < N; ++n)
        CUDA_TRY(cudaMemcpyAsync(out[n], d_out[n].p, (size_t)imgH[n] * imgW[n] * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CELESTE_OK;
}

int celeste_tr_subproblem(int32_t batch, int32_t n, const double* g_dev, const double* H_dev, const double* delta_dev,
                          const uint8_t* mask_dev, double* s_dev, double* m_dev, int32_t* interior_dev,
                          void* cuda_stream) {
    if (batch < 0 || n < 1 || n > TR_MAXN - 1 || !g_dev || !H_dev || !delta_dev || !s_dev || !m_dev || !interior_dev) {
        set_detail("tr_subproblem: bad arguments (batch=%d n=%d, n must be 1..%d)", batch, n, TR_MAXN - 1);
        return CELESTE_ERR_BAD_ARG;
    }
    if (batch == 0) return CELESTE_OK;
    int st0 = ensure_device_ready();
    if (st0 != CELESTE_OK) return st0;
    tr_subproblem_kernel<<<batch, TR_THREADS, 0, (cudaStream_t)cuda_stream>>>(n, g_dev, H_dev, delta_dev, mask_dev, s_dev,
                                                                              m_dev, interior_dev);
    CUDA_TRY(cudaGetLastError());
    return CELESTE_OK;
}

int celeste_newton_step(int32_t phase, int32_t batch, const celeste_newton_buffers* nb, void* cuda_stream) {
    if (phase < 0 || phase > 2 || batch < 0 || !nb) {
        set_detail("newton_step: bad arguments (phase=%d batch=%d)", phase, batch);
        return CELESTE_ERR_BAD_ARG;
    }
    if (!nb->x || !nb->lo || !nb->hi || !nb->vp_all || !nb->aslot ||
        (phase < 2 && (!nb->f || !nb->g || !nb->H || !nb->delta || !nb->x_new || !nb->m_pred || !nb->interior ||
                       !nb->active || !nb->converged || !nb->iters || !nb->f_calls || !nb->v || !nb->d || !nb->h ||
                       !nb->flags))) {
        set_detail("newton_step: null buffer");
        return CELESTE_ERR_BAD_ARG;
    }
    if (batch == 0) return CELESTE_OK;
    int st0 = ensure_device_ready();
    if (st0 != CELESTE_OK) return st0;
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
    newton_step_kernel<<<batch, TR_THREADS, 0, (cudaStream_t)cuda_stream>>>(dev, phase);
    CUDA_TRY(cudaGetLastError());
    return CELESTE_OK;
}

int celeste_fp64_peak(double* tflops_out, void* cuda_stream) {
    if (!tflops_out) return CELESTE_ERR_BAD_ARG;
    int st0 = ensure_device_ready();
    if (st0 != CELESTE_OK) return st0;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevBuf<double> out;
    CUDA_TRY(out.alloc(1));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sm_count() * 8, threads = 256, iters = 4096;
    dfma_peak_kernel<<<blocks, threads, 0, st>>>(out.p, 64, 1.0);   // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        dfma_peak_kernel<<<blocks, threads, 0, st>>>(out.p, iters, 1.0);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CUDA_TRY(cudaGetLastError());
    *tflops_out = best;
    return CELESTE_OK;
}

}  // extern "C"
