/*
 * celeste_cuda.h -- C ABI of the B200-native (sm_100a) ELBO hot path of Celeste.jl.
 *
 * Drop-in boundary: this library replaces exactly one
 * reference method,
 *
 *   DeterministicVI.elbo_likelihood(ea::ElboArgs, vp::VariationalParams{Float64},
 *                                   elbo_vars, bvn_bundle)
 *       src/deterministic_vi/elbo_objective.jl:400-474
 *
 * (and its direct callee tree: load_source_brightnesses source_brightness.jl:213,
 * load_bvn_mixtures! fsm_util.jl:111, add_pixel_term! elbo_objective.jl:330,
 * star_light_density! fsm_util.jl:225, populate_gal_fsm! fsm_util.jl:194,
 * calculate_G_s! elbo_objective.jl:17, add_elbo_log_term! elbo_objective.jl:274,
 * and the SensitiveFloats algebra SensitiveFloats.jl:83-250).
 *
 * Conventions (they mirror the reference's own FFI, src/SEP.jl:39-46,137-150):
 *   - every entry point returns an int status, 0 == CELESTE_OK; a message for a
 *     status is obtained with celeste_get_errmsg (cf. sep_get_errmsg);
 *   - opaque handles are created/destroyed by the library; every data buffer is
 *     caller-owned and only has to stay alive for the duration of the call;
 *   - all matrices are column-major (Julia layout), "h fastest";
 *   - integer indices that name sources are 1-based, exactly the values the
 *     Julia side holds (ea.active_sources, rows of ea.patches);
 *   - no exception / longjmp ever crosses this boundary;
 *   - the library never silently falls back to a CPU path: without a usable
 *     CUDA device every compute entry point returns CELESTE_ERR_NO_DEVICE.
 *
 * Only T == Float64 is redirected here; ForwardDiff.Dual calls stay in Julia
 * (test/test_elbo.jl:232).
 */
#ifndef CELESTE_CUDA_H
#define CELESTE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- layout constants of the reference parameterisation ------------------ */
/* length(CanonicalParams), src/model/param_set.jl:107 */
#define CELESTE_NUM_PARAMS 44
/* NUM_BANDS (u g r i z) */
#define CELESTE_NUM_BANDS 5
/* 1-based canonical ids, src/model/param_set.jl:88-100 */
#define CELESTE_ID_POS 1           /* 1:2   */
#define CELESTE_ID_GAL_FRAC_DEV 3
#define CELESTE_ID_GAL_AXIS_RATIO 4
#define CELESTE_ID_GAL_ANGLE 5
#define CELESTE_ID_GAL_RADIUS_PX 6
#define CELESTE_ID_FLUX_LOC 7      /* 7:8   (star, galaxy) */
#define CELESTE_ID_FLUX_SCALE 9    /* 9:10  */
#define CELESTE_ID_COLOR_MEAN 11   /* 11:18 (4 x 2 col-major) */
#define CELESTE_ID_COLOR_VAR 19    /* 19:26 */
#define CELESTE_ID_IS_STAR 27      /* 27:28 */
#define CELESTE_ID_K 29            /* 29:44 (never receives likelihood derivatives) */

/* ---- status codes --------------------------------------------------------- */
#define CELESTE_OK 0
#define CELESTE_ERR_NO_DEVICE 1     /* no CUDA device / driver; nothing was computed */
#define CELESTE_ERR_BAD_ARG 2
#define CELESTE_ERR_ALLOC 3
#define CELESTE_ERR_CUDA 4          /* a CUDA runtime call failed; see errmsg detail */
#define CELESTE_ERR_UNSUPPORTED 5   /* e.g. Sa > 1 in this build: caller keeps the Julia path */
#define CELESTE_ERR_NONFINITE 6     /* assert_all_finite (elbo_args.jl:145) would have thrown */
#define CELESTE_ERR_STATE 7

/* evaluation mode == (elbo.has_gradient, elbo.has_hessian), elbo_objective.jl:69,95 */
#define CELESTE_MODE_VALUE 0
#define CELESTE_MODE_GRAD 1
#define CELESTE_MODE_HESS 2

/* per-task flag bits written to `flags` */
#define CELESTE_FLAG_NONFINITE 1

/*
 * One Model.Image (src/model/image_model.jl:6-38), flattened.
 * `sky` is the DENSE matrix img.sky[h, w] materialised by the host (for SDSS the
 * Float32 bilinear rule of src/SDSSIO.jl:70-98 must be evaluated by the caller
 * so that no Float32 rounding can differ).  `log_iota[h]` must be
 * Float64(log(nelec_per_nmgy[h]::Float32)) -- the reference evaluates that log
 * in Float32 (elbo_objective.jl:292); NULL lets the library do (double)logf().
 */
typedef struct celeste_image {
    int32_t H, W;                  /* image_model.jl:7-8 */
    int32_t band;                  /* img.b, 1..5 */
    const float*  pixels;          /* H x W, electrons, NaN == masked */
    const float*  sky;             /* H x W, nmgy */
    const float*  nelec_per_nmgy;  /* H (varies by row) */
    const double* log_iota;        /* H or NULL */
} celeste_image;

/*
 * One Model.ImagePatch (src/model/imaged_sources.jl:60-71), flattened.
 * psf      : K x 7 doubles per component k: alphaBar, xiBar[1:2], tauBar[1,1],
 *            tauBar[2,1], tauBar[1,2], tauBar[2,2]   (psf_model.jl:17-29)
 * itp_coefs: the prefiltered cubic B-spline coefficient array of patch.itp_psf
 *            (Interpolations.jl `itp.coefs`, padded by one on each side:
 *            (n1+2) x (n2+2), 53 x 53 for the 51 x 51 stamp), column-major.
 *            Patches that share one array may pass the same pointer; the
 *            library de-duplicates by (pointer, dims).
 */
typedef struct celeste_patch {
    int64_t bitmap_offset[2];            /* imaged_sources.jl:69 (box corner - 1) */
    int32_t H2, W2;                      /* size(active_pixel_bitmap) */
    const uint8_t* active_pixel_bitmap;  /* H2 x W2, 0/1 */
    double wcs_jacobian[4];              /* 2 x 2 col-major */
    double world_center[2];
    double pixel_center[2];
    int32_t K;                           /* psf components (ea.psf_K) */
    const double* psf;                   /* K x 7 */
    const double* itp_coefs;
    int32_t itp_dims[2];                 /* padded dims, e.g. {53, 53} */
} celeste_patch;

typedef struct celeste_field celeste_field;

/* Like sep_get_errmsg (src/SEP.jl:41): short message for a status code. buf >= 61 bytes. */
void celeste_get_errmsg(int status, char* buf);
/* Longer, thread-local detail of the most recent failure on the calling thread. buf >= 512. */
void celeste_get_errdetail(char* buf);

/* Library/ABI version: major*10000 + minor*100 + patch. */
int celeste_version(void);

/*
 * Select / probe the CUDA device this process evaluates on (one process per GPU:
 * ParallelRun shards sources across processes).  device < 0 keeps the
 * current device.  n_devices_out (may be NULL) receives cudaGetDeviceCount.
 */
int celeste_init(int device, int* n_devices_out);

/*
 * Upload the images of one inference box (ea.images, elbo_args.jl:178) once.
 * Pixel planes stay resident in HBM until celeste_field_destroy.
 */
int celeste_field_create(celeste_field** out, int32_t N, const celeste_image* imgs);

/*
 * Upload the S_tot x N patch matrix (`patches` of ParallelRun._infer_box,
 * ParallelRun.jl:610-637; column-major: patch of source s (1-based) in image n
 * is p[(s-1) + (n-1)*S_tot]).  May be called again to replace the patch set.
 */
int celeste_patches_set(celeste_field* f, int32_t S_tot, int32_t N, const celeste_patch* p);

/*
 * Evaluate elbo_likelihood for a batch of independent (ElboArgs, vp) tasks.
 *
 * Task t (0-based) owns the local sources  source_ids[task_ptr[t] .. task_ptr[t+1])
 * -- 1-based rows of the patch matrix, in the order of ea.patches' rows
 * (ParallelRun.jl:242,485: target first, then its neighbours) -- and the
 * variational parameters vp[44 * task_ptr[t] ...] (44 x S_t column-major, i.e.
 * vp[s] of the reference concatenated; re-read on every call, never cached,
 * SURVEY appendix B.10).  active_idx[active_ptr[t] .. active_ptr[t+1]) are
 * ea.active_sources (1-based LOCAL indices, 1..S_t).
 *
 * Outputs (caller-allocated, host memory), with Sa_t = active_ptr[t+1]-active_ptr[t]
 * and P = 44*Sa_t:
 *   v[t]                               elbo.v[]
 *   d[44*active_ptr[t] ...]            elbo.d, 44 x Sa_t col-major      (mode >= 1)
 *   h[h_ptr(t) ...]                    elbo.h, P x P col-major, symmetric (mode == 2),
 *                                      h_ptr(t) = sum_{u<t} (44*Sa_u)^2
 *   counters[2*t + {0,1}]              active / inactive pixel-visit counters
 *                                      (elbo_objective.jl:353-357)
 *   flags[t]                           CELESTE_FLAG_* bits
 * d / h may be NULL when the mode does not produce them.
 *
 * Returns CELESTE_ERR_NONFINITE if any task produced a non-finite value (all
 * outputs are still written; inspect flags[]).  1 <= Sa_t <= 8 (production uses 1,
 * ParallelRun.jl:253,489; the reference's unit tests use 2): each pixel is visited
 * once (`already_visited`, elbo_objective.jl:450-455), the Hessian carries the
 * cross-source blocks of combine_sfs_hessian! (SensitiveFloats.jl:114-126);
 * CELESTE_ERR_UNSUPPORTED for Sa_t > 8.
 */
int celeste_elbo_batch(celeste_field* f, int32_t n_tasks,
                       const int32_t* task_ptr, const int32_t* source_ids,
                       const int32_t* active_ptr, const int32_t* active_idx,
                       const double* vp, int32_t mode,
                       double* v, double* d, double* h,
                       int64_t* counters, int32_t* flags);

/*
 * Single-task, thread-safe convenience with the shape of the reference call
 * (one ElboArgs): used by the elbo_likelihood method override.
 */
int celeste_elbo_single(celeste_field* f, int32_t S, const int32_t* source_ids,
                        int32_t Sa, const int32_t* active_idx,
                        const double* vp, int32_t mode,
                        double* v, double* d, double* h,
                        int64_t* counters, int32_t* flags);

/*
 * Device-resident variant: every pointer argument is a DEVICE pointer on the
 * current device, work is enqueued on `cuda_stream` (a cudaStream_t, NULL == the
 * legacy default stream) and the call returns without synchronising.  The task
 * plan (task_ptr/source_ids/active_*) must have been registered with
 * celeste_plan_create.  Used to time the kernels with inputs resident in HBM.
 */
typedef struct celeste_plan celeste_plan;
int celeste_plan_create(celeste_field* f, celeste_plan** out, int32_t n_tasks,
                        const int32_t* task_ptr, const int32_t* source_ids,
                        const int32_t* active_ptr, const int32_t* active_idx);
/*
 * One plan over the tasks of SEVERAL inference boxes (fields with the same image count N, e.g. the
 * fields of a stripe handled by one process): task t belongs to fields[task_field[t]] (0-based) and
 * its source ids index that field's patch matrix.  All tasks are evaluated by the same three kernel
 * launches, so small boxes do not pay a launch tail each.
 */
int celeste_plan_create_multi(int32_t n_fields, celeste_field* const* fields, celeste_plan** out, int32_t n_tasks,
                              const int32_t* task_field, const int32_t* task_ptr, const int32_t* source_ids,
                              const int32_t* active_ptr, const int32_t* active_idx);
void celeste_plan_destroy(celeste_plan* p);
/* number of kernel launches one celeste_elbo_plan_device call enqueues */
int celeste_plan_launches(const celeste_plan* p, int32_t mode);
/*
 * Name of the kernel family that carries the pixel loop of `mode` for this plan, written to buf (<= 31 chars + NUL):
 * "unit_kernel" (every mode when every task has Sa = 1 and every patch K = 2 -- the production shape:
 * ParallelRun.jl:253,489, elbo_args.jl:197 -- unit_bg_kernel + unit_walk_kernel<mode> [+ unit_moment_kernel]),
 * "march_kernel" / "task_kernel" (value / gradient: CELESTE_GRAD_KERNEL=march / =task, or any other shape),
 * "pixel_kernel" (Hessian: CELESTE_HESS_KERNEL=pixel, or any other shape).  For profiling / bench reports.
 */
int celeste_plan_kernel_name(const celeste_plan* p, int32_t mode, char* buf);
int celeste_elbo_plan_device(celeste_plan* p, const double* vp_dev, int32_t mode,
                             double* v_dev, double* d_dev, double* h_dev,
                             int64_t* counters_dev, int32_t* flags_dev,
                             void* cuda_stream);
/* Same plan, HOST buffers, pinned staging + H2D/D2H inside the call (synchronous). */
int celeste_elbo_plan_host(celeste_plan* p, const double* vp, int32_t mode,
                           double* v, double* d, double* h,
                           int64_t* counters, int32_t* flags);

/*
 * Optional instrumentation: when enabled, celeste_elbo_plan_device / _host bracket each of their
 * kernels with CUDA events on the launch stream; celeste_plan_kernel_times synchronises on the last
 * evaluation and returns its device times in milliseconds: ms[0] setup, ms[1] pixel (the hot kernel),
 * ms[2] epilogue.  Used by bench.py for the roofline of the dominant kernel.
 */
int celeste_plan_enable_timing(celeste_plan* p, int32_t on);
int celeste_plan_kernel_times(celeste_plan* p, float ms[3]);
/* the "pixel" share of the last evaluation split by kernel when the unit kernels ran (else zeros):
 * ms[0] unit_bg_kernel (neighbours), ms[1] unit_walk_kernel (the dominant kernel), ms[2] unit_moment_kernel */
int celeste_plan_unit_times(celeste_plan* p, float ms[3]);
/*
 * Layout of the Hessian a plan writes (mode 2).  CELESTE_HESS_DENSE (default): per task the (44 Sa) x (44 Sa)
 * column-major matrix of the reference's SensitiveFloat (SensitiveFloats.jl:29-31).  CELESTE_HESS_PACKED28 (plans
 * whose tasks all have Sa = 1): per task the 406 doubles of the upper triangle (row-major: (0,0) (0,1) .. (0,27)
 * (1,1) ..) of the 28 x 28 block of canonical ids 1..28 -- everything else of the 44 x 44 matrix is exactly zero
 * (ids.k never receives a likelihood derivative) and the matrix is exactly symmetric, so nothing is lost; it is 4.8x
 * less device->host traffic.  h / h_dev of the evaluation calls then hold 406 doubles per task.
 */
#define CELESTE_HESS_DENSE 0
#define CELESTE_HESS_PACKED28 1
#define CELESTE_HESS_PACKED28_LEN 406
int celeste_plan_set_hessian_layout(celeste_plan* p, int32_t layout);
/*
 * Optional per-task mask for the plan's evaluations: mask_dev is a DEVICE array of n_tasks bytes that the caller
 * may rewrite between calls; tasks whose byte is 0 are skipped and their outputs left untouched (the batched
 * Newton driver stops evaluating sources that have converged).  NULL removes the mask.
 */
int celeste_plan_set_task_mask(celeste_plan* p, const uint8_t* mask_dev);
/* chunk size (pixels per pixel-kernel block) used by plans created afterwards; 0 restores the default */
int celeste_set_chunk_pixels(int32_t chunk_pixels);

void celeste_field_destroy(celeste_field* f);

/*
 * Row f.4, ImagePatch construction on the device: what the ImagePatch(img, box) constructor
 * (src/model/imaged_sources.jl:80-117) computes from pixels and from the PSF stamp,
 *     active_pixel_bitmap = !isnan(pixels in box)  (:92-95),
 *     itp_psf = cubic B-spline of softpluslike(normalise(max(psfmap(center), 0) + 1e-6))  (:97-107),
 * built here from the images already resident in the field.  The host keeps the box arithmetic and the WCS
 * linearisation (clamp_box :10-14, pixel_center / world_center / pixel_world_jacobian :86-90; WCS.jl is a host
 * library) and passes them per patch.  `grid_psf`: the raw psfmap stamp at the patch centre (grid_n x grid_n doubles,
 * column-major, HOST pointer; identical pointers are processed once) or NULL to rasterise the K-component mixture
 * `psf` (render_psf, src/model/psf_model.jl:61-75).  3 <= grid_n <= 54 (51 in the reference).
 * Same effect as celeste_patches_set with host-built bitmaps and coefficient arrays.
 */
typedef struct celeste_patch_spec {
    int64_t bitmap_offset[2];   /* first(box[d]) - 1 of the clamped box                       */
    int32_t H2, W2;             /* size of the clamped box (0 allowed)                        */
    double wcs_jacobian[4];     /* column-major 2 x 2                                         */
    double world_center[2];
    double pixel_center[2];
    int32_t K;
    int32_t grid_n;
    const double* psf;          /* K x 7: alphaBar, xiBar[2], tauBar[4] (column-major)        */
    const double* grid_psf;     /* grid_n x grid_n raw stamp, or NULL                         */
} celeste_patch_spec;
int celeste_patches_build(celeste_field* f, int32_t S_tot, int32_t N, const celeste_patch_spec* specs /* S_tot x N col-major */);

/* Inspection: copy patch (s, n) (0-based) back to the host.  dims_out = {H2, W2, n1, n2}; bitmap (H2*W2 bytes) and
 * coefs (n1*n2 doubles) may be NULL to query the sizes only. */
int celeste_patch_readback(celeste_field* f, int32_t s, int32_t n, int32_t* dims_out, uint8_t* bitmap, double* coefs);

/*
 * find_neighbors (src/model/imaged_sources.jl:232-244) for every source of the field's patch matrix at once:
 * nbr_ptr (S_tot + 1) and nbr (0-based source indices, ascending per target) in CSR form.  If `capacity` is too
 * small nothing is written to nbr, *needed_out receives the required length and CELESTE_ERR_BAD_ARG is returned
 * (call with capacity 0 first to size the buffer).
 */
int celeste_find_neighbors(celeste_field* f, int32_t* nbr_ptr, int32_t* nbr, int64_t capacity, int64_t* needed_out);

/*
 * Row f.4, the value-only full-image render: fill_celeste_expectation! (bin/write_celeste_expectation.jl:111-156),
 * which calls add_pixel_term! (elbo_objective.jl:330-392) in value mode on EVERY pixel of every image of the
 * field.  For image n (n = 0..N-1), out[n] (HOST, H x W doubles, column-major like the image) receives
 *     E_G(h, w) - sky(h, w) = sum over the S sources whose patch covers (h, w)  [same in-patch test as the ELBO,
 *                             incl. the strict w2 < W2 of :349]  of  a_star E_l_star f_star + a_gal E_l_gal f_gal
 * in nanomaggies, sources added in the order given (the reference then does image.pixels[h, w] += that).
 * source_ids: 1-based rows of the patch matrix; vp: 44 x S column-major.  Synchronous.
 */
int celeste_render_expectation(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp,
                               double* const* out);
/*
 * The same render over EVERY column of each source's box (no strict w2 < W2): what Synthetic.gen_image!
 * (src/Synthetic.jl:30-47) adds for one body -- its expected flux on the whole radius-25 box -- when vp holds the
 * catalog's fluxes exactly (is_star 0 / 1, flux_scale = color_var = 0: E_l = exp(flux_loc + colours) = catalog flux).
 * The caller adds the sky, multiplies by iota and draws the Poisson sample (synthetic.gen_images_device).
 */
int celeste_render_boxes(celeste_field* f, int32_t S, const int32_t* source_ids, const double* vp, double* const* out);

/*
 * Row f.2 (batched Newton trust region; the step ElboMaximize.maximize! delegates to Optim.NewtonTrustRegion,
 * src/deterministic_vi/ElboMaximize.jl:105-108,235): for each of `batch` sources solve
 *     min_s  g's + 1/2 s'Hs   subject to |s| <= delta
 * exactly (Jacobi eigen-decomposition + secular equation, hard case included).  All pointers are DEVICE
 * pointers; g: batch x n, H: batch x n x n (symmetric), delta: batch; outputs s: batch x n, m: batch
 * (predicted change of the objective), interior: batch (1 when the unconstrained Newton step was taken).
 * mask_dev (nullable): batch bytes; sources with a zero byte are skipped and their outputs left untouched.
 * n <= 47 (41 free parameters in Celeste).  Asynchronous on `cuda_stream`.
 */
int celeste_tr_subproblem(int32_t batch, int32_t n, const double* g_dev, const double* H_dev, const double* delta_dev,
                          const uint8_t* mask_dev, double* s_dev, double* m_dev, int32_t* interior_dev,
                          void* cuda_stream);

/*
 * Rows f.1 + f.2 + f.3 fused: one lock-step iteration of ElboMaximize.maximize! (src/deterministic_vi/
 * ElboMaximize.jl:228-242; evaluate! :161-172; Optim.NewtonTrustRegion options :95-108) for `batch` sources at
 * once, everything that is not the likelihood itself:
 *   -KL (elbo_kl.jl:143-154) added to the plan's 44-space value / gradient / Hessian, propagate_derivatives!
 *   (ConstraintTransforms.jl:373-396) to the 41 free coordinates, the trust-region accept / radius / convergence
 *   update, and -- for sources still iterating -- the next subproblem solve, the candidate x + s and to_bound!
 *   (:189-196) of the candidate written into vp_all, ready for the next plan evaluation.
 * All pointers are DEVICE pointers owned by the caller (sizes in units of `batch` = B):
 */
typedef struct celeste_newton_buffers {
    double* x;            /* B x 41  accepted iterate (free coordinates)                           in/out */
    double* f;            /* B       objective -ELBO at x                                          in/out */
    double* g;            /* B x 41  its gradient                                                  in/out */
    double* H;            /* B x 41 x 41  its Hessian                                              in/out */
    double* delta;        /* B       trust-region radius                                           in/out */
    double* x_new;        /* B x 41  candidate under evaluation                                    in/out */
    double* m_pred;       /* B       change the subproblem predicted for the candidate             in/out */
    int32_t* interior;    /* B       1 when that candidate was the unconstrained Newton step       in/out */
    uint8_t* active;      /* B       1 while the source iterates; usable as celeste_plan_set_task_mask    */
    uint8_t* converged;   /* B       1 once an x / f / g tolerance was met                         out    */
    int32_t* iters;       /* B       iterations taken                                              out    */
    int32_t* f_calls;     /* B       ELBO evaluations consumed                                     out    */
    const double* lo;     /* B x 26  box lower bounds (ElboMaximize.jl:70-85)                             */
    const double* hi;     /* B x 26  box upper bounds                                                     */
    const double* v;      /* B       plan outputs at the candidate (celeste_elbo_plan_device, mode 2)     */
    const double* d;      /* B x 44                                                                       */
    const double* h;      /* B x 44 x 44, or B x 406 when h_layout = CELESTE_HESS_PACKED28                */
    const int32_t* flags; /* B                                                                            */
    double* vp_all;       /* n_slots x 44  the plan's bound parameters; row aslot[b] belongs to source b  */
    const int64_t* aslot; /* B                                                                            */
    const double* prior;  /* 360 doubles (kl.KLTerm.packed layout, see maximize_kernels.cuh) or NULL: no KL */
    int64_t h_layout;     /* CELESTE_HESS_DENSE or CELESTE_HESS_PACKED28: layout of `h`                   */
} celeste_newton_buffers;
/*
 * phase 0: the plan was evaluated at to_bound(x): initialise f, g, H, delta = 1, active, and emit the first
 * candidate.  phase 1: the plan was evaluated at the candidate: accept / reject, update, emit the next
 * candidate for sources still active.  phase 2: write to_bound(x) of every source into vp_all (maximize! :239).
 * Asynchronous on `cuda_stream`.
 */
int celeste_newton_step(int32_t phase, int32_t batch, const celeste_newton_buffers* buffers, void* cuda_stream);

/*
 * Measured FP64 FMA peak of the current device (a register-resident DFMA chain,
 * 2 flop per FMA), in TFLOP/s: the denominator of the roofline of this
 * FP64-pipe-bound path.  Runs on `cuda_stream`, synchronises.
 */
int celeste_fp64_peak(double* tflops_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* CELESTE_CUDA_H */
