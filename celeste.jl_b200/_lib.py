"""ctypes binding of libceleste_cuda.so (include/celeste_cuda.h).

This is the Python analogue of the `ccall` shim a Celeste.jl maintainer would add
(INTEGRATION.md shows the Julia version).  There is NO fallback: if the CUDA
library is missing, cannot be loaded, or finds no device, every compute call
raises -- the product never routes through a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CELESTE_CUDA_LIB selects another build of the same library (kernel-tuning variants); never a CPU path
LIB_PATH = os.environ.get("CELESTE_CUDA_LIB") or os.path.join(_HERE, "libceleste_cuda.so")

CELESTE_OK = 0
CELESTE_ERR_NO_DEVICE = 1
CELESTE_ERR_BAD_ARG = 2
CELESTE_ERR_ALLOC = 3
CELESTE_ERR_CUDA = 4
CELESTE_ERR_UNSUPPORTED = 5
CELESTE_ERR_NONFINITE = 6
CELESTE_ERR_STATE = 7

MODE_VALUE, MODE_GRAD, MODE_HESS = 0, 1, 2
FLAG_NONFINITE = 1


class celeste_image(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("band", C.c_int32),
                ("pixels", C.c_void_p), ("sky", C.c_void_p),
                ("nelec_per_nmgy", C.c_void_p), ("log_iota", C.c_void_p)]


class celeste_patch(C.Structure):
    _fields_ = [("bitmap_offset", C.c_int64 * 2), ("H2", C.c_int32), ("W2", C.c_int32),
                ("active_pixel_bitmap", C.c_void_p),
                ("wcs_jacobian", C.c_double * 4), ("world_center", C.c_double * 2),
                ("pixel_center", C.c_double * 2), ("K", C.c_int32),
                ("psf", C.c_void_p), ("itp_coefs", C.c_void_p), ("itp_dims", C.c_int32 * 2)]


class celeste_patch_spec(C.Structure):
    _fields_ = [("bitmap_offset", C.c_int64 * 2), ("H2", C.c_int32), ("W2", C.c_int32),
                ("wcs_jacobian", C.c_double * 4), ("world_center", C.c_double * 2),
                ("pixel_center", C.c_double * 2), ("K", C.c_int32), ("grid_n", C.c_int32),
                ("psf", C.c_void_p), ("grid_psf", C.c_void_p)]


class CelesteError(RuntimeError):
    def __init__(self, status: int, msg: str, detail: str = ""):
        super().__init__(f"celeste_cuda status {status}: {msg}" + (f" [{detail}]" if detail else ""))
        self.status = status


class NonFiniteError(CelesteError):
    """assert_all_finite (elbo_args.jl:145-149) would have thrown in the reference."""


# every symbol include/celeste_cuda.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "celeste_get_errmsg", "celeste_get_errdetail", "celeste_version", "celeste_init",
    "celeste_field_create", "celeste_patches_set", "celeste_elbo_batch", "celeste_elbo_single",
    "celeste_plan_create", "celeste_plan_destroy", "celeste_plan_launches",
    "celeste_elbo_plan_device", "celeste_elbo_plan_host", "celeste_field_destroy",
    "celeste_fp64_peak", "celeste_plan_enable_timing", "celeste_plan_kernel_times", "celeste_set_chunk_pixels",
    "celeste_plan_create_multi", "celeste_tr_subproblem", "celeste_plan_set_task_mask", "celeste_newton_step",
    "celeste_render_expectation", "celeste_patches_build", "celeste_patch_readback", "celeste_find_neighbors",
    "celeste_plan_kernel_name", "celeste_plan_unit_times", "celeste_plan_set_hessian_layout", "celeste_render_boxes",
]


class celeste_newton_buffers(C.Structure):
    """include/celeste_cuda.h: device pointers of one batched Newton trust-region solve."""
    FIELDS = ["x", "f", "g", "H", "delta", "x_new", "m_pred", "interior", "active", "converged", "iters", "f_calls",
              "lo", "hi", "v", "d", "h", "flags", "vp_all", "aslot", "prior"]
    _fields_ = [(name, C.c_void_p) for name in FIELDS] + [("h_layout", C.c_int64)]


_lib = None


def load():
    """Load the CUDA library, or raise.  (No CPU fallback exists.)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    i32 = C.c_int32
    lib.celeste_get_errmsg.argtypes = [C.c_int, C.c_char_p]
    lib.celeste_get_errmsg.restype = None
    lib.celeste_get_errdetail.argtypes = [C.c_char_p]
    lib.celeste_get_errdetail.restype = None
    lib.celeste_version.restype = C.c_int
    lib.celeste_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.celeste_field_create.argtypes = [C.POINTER(vp), i32, C.POINTER(celeste_image)]
    lib.celeste_patches_set.argtypes = [vp, i32, i32, C.POINTER(celeste_patch)]
    lib.celeste_elbo_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.celeste_elbo_single.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.celeste_plan_create.argtypes = [vp, C.POINTER(vp), i32, vp, vp, vp, vp]
    lib.celeste_plan_create_multi.argtypes = [i32, C.POINTER(vp), C.POINTER(vp), i32, vp, vp, vp, vp, vp]
    lib.celeste_plan_destroy.argtypes = [vp]
    lib.celeste_plan_destroy.restype = None
    lib.celeste_plan_launches.argtypes = [vp, i32]
    lib.celeste_plan_kernel_name.argtypes = [vp, i32, C.c_char_p]
    lib.celeste_elbo_plan_device.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.celeste_elbo_plan_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.celeste_field_destroy.argtypes = [vp]
    lib.celeste_field_destroy.restype = None
    lib.celeste_fp64_peak.argtypes = [C.POINTER(C.c_double), vp]
    lib.celeste_plan_enable_timing.argtypes = [vp, i32]
    lib.celeste_plan_kernel_times.argtypes = [vp, C.POINTER(C.c_float * 3)]
    lib.celeste_plan_unit_times.argtypes = [vp, C.POINTER(C.c_float * 3)]
    lib.celeste_plan_set_hessian_layout.argtypes = [vp, i32]
    lib.celeste_set_chunk_pixels.argtypes = [i32]
    lib.celeste_plan_set_task_mask.argtypes = [vp, vp]
    lib.celeste_tr_subproblem.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.celeste_render_expectation.argtypes = [vp, i32, vp, vp, vp]
    lib.celeste_render_boxes.argtypes = [vp, i32, vp, vp, vp]
    lib.celeste_patches_build.argtypes = [vp, i32, i32, C.POINTER(celeste_patch_spec)]
    lib.celeste_patch_readback.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.celeste_find_neighbors.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    lib.celeste_newton_step.argtypes = [i32, i32, C.POINTER(celeste_newton_buffers), vp]
    _lib = lib
    return lib


def errmsg(status: int) -> str:
    buf = C.create_string_buffer(64)
    load().celeste_get_errmsg(status, buf)
    return buf.value.decode()


def errdetail() -> str:
    buf = C.create_string_buffer(512)
    load().celeste_get_errdetail(buf)
    return buf.value.decode()


def check(status: int, allow_nonfinite: bool = False):
    if status == CELESTE_OK:
        return
    if status == CELESTE_ERR_NONFINITE:
        if allow_nonfinite:
            return
        raise NonFiniteError(status, errmsg(status), errdetail())
    raise CelesteError(status, errmsg(status), errdetail())
