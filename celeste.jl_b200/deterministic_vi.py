"""`DeterministicVI` surface of the reference, backed by the CUDA library.

Mirrors, for the hot path only:
  src/DeterministicVI.jl:39-105           generic_init_source, catalog_init_source, init_sources
  src/deterministic_vi/elbo_args.jl:165-211   ElboArgs
  src/SensitiveFloats.jl:23-47            SensitiveFloat (result container)
  src/deterministic_vi/elbo_objective.jl:400-492  elbo_likelihood / elbo
`elbo_likelihood(ea, vp)` has the reference's meaning and error behaviour (raises on a
non-finite result like assert_all_finite, elbo_args.jl:145); the arithmetic runs in
libceleste_cuda.so on the current CUDA device.  `DeviceField` is the batched entry the
reference lacks (many (ElboArgs, vp) per launch; SURVEY.md 8b "Threading").
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .flatten import FlatImages, FlatPatches, csr_tasks, out_sizes
from .model import (CatalogEntry, Image, ImagePatch, NUM_PARAMS, ids)

VariationalParams = List[np.ndarray]   # Vector{Vector{Float64}}: one 44-vector per source


# ------------------------------------------------------------------ DeterministicVI.jl
def generic_init_source(init_pos) -> np.ndarray:
    """DeterministicVI.jl:39-53."""
    ret = np.empty(NUM_PARAMS)
    ret[ids.is_star] = 0.5
    ret[ids.pos] = init_pos
    ret[ids.flux_loc] = math.log(2.0)
    ret[ids.flux_scale] = 1e-3
    ret[ids.gal_frac_dev] = 0.5
    ret[ids.gal_axis_ratio] = 0.5
    ret[ids.gal_angle] = 0.0
    ret[ids.gal_radius_px] = 1.0
    ret[ids.k] = 1.0 / ids.k.shape[0]
    ret[ids.color_mean] = 0.0
    ret[ids.color_var] = 1e-2
    return ret


def catalog_init_source(ce: CatalogEntry, max_gal_radius_px=float("inf")) -> np.ndarray:
    """DeterministicVI.jl:59-91."""
    ret = generic_init_source(ce.pos)
    ret[ids.is_star[0]] = 0.8 if ce.is_star else 0.2
    ret[ids.is_star[1]] = 0.2 if ce.is_star else 0.8
    ret[ids.flux_loc[0]] = math.log(max(0.1, ce.star_fluxes[2]))
    ret[ids.flux_loc[1]] = math.log(max(0.1, ce.gal_fluxes[2]))

    def get_color(c_var, c_mean):
        if c_var > 0 and c_mean > 0:
            return min(max(math.log(c_var / c_mean), -9.0), 9.0)
        if c_var > 0 and c_mean <= 0:
            return 3.0
        if c_var <= 0 and c_mean > 0:
            return -3.0
        return 0.0

    def get_colors(raw):
        return [get_color(raw[c + 1], raw[c]) for c in range(4)]

    ret[ids.color_mean[:, 0]] = get_colors(ce.star_fluxes)
    ret[ids.color_mean[:, 1]] = get_colors(ce.gal_fluxes)
    ret[ids.gal_frac_dev] = min(max(ce.gal_frac_dev, 0.015), 0.985)
    ret[ids.gal_axis_ratio] = 0.8 if ce.is_star else min(max(ce.gal_axis_ratio, 0.015), 0.985)
    ret[ids.gal_angle] = ce.gal_angle
    ret[ids.gal_radius_px] = 0.2 if ce.is_star else min(max_gal_radius_px, max(ce.gal_radius_px, 0.2))
    return ret


def init_sources(target_sources: Sequence[int], catalog: Sequence[CatalogEntry]) -> VariationalParams:
    """DeterministicVI.jl:94-103 (target_sources 0-based here)."""
    ret = [catalog_init_source(ce) for ce in catalog]
    for s in target_sources:
        ret[s][:] = generic_init_source(catalog[s].pos)
    return ret


# ------------------------------------------------------------------ SensitiveFloats.jl
class SensitiveFloat:
    """SensitiveFloats.jl:23-47: value, d (local_P x local_S), h ((P S) x (P S), p fastest)."""

    def __init__(self, local_P: int, local_S: int, has_gradient=True, has_hessian=True):
        assert has_gradient or not has_hessian
        self.local_P, self.local_S = local_P, local_S
        self.has_gradient, self.has_hessian = has_gradient, has_hessian
        self.v = 0.0
        self.d = np.zeros((local_P * has_gradient, local_S * has_gradient), order="F")
        hd = local_P * local_S * has_hessian
        self.h = np.zeros((hd, hd), order="F")

    @property
    def mode(self) -> int:
        return _lib.MODE_HESS if self.has_hessian else (_lib.MODE_GRAD if self.has_gradient else _lib.MODE_VALUE)


class ElboIntermediateVariables:
    """elbo_args.jl:29-113: only the parts a caller observes -- the result `elbo` (whose
    has_gradient/has_hessian flags select the mode, elbo_objective.jl:69,95) and the two
    pixel-visit counters (elbo_args.jl:62-63).  The scratch itself lives on the device."""

    def __init__(self, num_active_sources: int, calculate_gradient=True, calculate_hessian=True):
        self.elbo = SensitiveFloat(NUM_PARAMS, num_active_sources, calculate_gradient,
                                   calculate_gradient and calculate_hessian)
        self.active_pixel_counter = 0
        self.inactive_pixel_counter = 0


# ------------------------------------------------------------------ device-resident box
class DeviceField:
    """Images + patch matrix of one inference box, resident in HBM (celeste_field)."""

    def __init__(self, images: Optional[Sequence[Image]], patches: Optional[np.ndarray], device: int = -1,
                 flat_images=None, flat_patches=None, specs: Optional[np.ndarray] = None):
        """Either (images, patches) model objects, or pre-flattened descriptor arrays
        (`flat_images.arr/.N`, `flat_patches.arr/.S_tot/.N`, e.g. golden fixtures), or (images, specs): an
        S x N matrix of model.PatchSpec from which the device builds bitmaps and PSF splines itself."""
        lib = _lib.load()
        ndev = C.c_int(0)
        _lib.check(lib.celeste_init(device, C.byref(ndev)))
        self.images = list(images) if images is not None else None
        self.patches = patches
        self._flat_images = flat_images if flat_images is not None else FlatImages(self.images)
        self._handle = C.c_void_p()
        _lib.check(lib.celeste_field_create(C.byref(self._handle), self._flat_images.N, self._flat_images.arr))
        self._finalizer = weakref.finalize(self, lib.celeste_field_destroy, self._handle)
        self._row = {}
        if specs is not None:
            self.build_patches(specs)
        else:
            self.set_patches(patches, flat_patches)

    def build_patches(self, specs: np.ndarray):
        """celeste_patches_build: ImagePatch construction on the device (imaged_sources.jl:80-117) from
        model.PatchSpec objects; no bitmap / spline coefficient leaves or enters the host."""
        from .flatten import FlatPatchSpecs
        fs = FlatPatchSpecs(specs)
        _lib.check(_lib.load().celeste_patches_build(self._handle, fs.S_tot, fs.N, fs.arr))
        self.patches = None
        self.specs = specs
        self.S_tot = fs.S_tot
        self._row = {id(specs[s, 0]): s for s in range(specs.shape[0])} if specs.shape[1] else {}

    def patch_readback(self, s: int, n: int):
        """(active_pixel_bitmap H2 x W2 bool, itp coefficient array n1 x n2) of patch (s, n) (0-based) as resident on
        the device."""
        lib = _lib.load()
        dims = (C.c_int32 * 4)()
        _lib.check(lib.celeste_patch_readback(self._handle, s, n, dims, None, None))
        H2, W2, n1, n2 = dims
        bm = np.zeros((H2, W2), dtype=np.uint8, order="F")
        co = np.zeros((n1, n2), dtype=np.float64, order="F")
        _lib.check(lib.celeste_patch_readback(self._handle, s, n, dims, bm.ctypes.data if bm.size else None, co.ctypes.data))
        return bm.astype(bool), co

    def find_all_neighbors(self) -> List[List[int]]:
        """celeste_find_neighbors: model.find_all_neighbors (imaged_sources.jl:232-244) evaluated on the device's
        patch matrix; 0-based indices."""
        lib = _lib.load()
        ptr = np.zeros(self.S_tot + 1, dtype=np.int32)
        need = C.c_int64(0)
        st = lib.celeste_find_neighbors(self._handle, ptr.ctypes.data, None, 0, C.byref(need))
        if need.value == 0:
            _lib.check(st)
            return [[] for _ in range(self.S_tot)]
        nbr = np.zeros(need.value, dtype=np.int32)
        _lib.check(lib.celeste_find_neighbors(self._handle, ptr.ctypes.data, nbr.ctypes.data, need.value, C.byref(need)))
        return [nbr[ptr[t]:ptr[t + 1]].tolist() for t in range(self.S_tot)]

    def set_patches(self, patches: Optional[np.ndarray], flat_patches=None):
        lib = _lib.load()
        fp = flat_patches if flat_patches is not None else FlatPatches(patches)
        _lib.check(lib.celeste_patches_set(self._handle, fp.S_tot, fp.N, fp.arr))
        self.patches = patches
        self.S_tot = fp.S_tot
        self._row = {}
        if patches is not None and patches.shape[1]:
            self._row = {id(patches[s, 0]): s for s in range(patches.shape[0])}

    def elbo_csr(self, task_ptr, src, active_ptr, act, vp, mode: int = _lib.MODE_HESS, check_finite: bool = True):
        """celeste_elbo_batch on raw CSR arrays (the layout include/celeste_cuda.h documents)."""
        lib = _lib.load()
        n = len(task_ptr) - 1
        nd, nh = out_sizes(active_ptr)
        v = np.zeros(n)
        d = np.zeros(nd if mode >= 1 else 0)
        h = np.zeros(nh if mode >= 2 else 0)
        counters = np.zeros(2 * n, dtype=np.int64)
        flags = np.zeros(n, dtype=np.int32)

        def p(a):
            return a.ctypes.data if a.size else None
        st = lib.celeste_elbo_batch(self._handle, n, p(task_ptr), p(src), p(active_ptr), p(act), p(vp), mode,
                                    p(v), p(d), p(h), p(counters), p(flags))
        _lib.check(st, allow_nonfinite=not check_finite)
        return {"v": v, "d": d, "h": h, "counters": counters.reshape(n, 2), "flags": flags,
                "active_ptr": active_ptr}

    def row_of(self, patch: ImagePatch) -> Optional[int]:
        return self._row.get(id(patch))

    def elbo_batch(self, tasks, mode: int = _lib.MODE_HESS, check_finite: bool = True):
        """tasks: list of (rows_1based, active_local_1based, vp 44 x S).  Returns dict of arrays."""
        task_ptr, src, active_ptr, act, vp = csr_tasks(tasks)
        return self.elbo_csr(task_ptr, src, active_ptr, act, vp, mode, check_finite)

    def make_plan(self, tasks_rows, tasks_active):
        return Plan(self, tasks_rows, tasks_active)

    def render_expectation(self, rows, vp, full_box: bool = False) -> List[np.ndarray]:
        """celeste_render_expectation: for every image the H x W float64 array of E_G - sky (nanomaggies) summed
        over the sources `rows` (1-based rows of the patch matrix) at variational parameters vp (44 x S).
        full_box = celeste_render_boxes: every column of each box (Synthetic.gen_image!), not the ELBO's w2 < W2."""
        lib = _lib.load()
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        vpm = np.asfortranarray(vp, dtype=np.float64)
        S = len(rows)
        assert vpm.shape == (44, S) or S == 0
        fi = self._flat_images
        outs = [np.zeros((fi.arr[n].H, fi.arr[n].W), dtype=np.float64, order="F") for n in range(fi.N)]
        ptrs = (C.c_void_p * max(fi.N, 1))(*[o.ctypes.data for o in outs])
        fn = lib.celeste_render_boxes if full_box else lib.celeste_render_expectation
        _lib.check(fn(self._handle, S, rows.ctypes.data if S else None, vpm.ctypes.data if S else None, ptrs))
        return outs


def fill_celeste_expectation(images: Sequence[Image], patches: np.ndarray, vps, field: Optional[DeviceField] = None):
    """bin/write_celeste_expectation.jl:111-156 (fill_celeste_expectation!): add the model's expected source flux
    (nanomaggies) of ALL sources to every pixel of every image, `image.pixels[h, w] += E_G - sky[h, w]`
    (Float32 pixels as in the reference).  vps: the S variational parameter vectors, in patch-row order."""
    field = field or DeviceField(images, patches)
    S = patches.shape[0]
    add = field.render_expectation(np.arange(1, S + 1), np.stack(vps, axis=1) if S else np.zeros((44, 0)))
    for im, a in zip(images, add):
        im.pixels = (im.pixels.astype(np.float64) + a).astype(np.float32)     # Float32 += Float64, rounded once
    return images


class Plan:
    """A registered task list (celeste_plan): evaluate it repeatedly with new vp."""

    def __init__(self, field, tasks_rows, tasks_active, task_field=None):
        """`field`: one DeviceField, or a list of them (celeste_plan_create_multi) with `task_field[t]`
        the 0-based field of task t."""
        lib = _lib.load()
        self.fields = list(field) if isinstance(field, (list, tuple)) else [field]
        self.field = self.fields[0]
        dummy = [(r, a, np.zeros((44, len(r)))) for r, a in zip(tasks_rows, tasks_active)]
        self.task_ptr, self.src, self.active_ptr, self.act, _ = csr_tasks(dummy)
        self.n_tasks = len(dummy)
        self.n_src = int(self.task_ptr[-1])
        self.nd, self.nh = out_sizes(self.active_ptr)
        self.hess_packed = False
        self._handle = C.c_void_p()
        tf = np.zeros(self.n_tasks, dtype=np.int32) if task_field is None else np.asarray(task_field, dtype=np.int32)
        assert tf.shape == (self.n_tasks,)
        harr = (C.c_void_p * len(self.fields))(*[f._handle for f in self.fields])
        _lib.check(lib.celeste_plan_create_multi(len(self.fields), harr, C.byref(self._handle), self.n_tasks,
                                                 tf.ctypes.data, self.task_ptr.ctypes.data, self.src.ctypes.data,
                                                 self.active_ptr.ctypes.data, self.act.ctypes.data))
        self._finalizer = weakref.finalize(self, lib.celeste_plan_destroy, self._handle)

    def set_task_mask(self, mask_dev_ptr: int):
        """celeste_plan_set_task_mask: device pointer to n_tasks bytes (0 = skip the task), or 0 to clear."""
        _lib.check(_lib.load().celeste_plan_set_task_mask(self._handle, mask_dev_ptr or None))

    def enable_timing(self, on: bool = True):
        _lib.check(_lib.load().celeste_plan_enable_timing(self._handle, int(on)))

    def kernel_times_ms(self):
        """(setup, pixel, epilogue) device milliseconds of the last evaluation (synchronises)."""
        ms = (C.c_float * 3)()
        _lib.check(_lib.load().celeste_plan_kernel_times(self._handle, C.byref(ms)))
        return tuple(float(x) for x in ms)

    def unit_times_ms(self):
        """(unit_bg_kernel, unit_walk_kernel, unit_moment_kernel) device milliseconds of the last evaluation when the
        unit kernels ran (zeros otherwise)."""
        ms = (C.c_float * 3)()
        _lib.check(_lib.load().celeste_plan_unit_times(self._handle, C.byref(ms)))
        return tuple(float(x) for x in ms)

    def set_hessian_layout(self, packed: bool):
        """celeste_plan_set_hessian_layout: packed = 406 doubles per task (upper triangle of the live 28 x 28 block,
        row-major) instead of the dense 44 x 44 SensitiveFloat matrix.  Sa = 1 plans only."""
        _lib.check(_lib.load().celeste_plan_set_hessian_layout(self._handle, 1 if packed else 0))
        self.hess_packed = bool(packed)
        self.nh = (406 * self.n_tasks) if packed else out_sizes(self.active_ptr)[1]

    def launches(self, mode: int) -> int:
        return int(_lib.load().celeste_plan_launches(self._handle, mode))

    def kernel_name(self, mode: int) -> str:
        """Which kernel carries the pixel loop of `mode` (march_kernel / task_kernel / pixel_kernel)."""
        buf = C.create_string_buffer(32)
        _lib.check(_lib.load().celeste_plan_kernel_name(self._handle, mode, buf))
        return buf.value.decode()

    def run_host(self, vp_flat: np.ndarray, mode: int, out=None, check_finite=True):
        lib = _lib.load()
        n = self.n_tasks
        if out is None:
            out = {"v": np.zeros(n), "d": np.zeros(self.nd if mode >= 1 else 0),
                   "h": np.zeros(self.nh if mode >= 2 else 0),
                   "counters": np.zeros(2 * n, dtype=np.int64), "flags": np.zeros(n, dtype=np.int32)}

        def p(a):
            return a.ctypes.data if a.size else None
        st = lib.celeste_elbo_plan_host(self._handle, p(vp_flat), mode, p(out["v"]), p(out["d"]), p(out["h"]),
                                        p(out["counters"]), p(out["flags"]))
        _lib.check(st, allow_nonfinite=not check_finite)
        return out

    def run_device(self, vp_dev_ptr: int, mode: int, v_ptr: int, d_ptr: int, h_ptr: int,
                   counters_ptr: int, flags_ptr: int, stream: int = 0):
        st = _lib.load().celeste_elbo_plan_device(self._handle, vp_dev_ptr, mode, v_ptr, d_ptr or None, h_ptr or None,
                                                  counters_ptr, flags_ptr, stream or None)
        _lib.check(st)


_KL_CPU = None


def _kl_term_cpu():
    """One KLTerm for the process (it reads the prior file and inverts 16 covariance matrices when built)."""
    global _KL_CPU
    if _KL_CPU is None:
        from .kl import KLTerm
        _KL_CPU = KLTerm("cpu")
    return _KL_CPU


# ------------------------------------------------------------------ elbo_args.jl:165-211
class ElboArgs:
    """elbo_args.jl:165-211.  `patches` is the S x N object matrix of ImagePatch;
    `active_sources` are 1-based local indices exactly as in the reference."""

    def __init__(self, images: Sequence[Image], patches: np.ndarray, active_sources: Sequence[int],
                 psf_K: int = 2, include_kl: bool = True, field: Optional[DeviceField] = None):
        self.S = patches.shape[0]
        self.Sa = len(active_sources)
        self.N = len(images)
        assert patches.shape[1] == self.N
        assert psf_K > 0
        self.psf_K = psf_K
        self.images = list(images)
        self.patches = patches
        self.active_sources = list(active_sources)
        self.include_kl = include_kl
        self._field = field
        self._rows = None

    def device_rows(self):
        """(DeviceField, 1-based rows of ea.patches inside it); uploads on first use."""
        if self._field is not None and self._rows is None:
            rows = [self._field.row_of(self.patches[s, 0]) if self.N else None for s in range(self.S)]
            if any(r is None for r in rows):
                self._field = None
            else:
                self._rows = [r + 1 for r in rows]
        if self._field is None:
            self._field = DeviceField(self.images, self.patches)
            self._rows = list(range(1, self.S + 1))
        return self._field, self._rows


def _vp_matrix(vp: VariationalParams) -> np.ndarray:
    return np.stack([np.asarray(v, dtype=np.float64) for v in vp], axis=1)


def elbo_likelihood(ea: ElboArgs, vp: VariationalParams,
                    elbo_vars: Optional[ElboIntermediateVariables] = None) -> SensitiveFloat:
    """elbo_objective.jl:400-474 on the GPU.  Returns elbo_vars.elbo filled in place."""
    if elbo_vars is None:
        elbo_vars = ElboIntermediateVariables(ea.Sa)
    fieldobj, rows = ea.device_rows()
    res = elbo_vars.elbo
    out = fieldobj.elbo_batch([(rows, ea.active_sources, _vp_matrix(vp))], mode=res.mode)
    res.v = float(out["v"][0])
    if res.has_gradient:
        res.d[:, :] = out["d"].reshape((NUM_PARAMS, ea.Sa), order="F")
    if res.has_hessian:
        res.h[:, :] = out["h"].reshape((NUM_PARAMS * ea.Sa, NUM_PARAMS * ea.Sa), order="F")
    elbo_vars.active_pixel_counter += int(out["counters"][0, 0])
    elbo_vars.inactive_pixel_counter += int(out["counters"][0, 1])
    return res


def elbo(ea: ElboArgs, vp: VariationalParams,
         elbo_vars: Optional[ElboIntermediateVariables] = None) -> SensitiveFloat:
    """elbo_objective.jl:482-492: likelihood on the GPU, then `subtract_kl_all_sources!` (elbo_kl.jl:214-225)
    -- pixel-free, 44 numbers per active source -- added on the host exactly where the reference adds it."""
    for vs in vp:
        if not np.all(np.isfinite(vs)):
            raise AssertionError("vp contains NaNs or Infs")     # elbo_objective.jl:487
    res = elbo_likelihood(ea, vp, elbo_vars)
    if ea.include_kl:
        import torch
        from .kl import KLTerm
        order = 2 if res.has_hessian else (1 if res.has_gradient else 0)
        act = torch.tensor(np.stack([np.asarray(vp[s - 1], dtype=np.float64) for s in ea.active_sources]))
        kv, kg, kH = _kl_term_cpu()(act, order=order)
        res.v += float(kv.sum())
        for sa in range(ea.Sa):                                   # add_sources_sf!, SensitiveFloats.jl:215-250
            if res.has_gradient:
                res.d[:, sa] += kg[sa].numpy()
            if res.has_hessian:
                res.h[44 * sa:44 * (sa + 1), 44 * sa:44 * (sa + 1)] += kH[sa].numpy()
        if not (np.isfinite(res.v) and np.isfinite(res.d).all() and np.isfinite(res.h).all()):
            raise _lib.NonFiniteError(_lib.CELESTE_ERR_NONFINITE, "ELBO contains Inf/NaNs")   # :490
    return res
