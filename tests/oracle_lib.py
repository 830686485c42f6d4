"""ctypes binding of oracle/libceleste_oracle.so -- the CHECKER (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

import celeste_jl_b200 as cj
from celeste_jl_b200.flatten import FlatImages, FlatPatches, csr_tasks, out_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libceleste_oracle.so")
_lib = None


def load(path=None):
    global _lib
    if path is None and _lib is not None:
        return _lib
    p = path or ORACLE_SO
    if path is None and not os.path.exists(p):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    lib = C.CDLL(p)
    vp, i32 = C.c_void_p, C.c_int32
    lib.oracle_elbo_batch.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32]
    lib.oracle_get_bvn_cov.argtypes = [C.c_double, C.c_double, C.c_double, vp]
    lib.oracle_get_bvn_cov.restype = None
    lib.oracle_spline_eval.argtypes = [vp, i32, i32, C.c_double, C.c_double, vp]
    lib.oracle_spline_eval.restype = None
    lib.oracle_galaxy_prototypes.argtypes = [vp, vp]
    lib.oracle_galaxy_prototypes.restype = None
    lib.oracle_render_expectation.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp, i32]
    lib.oracle_calculate_G_s_probe.argtypes = [i32, vp, i32, vp, i32, vp, vp, i32, vp, vp, vp]
    if path is None:
        _lib = lib
    return lib


class OracleField:
    """Same call shape as cj.DeviceField.elbo_batch, evaluated by the CPU oracle."""

    def __init__(self, images, patches, lib=None, flat_images=None, flat_patches=None):
        self.fi = flat_images if flat_images is not None else FlatImages(images)
        self.fp = flat_patches if flat_patches is not None else FlatPatches(patches)
        self.lib = lib or load()

    def elbo_batch(self, tasks, mode=2, n_threads=1):
        task_ptr, src, active_ptr, act, vp = csr_tasks(tasks)
        return self.elbo_csr(task_ptr, src, active_ptr, act, vp, mode, n_threads)

    def elbo_csr(self, task_ptr, src, active_ptr, act, vp, mode=2, n_threads=1):
        n = len(task_ptr) - 1
        nd, nh = out_sizes(active_ptr)
        v = np.zeros(n)
        d = np.zeros(nd if mode >= 1 else 0)
        h = np.zeros(nh if mode >= 2 else 0)
        counters = np.zeros(2 * n, dtype=np.int64)
        flags = np.zeros(n, dtype=np.int32)
        p = lambda a: a.ctypes.data if a.size else None
        st = self.lib.oracle_elbo_batch(self.fi.N, C.addressof(self.fi.arr), self.fp.S_tot, C.addressof(self.fp.arr),
                                        n, p(task_ptr), p(src), p(active_ptr), p(act), p(vp), mode,
                                        p(v), p(d), p(h), p(counters), p(flags), n_threads)
        assert st == 0
        return {"v": v, "d": d, "h": h, "counters": counters.reshape(n, 2), "flags": flags, "active_ptr": active_ptr}


def _render(fn, fi, fp, rows, vp, *extra):
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    vpm = np.asfortranarray(vp, dtype=np.float64)
    outs = [np.zeros((fi.arr[n].H, fi.arr[n].W), dtype=np.float64, order="F") for n in range(fi.N)]
    ptrs = (C.c_void_p * max(fi.N, 1))(*[o.ctypes.data for o in outs])
    st = fn(fi.N, C.addressof(fi.arr), fp.S_tot, C.addressof(fp.arr), len(rows), rows.ctypes.data if len(rows) else None,
            vpm.ctypes.data if len(rows) else None, ptrs, *extra)
    assert st == 0
    return outs


def oracle_render_expectation(images, patches, rows, vp, n_threads=8):
    """fill_celeste_expectation! by the oracle: per image the H x W array of E_G - sky."""
    of = OracleField(images, patches)
    return _render(of.lib.oracle_render_expectation, of.fi, of.fp, rows, vp, n_threads)


def oracle_elbo(images, patches, vp, active_sources, mode=2):
    """elbo_likelihood(ElboArgs(images, patches, active_sources), vp) by the oracle.
    Returns (v, d[44, Sa], h[44 Sa, 44 Sa], counters)."""
    S = patches.shape[0]
    of = OracleField(images, patches)
    vpm = np.stack(vp, axis=1)
    out = of.elbo_batch([(list(range(1, S + 1)), list(active_sources), vpm)], mode=mode)
    Sa = len(active_sources)
    d = out["d"].reshape((44, Sa), order="F") if mode >= 1 else None
    h = out["h"].reshape((44 * Sa, 44 * Sa), order="F") if mode >= 2 else None
    return float(out["v"][0]), d, h, out["counters"][0]
