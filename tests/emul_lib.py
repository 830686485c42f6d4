"""ctypes binding of tests/host_emul/libceleste_emul.so: the product's kernel source run under a
host emulation of the CUDA execution model (test infrastructure; lets kernel logic be checked
against the oracle without a GPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

from celeste_jl_b200.flatten import FlatImages, FlatPatches, csr_tasks, out_sizes

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
SO = os.path.join(DIR, "libceleste_emul.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", DIR, "-s"])
        _lib = C.CDLL(SO)
        vp, i32 = C.c_void_p, C.c_int32
        _lib.emul_elbo_batch.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32]
        _lib.emul_set_grad_kernel.argtypes = [i32]
        _lib.emul_set_hess_kernel.argtypes = [i32]
        _lib.emul_set_unit_target.argtypes = [C.c_int64]
        _lib.emul_march_blocks.argtypes = [i32, i32, vp, i32, vp, vp, C.c_int64, i32, vp, vp]
        _lib.emul_tr_subproblem.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp]
        _lib.emul_newton_step.argtypes = [i32, i32, vp]
        _lib.emul_spline_build.argtypes = [i32, vp, i32, vp, vp]
        _lib.emul_bitmap_build.argtypes = [i32, i32, vp, i32, i32, i32, i32, vp]
        _lib.emul_find_neighbors.argtypes = [i32, i32, vp, vp, vp]
        _lib.emul_render_expectation.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp]
        _lib.emul_render_boxes.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp]
    return _lib


class EmulField:
    def __init__(self, images, patches, flat_images=None, flat_patches=None):
        self.fi = flat_images if flat_images is not None else FlatImages(images)
        self.fp = flat_patches if flat_patches is not None else FlatPatches(patches)

    def elbo_batch(self, tasks, mode=2, chunk_pixels=512):
        return self.elbo_csr(*csr_tasks(tasks), mode=mode, chunk_pixels=chunk_pixels)

    def elbo_csr(self, task_ptr, src, active_ptr, act, vp, mode=2, chunk_pixels=512):
        n = len(task_ptr) - 1
        nd, nh = out_sizes(active_ptr)
        v = np.zeros(n)
        d = np.zeros(max(nd, 1))
        h = np.zeros(max(nh, 1))
        counters = np.zeros(2 * n, dtype=np.int64)
        flags = np.zeros(n, dtype=np.int32)
        p = lambda a: a.ctypes.data if a.size else None
        st = load().emul_elbo_batch(self.fi.N, C.addressof(self.fi.arr), self.fp.S_tot, C.addressof(self.fp.arr),
                                    n, p(task_ptr), p(src), p(active_ptr), p(act), p(vp), mode,
                                    p(v), p(d), p(h), p(counters), p(flags), chunk_pixels)
        assert st == 0, st
        return {"v": v, "d": d[:nd] if mode >= 1 else d[:0], "h": h[:nh] if mode >= 2 else h[:0],
                "counters": counters.reshape(n, 2), "flags": flags, "active_ptr": active_ptr}


def tr_subproblem(g, H, delta):
    """tr_subproblem_kernel under emulation: g B x n, H B x n x n, delta B -> (s, m, interior)."""
    g = np.ascontiguousarray(g, dtype=np.float64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    B, n = g.shape
    s = np.zeros((B, n))
    m = np.zeros(B)
    interior = np.zeros(B, dtype=np.int32)
    st = load().emul_tr_subproblem(B, n, g.ctypes.data, H.ctypes.data, delta.ctypes.data, s.ctypes.data,
                                   m.ctypes.data, interior.ctypes.data)
    assert st == 0
    return s, m, interior.astype(bool)


def newton_stepper(phase, n, buffers):
    """newton_step_kernel under emulation, as the `stepper` hook of elbo_maximize.BatchMaximizer (CPU tensors)."""
    st = load().emul_newton_step(phase, n, C.addressof(buffers))
    assert st == 0


def render_expectation(images, patches, rows, vp, full_box=False):
    """render_kernel (+ setup_kernel, host tile binning) under emulation."""
    from oracle_lib import _render
    fi, fp = FlatImages(images), FlatPatches(patches)
    return _render(load().emul_render_boxes if full_box else load().emul_render_expectation, fi, fp, rows, vp)


def spline_build(grid_n, raw=None, psf=None):
    """spline_build_kernel under emulation: raw stamp (n x n) or PSF mixture (list of PsfComponent) -> (n+2)^2 coefficients."""
    coefs = np.zeros((grid_n + 2, grid_n + 2), order="F")
    rawa = np.asfortranarray(raw, dtype=np.float64) if raw is not None else None
    psfa = np.ascontiguousarray(np.concatenate([pc.flat7() for pc in psf]), dtype=np.float64) if psf is not None else np.zeros(7)
    load().emul_spline_build(grid_n, rawa.ctypes.data if rawa is not None else None, len(psf) if psf is not None else 0,
                             psfa.ctypes.data, coefs.ctypes.data)
    return coefs


def bitmap_build(pixels, off_h, off_w, H2, W2):
    px = np.asfortranarray(pixels, dtype=np.float32)
    bm = np.zeros((H2, W2), dtype=np.uint8, order="F")
    load().emul_bitmap_build(px.shape[0], px.shape[1], px.ctypes.data, off_h, off_w, H2, W2, bm.ctypes.data if bm.size else None)
    return bm.astype(bool)


def find_all_neighbors(patches):
    """neighbor_kernel under emulation on an S x N matrix of objects with bitmap_offset and shape / (H2, W2)."""
    S, N = patches.shape
    boxes = np.zeros((N, S, 4), dtype=np.int32)
    for n in range(N):
        for s in range(S):
            p = patches[s, n]
            H2, W2 = (p.H2, p.W2) if hasattr(p, "H2") else p.active_pixel_bitmap.shape
            boxes[n, s] = [p.bitmap_offset[0], p.bitmap_offset[1], H2, W2]
    ptr = np.zeros(S + 1, dtype=np.int32)
    need = load().emul_find_neighbors(S, N, boxes.ctypes.data, ptr.ctypes.data, None)
    nbr = np.zeros(max(need, 1), dtype=np.int32)
    load().emul_find_neighbors(S, N, boxes.ctypes.data, ptr.ctypes.data, nbr.ctypes.data)
    return [nbr[ptr[t]:ptr[t + 1]].tolist() for t in range(S)]


def march_blocks(patches, tasks, split=10**9):
    """build_march_blocks (host half of march_kernels.cuh) -> (list of block dicts, part_ptr)."""
    fp = FlatPatches(patches)
    task_ptr, src, _active_ptr, _act, _vp = csr_tasks(tasks)
    n = len(task_ptr) - 1
    cap = 64 * max(n, 1) * max(fp.N, 1)
    out = np.zeros(12 * cap, dtype=np.int32)
    part = np.zeros(n + 1, dtype=np.int32)
    nb = load().emul_march_blocks(fp.N, fp.S_tot, C.addressof(fp.arr), n, task_ptr.ctypes.data, src.ctypes.data, int(split), cap,
                                  out.ctypes.data, part.ctypes.data)
    assert nb >= 0, nb
    keys = ["aslot", "slot0", "slot1", "n0", "n1", "pidx", "nseg", "hasbg", "walks", "sub", "task", "npair"]
    return [dict(zip(keys, out[12 * i:12 * i + 12].tolist())) for i in range(nb)], part
