"""Flatten `Image` / `ImagePatch` objects into the plain C structs of include/celeste_cuda.h.

This is the Python twin of the descriptor-building half of the Julia shim
(INTEGRATION.md): done once per inference box, not per evaluation.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from ._lib import celeste_image, celeste_patch, celeste_patch_spec
from .model import Image, ImagePatch


class FlatImages:
    """Owns the numpy buffers the `celeste_image` array points into."""

    def __init__(self, images: Sequence[Image]):
        self.images = list(images)
        self.N = len(self.images)
        self.arr = (celeste_image * max(self.N, 1))()
        self._keep = []
        for n, im in enumerate(self.images):
            px = np.asfortranarray(im.pixels, dtype=np.float32)
            sky = np.asfortranarray(im.sky, dtype=np.float32)
            iota = np.ascontiguousarray(im.nelec_per_nmgy, dtype=np.float32)
            li = np.ascontiguousarray(im.log_iota(), dtype=np.float64)
            self._keep += [px, sky, iota, li]
            a = self.arr[n]
            a.H, a.W, a.band = im.H, im.W, im.b
            a.pixels = px.ctypes.data
            a.sky = sky.ctypes.data
            a.nelec_per_nmgy = iota.ctypes.data
            a.log_iota = li.ctypes.data


class FlatPatches:
    """S_tot x N patch matrix as a column-major `celeste_patch` array."""

    def __init__(self, patches: np.ndarray):
        assert patches.ndim == 2
        self.S_tot, self.N = patches.shape
        self.arr = (celeste_patch * max(self.S_tot * self.N, 1))()
        self._keep = []
        for n in range(self.N):
            for s in range(self.S_tot):
                p: ImagePatch = patches[s, n]
                a = self.arr[s + n * self.S_tot]
                bm = np.asfortranarray(p.active_pixel_bitmap, dtype=np.uint8)
                psf = np.ascontiguousarray(np.concatenate([pc.flat7() for pc in p.psf]), dtype=np.float64)
                coefs = p.itp_coefs
                if not (coefs.flags.f_contiguous and coefs.dtype == np.float64):
                    coefs = np.asfortranarray(coefs, dtype=np.float64)
                self._keep += [bm, psf, coefs]
                a.bitmap_offset[0] = int(p.bitmap_offset[0])
                a.bitmap_offset[1] = int(p.bitmap_offset[1])
                a.H2, a.W2 = bm.shape if bm.ndim == 2 else (0, 0)
                a.active_pixel_bitmap = bm.ctypes.data if bm.size else None
                J = np.asarray(p.wcs_jacobian, dtype=np.float64)
                a.wcs_jacobian[0], a.wcs_jacobian[1] = J[0, 0], J[1, 0]
                a.wcs_jacobian[2], a.wcs_jacobian[3] = J[0, 1], J[1, 1]
                a.world_center[0], a.world_center[1] = p.world_center
                a.pixel_center[0], a.pixel_center[1] = p.pixel_center
                a.K = len(p.psf)
                a.psf = psf.ctypes.data
                a.itp_coefs = coefs.ctypes.data
                a.itp_dims[0], a.itp_dims[1] = coefs.shape


class FlatPatchSpecs:
    """S_tot x N PatchSpec matrix as a column-major `celeste_patch_spec` array (celeste_patches_build)."""

    def __init__(self, specs: np.ndarray):
        assert specs.ndim == 2
        self.S_tot, self.N = specs.shape
        self.arr = (celeste_patch_spec * max(self.S_tot * self.N, 1))()
        self._keep = []
        stamps = {}                                           # id(array) -> contiguous copy (identical pointers dedupe)
        psfs = {}
        for n in range(self.N):
            for s in range(self.S_tot):
                p = specs[s, n]
                a = self.arr[s + n * self.S_tot]
                key = id(p.psf)
                if key not in psfs:
                    psfs[key] = np.ascontiguousarray(np.concatenate([pc.flat7() for pc in p.psf]), dtype=np.float64)
                psf = psfs[key]
                a.bitmap_offset[0] = int(p.bitmap_offset[0])
                a.bitmap_offset[1] = int(p.bitmap_offset[1])
                a.H2, a.W2 = int(p.H2), int(p.W2)
                J = np.asarray(p.wcs_jacobian, dtype=np.float64)
                a.wcs_jacobian[0], a.wcs_jacobian[1] = J[0, 0], J[1, 0]
                a.wcs_jacobian[2], a.wcs_jacobian[3] = J[0, 1], J[1, 1]
                a.world_center[0], a.world_center[1] = p.world_center
                a.pixel_center[0], a.pixel_center[1] = p.pixel_center
                a.K = len(p.psf)
                a.grid_n = int(p.grid_n)
                a.psf = psf.ctypes.data
                if p.grid_psf is not None:
                    k2 = id(p.grid_psf)
                    if k2 not in stamps:
                        stamps[k2] = np.asfortranarray(p.grid_psf, dtype=np.float64)
                    a.grid_psf = stamps[k2].ctypes.data
                else:
                    a.grid_psf = None
        self._keep += list(stamps.values()) + list(psfs.values())


def csr_tasks(tasks):
    """tasks: list of (source_rows_1based, active_local_1based, vp 44 x S array).

    Returns (task_ptr, source_ids, active_ptr, active_idx, vp_flat) as numpy arrays with
    the layout celeste_elbo_batch documents."""
    task_ptr = [0]
    active_ptr = [0]
    src: List[int] = []
    act: List[int] = []
    vps = []
    for rows, active, vp in tasks:
        rows = list(rows)
        active = list(active)
        vp = np.asarray(vp, dtype=np.float64)
        assert vp.shape == (44, len(rows)), (vp.shape, len(rows))
        src += rows
        act += active
        task_ptr.append(len(src))
        active_ptr.append(len(act))
        vps.append(np.asfortranarray(vp).ravel(order="F"))
    return (np.array(task_ptr, dtype=np.int32), np.array(src, dtype=np.int32),
            np.array(active_ptr, dtype=np.int32), np.array(act, dtype=np.int32),
            np.concatenate(vps) if vps else np.zeros(0))


def out_sizes(active_ptr: np.ndarray):
    sa = np.diff(active_ptr).astype(np.int64)
    return int((44 * sa).sum()), int(((44 * sa) ** 2).sum())
