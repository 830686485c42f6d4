"""Golden fixtures: flat ABI-level inputs + oracle outputs, stored as .npz under tests/golden/.

The reference's test-suite holds no stored numbers for this path and the reference cannot run
offline, so the golden vectors are produced by the oracle (tools/make_golden.py)
after the oracle itself has been pinned against autograd (tests/test_oracle.py).  A fixture stores
the *flattened descriptors* (exactly what crosses the C ABI), so it does not depend on the Python
model code or on RNG reproducibility.
"""
import ctypes as C

import numpy as np

from celeste_jl_b200._lib import celeste_image, celeste_patch

IMG_KEYS = ("H", "W", "band")


def dump(path, fi, fp, csr, outs):
    """fi/fp: FlatImages/FlatPatches; csr: (task_ptr, src, active_ptr, act, vp); outs: {mode: result dict}."""
    z = {}
    N, S_tot = fi.N, fp.S_tot
    z["N"], z["S_tot"] = N, S_tot
    for n in range(N):
        a = fi.arr[n]
        H, W = a.H, a.W
        z[f"img{n}_meta"] = np.array([H, W, a.band])
        z[f"img{n}_pixels"] = np.ctypeslib.as_array(C.cast(a.pixels, C.POINTER(C.c_float)), (H * W,)).copy()
        z[f"img{n}_sky"] = np.ctypeslib.as_array(C.cast(a.sky, C.POINTER(C.c_float)), (H * W,)).copy()
        z[f"img{n}_iota"] = np.ctypeslib.as_array(C.cast(a.nelec_per_nmgy, C.POINTER(C.c_float)), (H,)).copy()
        z[f"img{n}_logiota"] = np.ctypeslib.as_array(C.cast(a.log_iota, C.POINTER(C.c_double)), (H,)).copy()
    coef_tables, coef_index = [], {}
    meta = np.zeros((S_tot * N, 8), dtype=np.int64)
    geo = np.zeros((S_tot * N, 8))
    for i in range(S_tot * N):
        q = fp.arr[i]
        nb = q.H2 * q.W2
        z[f"p{i}_bitmap"] = (np.ctypeslib.as_array(C.cast(q.active_pixel_bitmap, C.POINTER(C.c_uint8)), (nb,)).copy()
                            if nb else np.zeros(0, dtype=np.uint8))
        z[f"p{i}_psf"] = np.ctypeslib.as_array(C.cast(q.psf, C.POINTER(C.c_double)), (7 * q.K,)).copy()
        key = (q.itp_coefs, q.itp_dims[0], q.itp_dims[1])
        if key not in coef_index:
            coef_index[key] = len(coef_tables)
            coef_tables.append(np.ctypeslib.as_array(C.cast(q.itp_coefs, C.POINTER(C.c_double)),
                                                     (q.itp_dims[0] * q.itp_dims[1],)).copy())
        meta[i] = [q.bitmap_offset[0], q.bitmap_offset[1], q.H2, q.W2, q.K, q.itp_dims[0], q.itp_dims[1],
                   coef_index[key]]
        geo[i] = list(q.wcs_jacobian) + list(q.world_center) + list(q.pixel_center)
    z["patch_meta"], z["patch_geo"] = meta, geo
    for k, t in enumerate(coef_tables):
        z[f"coefs{k}"] = t
    z["n_coefs"] = len(coef_tables)
    for name, a in zip(("task_ptr", "src", "active_ptr", "act", "vp"), csr):
        z["csr_" + name] = a
    for mode, o in outs.items():
        for k in ("v", "d", "h", "counters", "flags"):
            z[f"out{mode}_{k}"] = o[k]
    np.savez_compressed(path, **z)


class _Flat:
    pass


def load(path):
    """-> (flat_images, flat_patches, csr tuple, {mode: outputs}) with ctypes arrays rebuilt."""
    z = np.load(path)
    N, S_tot = int(z["N"]), int(z["S_tot"])
    fi, fp = _Flat(), _Flat()
    fi.N, fi.arr, fi._keep = N, (celeste_image * max(N, 1))(), []
    for n in range(N):
        H, W, band = (int(x) for x in z[f"img{n}_meta"])
        arrs = [np.ascontiguousarray(z[f"img{n}_{k}"]) for k in ("pixels", "sky", "iota", "logiota")]
        fi._keep += arrs
        a = fi.arr[n]
        a.H, a.W, a.band = H, W, band
        a.pixels, a.sky, a.nelec_per_nmgy, a.log_iota = (x.ctypes.data for x in arrs)
    coefs = [np.ascontiguousarray(z[f"coefs{k}"]) for k in range(int(z["n_coefs"]))]
    fp.S_tot, fp.N, fp.arr, fp._keep = S_tot, N, (celeste_patch * max(S_tot * N, 1))(), coefs
    meta, geo = z["patch_meta"], z["patch_geo"]
    for i in range(S_tot * N):
        q = fp.arr[i]
        bm = np.ascontiguousarray(z[f"p{i}_bitmap"])
        psf = np.ascontiguousarray(z[f"p{i}_psf"])
        fp._keep += [bm, psf]
        q.bitmap_offset[0], q.bitmap_offset[1] = int(meta[i, 0]), int(meta[i, 1])
        q.H2, q.W2, q.K = int(meta[i, 2]), int(meta[i, 3]), int(meta[i, 4])
        q.itp_dims[0], q.itp_dims[1] = int(meta[i, 5]), int(meta[i, 6])
        q.active_pixel_bitmap = bm.ctypes.data if bm.size else None
        q.psf = psf.ctypes.data
        q.itp_coefs = coefs[int(meta[i, 7])].ctypes.data
        for k in range(4):
            q.wcs_jacobian[k] = geo[i, k]
        q.world_center[0], q.world_center[1] = geo[i, 4], geo[i, 5]
        q.pixel_center[0], q.pixel_center[1] = geo[i, 6], geo[i, 7]
    csr = tuple(np.ascontiguousarray(z["csr_" + k]) for k in ("task_ptr", "src", "active_ptr", "act", "vp"))
    outs = {}
    for mode in (0, 1, 2):
        if f"out{mode}_v" in z:
            outs[mode] = {k: z[f"out{mode}_{k}"] for k in ("v", "d", "h", "counters", "flags")}
    return fi, fp, csr, outs
