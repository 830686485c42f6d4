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


# ---------------------------------------------------------------------------------------------------------------
# Reference-side dumps (tools/julia/dump_golden.jl): one self-describing little-endian file per case.
# Records: [int32 name_len][name][int32 dtype 0=f64 1=f32 2=i64 3=u8][int32 ndim][int64 dims...][column-major data]
_CELGOLD_DT = {0: np.dtype("<f8"), 1: np.dtype("<f4"), 2: np.dtype("<i8"), 3: np.dtype("u1")}


def read_celgold(path):
    """-> {name: ndarray (Fortran order, as Julia wrote it)}."""
    import struct
    b = open(path, "rb").read()
    pos, out = 0, {}
    while pos < len(b):
        (ln,) = struct.unpack_from("<i", b, pos)
        pos += 4
        name = b[pos:pos + ln].decode()
        pos += ln
        dt, nd = struct.unpack_from("<ii", b, pos)
        pos += 8
        dims = struct.unpack_from("<" + "q" * nd, b, pos)
        pos += 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        a = np.frombuffer(b, dtype=_CELGOLD_DT[dt], count=cnt, offset=pos).reshape(dims, order="F")
        pos += cnt * _CELGOLD_DT[dt].itemsize
        out[name] = a
    return out


def write_celgold(path, records):
    """The Python twin of dump_golden.jl's `rec` (used by the tests to exercise the reader without Julia)."""
    import struct
    code = {np.dtype("float64"): 0, np.dtype("float32"): 1, np.dtype("int64"): 2, np.dtype("uint8"): 3}
    with open(path, "wb") as f:
        for name, a in records:
            a = np.asarray(a)
            if a.ndim == 0:
                a = a.reshape(1)
            f.write(struct.pack("<i", len(name)) + name.encode())
            f.write(struct.pack("<ii", code[a.dtype], a.ndim))
            f.write(struct.pack("<" + "q" * a.ndim, *a.shape))
            f.write(np.asfortranarray(a).tobytes(order="F"))


def load_julia_dump(path):
    """A dump_golden.jl file -> (flat_images, flat_patches, csr, {mode: reference outputs}, extras).
    Images get log_iota = log(Float32 iota) in Float32 like elbo_objective.jl:292."""
    z = read_celgold(path)
    N, S = int(z["N"][0]), int(z["S"][0])
    fi, fp = _Flat(), _Flat()
    fi.N, fi.arr, fi._keep = N, (celeste_image * max(N, 1))(), []
    for n in range(N):
        H, W, band = (int(x) for x in z[f"img{n + 1}_meta"])
        px = np.asfortranarray(z[f"img{n + 1}_pixels"], dtype=np.float32)
        sky = np.asfortranarray(z[f"img{n + 1}_sky"], dtype=np.float32)
        iota = np.ascontiguousarray(z[f"img{n + 1}_iota"], dtype=np.float32)
        li = np.log(iota).astype(np.float32).astype(np.float64)
        fi._keep += [px, sky, iota, li]
        a = fi.arr[n]
        a.H, a.W, a.band = H, W, band
        a.pixels, a.sky, a.nelec_per_nmgy, a.log_iota = px.ctypes.data, sky.ctypes.data, iota.ctypes.data, li.ctypes.data
    fp.S_tot, fp.N, fp.arr, fp._keep = S, N, (celeste_patch * max(S * N, 1))(), []
    for n in range(N):
        for s in range(S):
            key = f"p{s + 1}_{n + 1}"
            q = fp.arr[s + n * S]
            bm = np.asfortranarray(z[key + "_bitmap"], dtype=np.uint8)
            psf = np.asfortranarray(z[key + "_psf"], dtype=np.float64)
            co = np.asfortranarray(z[key + "_itp_coefs"], dtype=np.float64)
            fp._keep += [bm, psf, co]
            q.bitmap_offset[0], q.bitmap_offset[1] = (int(x) for x in z[key + "_offset"])
            q.H2, q.W2 = bm.shape if bm.ndim == 2 else (0, 0)
            q.K = psf.shape[1]
            q.active_pixel_bitmap = bm.ctypes.data if bm.size else None
            q.psf = psf.ctypes.data
            q.itp_coefs = co.ctypes.data
            q.itp_dims[0], q.itp_dims[1] = co.shape
            J = z[key + "_wcs_jacobian"]
            q.wcs_jacobian[0], q.wcs_jacobian[1], q.wcs_jacobian[2], q.wcs_jacobian[3] = J[0, 0], J[1, 0], J[0, 1], J[1, 1]
            q.world_center[0], q.world_center[1] = z[key + "_world_center"]
            q.pixel_center[0], q.pixel_center[1] = z[key + "_pixel_center"]
    act = np.asarray(z["active_sources"], dtype=np.int32)
    csr = (np.array([0, S], dtype=np.int32), np.arange(1, S + 1, dtype=np.int32), np.array([0, len(act)], dtype=np.int32),
           act, np.asfortranarray(z["vp"], dtype=np.float64).ravel(order="F"))
    outs = {}
    for mode in (0, 1, 2):
        if f"out{mode}_v" in z:
            outs[mode] = {"v": np.asarray(z[f"out{mode}_v"], dtype=np.float64).reshape(1),
                          "d": z[f"out{mode}_d"].ravel(order="F") if mode >= 1 else np.zeros(0),
                          "h": z[f"out{mode}_h"].ravel(order="F") if mode >= 2 else np.zeros(0),
                          "counters": np.asarray(z[f"out{mode}_counters"], dtype=np.int64).reshape(1, 2),
                          "flags": np.zeros(1, dtype=np.int32)}
    extras = {k: z[k] for k in ("elbo_kl_v", "elbo_kl_d", "elbo_kl_h") if k in z}
    return fi, fp, csr, outs, extras
