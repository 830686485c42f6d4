"""Golden vectors (tests/golden/*.npz, made by tools/make_golden.py from the pinned oracle)."""
import glob
import os

import numpy as np
import pytest

import cases
import emul_lib
import golden_io
import oracle_lib

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def test_fixtures_present():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_reproduces_golden(path):
    fi, fp, csr, outs = golden_io.load(path)
    of = oracle_lib.OracleField(None, None, flat_images=fi, flat_patches=fp)
    for mode, ref in outs.items():
        got = of.elbo_csr(*csr, mode=mode)
        assert np.array_equal(ref["counters"], got["counters"]) and np.array_equal(ref["flags"], got["flags"])
        for k in ("v", "d", "h"):
            sc = np.abs(ref[k]).max() if ref[k].size else 1.0
            assert np.all(np.abs(ref[k] - got[k]) <= 1e-12 * max(sc, 1e-300)), (k, mode)


@pytest.mark.parametrize("path", GOLDEN[:3], ids=IDS[:3])
def test_emulated_kernels_match_golden(path):
    fi, fp, csr, outs = golden_io.load(path)
    got = emul_lib.EmulField(None, None, flat_images=fi, flat_patches=fp).elbo_csr(*csr, mode=2)
    cases.assert_parity(outs[2], got, 2, path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_cuda_matches_golden(path, mode):
    import celeste_jl_b200 as cj
    fi, fp, csr, outs = golden_io.load(path)
    field = cj.DeviceField(None, None, flat_images=fi, flat_patches=fp)
    got = field.elbo_csr(*csr, mode=mode)
    cases.assert_parity(outs[mode], got, mode, path)


# ---------------------------------------------------------------------------------------------------------------
# Reference-held vectors.  tools/julia/dump_golden.jl (run by anyone with Julia 0.6 + Celeste.jl) writes
# tests/golden/julia/*.celgold; the day such a file is committed these tests pin the oracle -- and through it every
# CUDA parity test -- to numbers the reference itself computed.  Until then the reader / comparison pipeline is
# exercised with files written in the same format by the Python twin of the Julia writer.
JULIA = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "julia", "*.celgold")))


def _as_celgold_records(name):
    """What dump_golden.jl's dump_case writes, produced from a tests/cases.py scene with the oracle in the role of
    the reference (all sources active, like SampleData.make_elbo_args)."""
    images, patches, tasks = cases.get(name)
    S, N = patches.shape
    vp = cases.all_vp(patches, tasks)
    act = list(range(1, S + 1))
    rec = [("N", np.array([N], dtype=np.int64)), ("S", np.array([S], dtype=np.int64)),
           ("active_sources", np.array(act, dtype=np.int64))]
    for n, im in enumerate(images):
        rec += [(f"img{n + 1}_meta", np.array([im.H, im.W, im.b], dtype=np.int64)),
                (f"img{n + 1}_pixels", np.asarray(im.pixels, dtype=np.float32)),
                (f"img{n + 1}_sky", np.asarray(im.sky, dtype=np.float32)),
                (f"img{n + 1}_iota", np.asarray(im.nelec_per_nmgy, dtype=np.float32))]
    for s in range(S):
        for n in range(N):
            p, key = patches[s, n], f"p{s + 1}_{n + 1}"
            rec += [(key + "_offset", np.array(p.bitmap_offset, dtype=np.int64)),
                    (key + "_bitmap", np.asarray(p.active_pixel_bitmap, dtype=np.uint8)),
                    (key + "_wcs_jacobian", np.asarray(p.wcs_jacobian, dtype=np.float64)),
                    (key + "_world_center", np.asarray(p.world_center, dtype=np.float64)),
                    (key + "_pixel_center", np.asarray(p.pixel_center, dtype=np.float64)),
                    (key + "_psf", np.stack([pc.flat7() for pc in p.psf], axis=1)),
                    (key + "_itp_coefs", np.asarray(p.itp_coefs, dtype=np.float64))]
    rec.append(("vp", vp))
    of = oracle_lib.OracleField(images, patches)
    for mode in (0, 1, 2):
        o = of.elbo_batch([(act, act, vp)], mode=mode)
        rec.append((f"out{mode}_v", o["v"]))
        if mode >= 1:
            rec.append((f"out{mode}_d", o["d"].reshape((44, S), order="F")))
        if mode >= 2:
            rec.append((f"out{mode}_h", o["h"].reshape((44 * S, 44 * S), order="F")))
        rec.append((f"out{mode}_counters", o["counters"].reshape(2).astype(np.int64)))
    return rec


def _check_dump(path, evaluate):
    fi, fp, csr, outs, _ = golden_io.load_julia_dump(path)
    assert outs, "dump holds no outputs"
    for mode, ref in outs.items():
        got = evaluate(fi, fp, csr, mode)
        cases.assert_parity(ref, got, mode, f"{os.path.basename(path)} mode {mode}")


@pytest.mark.parametrize("name", ["two_body", "masked"])
def test_celgold_reader_round_trip(tmp_path, name):
    path = str(tmp_path / (name + ".celgold"))
    golden_io.write_celgold(path, _as_celgold_records(name))
    _check_dump(path, lambda fi, fp, csr, mode: oracle_lib.OracleField(None, None, flat_images=fi, flat_patches=fp)
                .elbo_csr(*csr, mode=mode))


@pytest.mark.parametrize("path", JULIA, ids=[os.path.basename(p) for p in JULIA])
def test_oracle_matches_julia_dumps(path):
    """Pins the oracle to the reference's own numbers (empty until someone with Julia commits a dump)."""
    _check_dump(path, lambda fi, fp, csr, mode: oracle_lib.OracleField(None, None, flat_images=fi, flat_patches=fp)
                .elbo_csr(*csr, mode=mode))


@pytest.mark.gpu
@pytest.mark.parametrize("path", JULIA, ids=[os.path.basename(p) for p in JULIA])
def test_cuda_matches_julia_dumps(path):
    import celeste_jl_b200 as cj
    _check_dump(path, lambda fi, fp, csr, mode: cj.DeviceField(None, None, flat_images=fi, flat_patches=fp)
                .elbo_csr(*csr, mode=mode))
