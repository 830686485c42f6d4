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
