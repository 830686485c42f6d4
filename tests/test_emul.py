"""The product's kernel SOURCE (celeste.jl_b200/csrc/celeste_kernels.cuh), compiled for the host under
tests/host_emul's emulation of the CUDA execution model, against the oracle -- CPU only.  This is not
the product path (the product only runs these kernels on a GPU); it catches indexing / reduction /
chain-rule bugs before GPU time is spent.  The real parity tests are tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cases
import emul_lib
import oracle_lib

FAST = ["star_1band", "two_body", "masked", "clipped_and_empty", "psf_k1", "psf_k3", "crowded"]


@pytest.mark.parametrize("name", FAST)
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_emulated_kernels_match_oracle(name, mode):
    images, patches, tasks = cases.get(name)
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=mode)
    got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode, chunk_pixels=512)
    cases.assert_parity(ref, got, mode, name)


@pytest.mark.parametrize("name", ["config2_rotated_wcs", "small_field"])
def test_emulated_kernels_match_oracle_hessian(name):
    images, patches, tasks = cases.get(name)
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=2, n_threads=4)
    got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2, chunk_pixels=512)
    cases.assert_parity(ref, got, 2, name)


def test_chunking_does_not_change_counters_or_parity():
    images, patches, tasks = cases.get("two_body")
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=2)
    for chunk in (128, 200, 4096):
        got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2, chunk_pixels=chunk)
        cases.assert_parity(ref, got, 2, f"chunk={chunk}")


@pytest.mark.parametrize("name,active", [("two_body", [1, 2]), ("two_body", [2, 1]), ("masked", [1, 2]),
                                         ("clipped_and_empty", [4, 1, 2])])
def test_emulated_kernels_multiple_active_sources(name, active):
    """Sa > 1: dedupe of visited pixels, per-source passes and the pair kernel's cross-source Hessian blocks."""
    images, patches, tasks = cases.get(name)
    S = patches.shape[0]
    vp = [None] * S
    for rows, act, v in tasks:
        for j, r in enumerate(rows):
            vp[r - 1] = v[:, j]
    tk = [(list(range(1, S + 1)), active, np.stack(vp, axis=1))]
    for mode in (1, 2):
        ref = oracle_lib.OracleField(images, patches).elbo_batch(tk, mode=mode)
        got = emul_lib.EmulField(images, patches).elbo_batch(tk, mode=mode)
        cases.assert_parity(ref, got, mode, f"{name} {active}")


@pytest.mark.parametrize("name", ["two_body", "masked", "clipped_and_empty", "psf_k3", "config2_rotated_wcs", "small_field"])
def test_emulated_render_kernel_matches_oracle(name):
    """Row f.4: render_kernel + the host tile binning (fill_celeste_expectation!, bin/write_celeste_expectation.jl:
    111-156) against the oracle's per-pixel add_pixel_term! loop; subsets and reorderings of the source list too."""
    images, patches, tasks = cases.get(name)
    vp = cases.all_vp(patches, tasks)
    S = patches.shape[0]
    rows = np.arange(1, S + 1)
    ref = oracle_lib.oracle_render_expectation(images, patches, rows, vp)
    got = emul_lib.render_expectation(images, patches, rows, vp)
    cases.assert_render_parity(ref, got, name)
    assert any(np.abs(r).max() > 0 for r in ref)
    if S >= 2:
        sub = rows[::-1][: max(1, S // 2)]
        ref = oracle_lib.oracle_render_expectation(images, patches, sub, vp[:, sub - 1])
        got = emul_lib.render_expectation(images, patches, sub, vp[:, sub - 1])
        cases.assert_render_parity(ref, got, name + " subset")
    empty = emul_lib.render_expectation(images, patches, rows[:0], vp[:, :0])
    assert all(not e.any() for e in empty)
