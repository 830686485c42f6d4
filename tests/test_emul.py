"""The product's kernel SOURCE (celeste.jl_b200/csrc/celeste_kernels.cuh), compiled for the host under
tests/host_emul's emulation of the CUDA execution model, against the oracle -- CPU only.  This is not
the product path (the product only runs these kernels on a GPU); it catches indexing / reduction /
chain-rule bugs before GPU time is spent.  The real parity tests are tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cases
import emul_lib
import oracle_lib

FAST = ["star_1band", "two_body", "masked", "clipped_and_empty", "psf_k1", "psf_k3", "crowded", "sharp_psf"]


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


@pytest.mark.parametrize("name", ["two_body", "masked", "clipped_and_empty", "crowded", "config2_rotated_wcs", "small_field",
                                  "wide_patch", "seven_images", "sharp_psf"])
def test_march_kernel_matches_oracle_and_task_kernel(name):
    """march_kernels.cuh (row walks with the exp recurrence; the product's value / gradient path for Sa = 1, K = 2)
    against the oracle at the 1e-8 parity statement, and against task_kernel (direct evaluation of every pixel) at
    1e-11: the recurrence is restarted exactly every <= 16 pixels, so the two differ by accumulated rounding only.
    A walk is carried by a pair of lanes (one PSF component each) that exchange half of their sums by shuffles."""
    images, patches, tasks = cases.get(name)
    lib = emul_lib.load()
    for mode in (0, 1):
        ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=mode, n_threads=4)
        try:
            lib.emul_set_grad_kernel(0)
            direct = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
            lib.emul_set_grad_kernel(1)
            march = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
            lib.emul_set_grad_kernel(2)          # every source split into one block per image
            split = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
        finally:
            lib.emul_set_grad_kernel(3)
        cases.assert_parity(ref, march, mode, name)
        cases.assert_parity(ref, split, mode, name + " (one block per image)")
        cases.assert_parity(ref, direct, mode, name)
        assert np.array_equal(march["counters"], direct["counters"])
        fin = np.isfinite(direct["v"])
        assert np.all(np.abs(march["v"] - direct["v"])[fin] <= 1e-11 * np.abs(direct["v"])[fin])
        if mode == 1:
            n = len(tasks)
            a, b = march["d"].reshape(n, -1), direct["d"].reshape(n, -1)
            sc = np.abs(b).max(axis=1, keepdims=True)
            assert np.all(np.abs(a - b) <= 1e-11 * np.maximum(np.abs(b), sc * 1e-3)), np.abs(a - b).max()


@pytest.mark.parametrize("idx", [0, 2, 3, 5, 10, 26])
def test_march_kernel_propagates_nonfinite_parameters(idx):
    """A NaN anywhere in the active source's parameters must come out as a flagged, non-finite result (the reference
    throws in assert_all_finite, elbo_args.jl:145-149) -- in particular it must not be swallowed by the rule that puts
    underflowed components to sleep."""
    images, patches, tasks = cases.get("two_body")
    bad = [(r, a, v.copy()) for r, a, v in tasks]
    bad[0][2][idx, 0] = np.nan
    lib = emul_lib.load()
    try:
        for which in (1, 3):                      # march_kernel, unit kernels
            lib.emul_set_grad_kernel(which)
            for mode in (0, 1):
                ref = oracle_lib.OracleField(images, patches).elbo_batch(bad, mode=mode)
                got = emul_lib.EmulField(images, patches).elbo_batch(bad, mode=mode)
                assert ref["flags"].tolist() == [1, 0] and got["flags"].tolist() == [1, 0]
                assert not np.isfinite(got["v"][0]) and np.isfinite(got["v"][1])
    finally:
        lib.emul_set_grad_kernel(3)


@pytest.mark.parametrize("seed", [0, 3, 7, 12, 19, 23, 31, 38])
def test_march_kernel_at_the_corners_of_the_parameter_box(seed):
    """Random small scenes with extreme source parameters (cases.random_extreme_scene): the recurrence of march_kernel
    against the oracle at the parity statement, value and gradient."""
    images, patches, tasks = cases.random_extreme_scene(seed)
    lib = emul_lib.load()
    try:
        for which in (1, 3):                      # march_kernel, unit kernels
            lib.emul_set_grad_kernel(which)
            for mode in (0, 1):
                ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=mode, n_threads=4)
                got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
                cases.assert_parity(ref, got, mode, f"seed {seed} kernel {which}")
    finally:
        lib.emul_set_grad_kernel(3)


def test_chunking_does_not_change_counters_or_parity():
    images, patches, tasks = cases.get("two_body")
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=2)
    lib = emul_lib.load()
    try:
        lib.emul_set_hess_kernel(0)              # pixel_kernel<2> (the chunked kernel)
        for chunk in (128, 200, 4096):
            got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2, chunk_pixels=chunk)
            cases.assert_parity(ref, got, 2, f"chunk={chunk}")
    finally:
        lib.emul_set_hess_kernel(1)


def _tight(ref, got, mode, label, rtol):
    rv, gv = ref["v"], got["v"]
    fin = np.isfinite(rv)
    assert np.all(np.abs(rv - gv)[fin] <= rtol * np.abs(rv)[fin]), (label, np.abs(rv - gv).max())
    n = len(rv)
    for key, need in (("d", 1), ("h", 2)):
        if mode >= need:
            r, g = ref[key].reshape(n, -1), got[key].reshape(n, -1)
            sc = np.abs(r).max(axis=1, keepdims=True)
            err = np.abs(r - g) / np.maximum(np.maximum(np.abs(r), sc * 1e-3), 1e-300)
            assert np.nanmax(err) <= rtol, (label, key, np.nanmax(err))


@pytest.mark.parametrize("name", ["star_1band", "galaxy", "two_body", "masked", "clipped_and_empty", "crowded",
                                  "config2_rotated_wcs", "small_field", "wide_patch", "seven_images", "sharp_psf"])
def test_unit_kernel_matches_oracle_and_pixel_kernel(name):
    """unit_kernels.cuh -- the Hessian of the production shape as (phase A) row walks with the first-order mixture sums
    + (phase B) L5-weighted moments of every component folded by closed forms -- against the oracle's dense per-pixel
    chain rule at the parity statement AND at 1e-10 (floor 1e-3 of the row scale), and against pixel_kernel<2>
    (direct evaluation): identical pixel-visit counters.  Modes 0 / 1 of the same kernel too."""
    images, patches, tasks = cases.get(name)
    lib = emul_lib.load()
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=2, n_threads=4)
    unit = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2)
    try:
        lib.emul_set_hess_kernel(0)
        direct = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2)
    finally:
        lib.emul_set_hess_kernel(1)
    cases.assert_parity(ref, unit, 2, name + " unit_kernel")
    cases.assert_parity(ref, direct, 2, name + " pixel_kernel")
    _tight(ref, unit, 2, name, 1e-10)
    assert np.array_equal(unit["counters"], direct["counters"])
    assert not np.array_equal(unit["h"], direct["h"]) or not unit["h"].any(), "the switch did not change kernels"
    for mode in (0, 1):
        r = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=mode, n_threads=4)
        g = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
        cases.assert_parity(r, g, mode, name + f" unit_kernel mode {mode}")
        _tight(r, g, mode, name, 1e-11)


@pytest.mark.parametrize("name", ["two_body", "masked", "clipped_and_empty", "crowded", "wide_patch", "sharp_psf"])
@pytest.mark.parametrize("rows", [0, 3])
def test_unit_kernel_row_cuts(name, rows):
    """Every (source, image) is cut into units of at most CELESTE_UNIT_ROWS rows (one warp each; the partial vectors
    of the pieces meet in the epilogue).  Other cuts -- none at all, 3 rows per unit -- give the same parity and the
    same counters in every mode."""
    images, patches, tasks = cases.get(name)
    lib = emul_lib.load()
    try:
        lib.emul_set_unit_target(rows)
        for mode in (0, 1, 2):
            ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=mode, n_threads=4)
            got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=mode)
            cases.assert_parity(ref, got, mode, f"{name} mode {mode}, {rows} rows per unit")
            _tight(ref, got, mode, name, 1e-10)
    finally:
        lib.emul_set_unit_target(16)


@pytest.mark.parametrize("idx", [0, 2, 3, 5, 10, 26])
def test_unit_kernel_propagates_nonfinite_parameters(idx):
    images, patches, tasks = cases.get("two_body")
    bad = [(r, a, v.copy()) for r, a, v in tasks]
    bad[0][2][idx, 0] = np.nan
    ref = oracle_lib.OracleField(images, patches).elbo_batch(bad, mode=2)
    got = emul_lib.EmulField(images, patches).elbo_batch(bad, mode=2)
    assert ref["flags"].tolist() == [1, 0] and got["flags"].tolist() == [1, 0]
    assert np.isfinite(got["v"][1]) and np.isfinite(got["h"].reshape(2, -1)[1]).all()


@pytest.mark.parametrize("seed", [0, 3, 7, 12, 19, 23, 31, 38])
def test_unit_kernel_at_the_corners_of_the_parameter_box(seed):
    """cases.random_extreme_scene through unit_kernel<2>: the moment formulation at radius 0.02 .. 60 px, axis ratio
    0.02, sources far off their patch centres."""
    images, patches, tasks = cases.random_extreme_scene(seed)
    ref = oracle_lib.OracleField(images, patches).elbo_batch(tasks, mode=2, n_threads=4)
    got = emul_lib.EmulField(images, patches).elbo_batch(tasks, mode=2)
    cases.assert_parity(ref, got, 2, f"seed {seed}")


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


def test_emulated_patch_construction_kernels():
    """Row f.4, ImagePatch construction on the device (csrc/patch_kernels.cuh) against the host model
    (model.py: render_psf / psf_spline_coefs / ImagePatch / find_all_neighbors, imaged_sources.jl:80-117,232-244)."""
    from celeste_jl_b200 import model
    from celeste_jl_b200 import synthetic
    rng = np.random.default_rng(3)
    # spline of a raw stamp, incl. negative entries (max(., 0)) and a non-51 size
    for n in (51, 21, 4):
        raw = rng.normal(0.2, 1.0, (n, n)) ** 2 * np.exp(-0.1 * ((np.arange(n)[:, None] - n / 2) ** 2 + (np.arange(n)[None, :] - n / 2) ** 2))
        raw[rng.random((n, n)) < 0.05] = -0.3
        want = model.psf_spline_coefs(raw)
        got = emul_lib.spline_build(n, raw=raw)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    # spline from the mixture itself (render_psf on the device), K = 1, 2, 3
    images, patches, _ = cases.get("psf_k3")
    for psf in (images[0].psf, cases.get("two_body")[0][0].psf, cases.get("psf_k1")[0][0].psf):
        want = model.psf_spline_coefs(model.render_psf(psf, (51, 51)))
        got = emul_lib.spline_build(51, psf=psf)
        assert np.allclose(got, want, rtol=1e-11, atol=1e-12 * np.abs(want).max())
    # the interpolant reproduces the (transformed) grid at the grid points (SURVEY 8c must-hold self-check)
    stamp = model.render_psf(images[0].psf, (51, 51))
    g = np.maximum(stamp, 0) + 1e-6
    g = model.softpluslike(g / g.sum())
    co = emul_lib.spline_build(51, raw=stamp)
    rec = (co[:-2, :-2] + co[2:, :-2] + co[:-2, 2:] + co[2:, 2:]) / 36 + (co[1:-1, :-2] + co[1:-1, 2:] + co[:-2, 1:-1] + co[2:, 1:-1]) / 9 \
        + co[1:-1, 1:-1] * 4 / 9
    assert np.allclose(rec, g, rtol=1e-11, atol=1e-11)
    # bitmaps
    nan_seen = False
    images, patches, _ = cases.get("masked")
    for s in range(patches.shape[0]):
        for n in range(patches.shape[1]):
            p = patches[s, n]
            H2, W2 = p.active_pixel_bitmap.shape
            o = p.bitmap_offset
            got = emul_lib.bitmap_build(images[n].pixels, int(o[0]), int(o[1]), H2, W2)
            want = ~np.isnan(images[n].pixels[o[0]:o[0] + H2, o[1]:o[1] + W2])      # imaged_sources.jl:94-95
            assert np.array_equal(got, want)
            nan_seen = nan_seen or (~want).any()
    assert nan_seen
    # neighbours
    for name in ("clipped_and_empty", "crowded", "small_field"):
        _, patches, _ = cases.get(name)
        assert emul_lib.find_all_neighbors(patches) == model.find_all_neighbors(patches)


def test_emulated_box_render_reproduces_the_synthetic_image_generator():
    """Synthetic.gen_image! (Synthetic.jl:30-47) through the render kernel: celeste_render_boxes at the catalog's own
    fluxes (synthetic.catalog_truth_vp) + sky, x iota == the host renderer's expectation image (synthetic.gen_images,
    Float32 pixels), including bodies whose radius-25 box is clipped by the image border and the last column of every
    box (which the ELBO's strict w2 < W2 rule would leave out)."""
    from celeste_jl_b200 import synthetic
    from celeste_jl_b200.model import ImagePatch, box_around_point
    images = synthetic.blank_images(70, 64, bands=(1, 3, 5))
    catalog = [synthetic.sample_ce([30.3, 28.8], False), synthetic.sample_ce([8.1, 60.2], True),
               synthetic.sample_ce([66.0, 5.5], False), synthetic.sample_ce([40.7, 41.9], True)]
    catalog[2].gal_axis_ratio, catalog[2].gal_angle, catalog[2].gal_frac_dev = 0.4, 1.1, 0.8
    ref = [im for im in synthetic.blank_images(70, 64, bands=(1, 3, 5))]
    synthetic.gen_images(ref, catalog, expectation=True, device="cpu")
    patches = np.empty((len(catalog), len(images)), dtype=object)
    for n, img in enumerate(images):
        for s, ce in enumerate(catalog):
            patches[s, n] = ImagePatch(img, box_around_point(img.wcs, ce.pos, 25))
    vp = np.stack([synthetic.catalog_truth_vp(ce) for ce in catalog], axis=1)
    add = emul_lib.render_expectation(images, patches, np.arange(1, len(catalog) + 1), vp, full_box=True)
    strict = emul_lib.render_expectation(images, patches, np.arange(1, len(catalog) + 1), vp)
    for img, r, a, st in zip(images, ref, add, strict):
        lam = (a + np.asarray(img.sky, dtype=np.float64)) * np.asarray(img.nelec_per_nmgy, dtype=np.float64)[:, None]
        assert np.allclose(lam, r.pixels.astype(np.float64), rtol=2e-6, atol=1e-4), np.abs(lam - r.pixels).max()
        assert (a != st).any() and (a >= st - 1e-12).all()       # the strict rule leaves each box's last column out
