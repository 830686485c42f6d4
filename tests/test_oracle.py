"""Pins the ORACLE (oracle/celeste_oracle.cpp) -- CPU only.

The reference holds no golden numbers for this path.  Its own tests pin the hot
path by (i) one closed form, (ii) structural properties, (iii) hand derivatives == AD of the same
value code.  These tests restate exactly those, with torch.float64 autograd (tests/ad_model.py, an
independent value-only model) in the role of ForwardDiff.
"""
import math

import numpy as np
import pytest

import ad_model
import cases
import oracle_lib
from celeste_jl_b200 import synthetic
from celeste_jl_b200.model import cubic_bspline_prefilter, galaxy_prototypes, ids, psf_spline_coefs, render_psf, softpluslike
import ctypes as C


def test_bvn_cov_closed_form():
    """test/test_elbo.jl:45-61."""
    ab, angle, scale = .7, math.pi / 5, 2.
    out = np.zeros(3)
    oracle_lib.load().oracle_get_bvn_cov(ab, angle, scale, out.ctypes.data)
    assert out[0] == pytest.approx(scale ** 2 * (1 + (ab ** 2 - 1) * math.sin(angle) ** 2), rel=1e-14)
    assert out[1] == pytest.approx(scale ** 2 * (1 - ab ** 2) * math.cos(angle) * math.sin(angle), rel=1e-14)
    assert out[2] == pytest.approx(scale ** 2 * (1 + (ab ** 2 - 1) * math.cos(angle) ** 2), rel=1e-14)


def test_galaxy_prototypes():
    """light_source_model.jl:45-72: amplitudes normalised; variances rescaled by the effective radii."""
    eta, nu = np.zeros(16), np.zeros(16)
    oracle_lib.load().oracle_galaxy_prototypes(eta.ctypes.data, nu.ctypes.data)
    assert eta[:8].sum() == pytest.approx(1.0, abs=1e-15) and eta[8:14].sum() == pytest.approx(1.0, abs=1e-15)
    (de, dn), (ee, en) = galaxy_prototypes
    assert np.allclose(eta[:8], de, rtol=0, atol=1e-16) and np.allclose(nu[8:14], en, rtol=1e-15)
    assert nu[7] == pytest.approx(8.40215071 / 1.078031 ** 2, rel=1e-15)


def _spline(coefs, x, y):
    out = np.zeros(6)
    oracle_lib.load().oracle_spline_eval(coefs.ctypes.data, coefs.shape[0], coefs.shape[1], x, y, out.ctypes.data)
    return out


def test_spline_interpolates_grid_and_derivatives():
    """Must-hold self-check of the restated Interpolations.jl rule: the interpolant of
    prefilter(grid) reproduces grid at integer coordinates; derivatives are the exact polynomial ones."""
    stamp = render_psf(synthetic.make_two_component_psf(), (51, 51))
    g = np.maximum(stamp, 0) + 1e-6
    g = softpluslike(g / g.sum())
    coefs = np.asfortranarray(psf_spline_coefs(stamp))
    assert coefs.shape == (53, 53)
    for (i, j) in [(1, 1), (26, 26), (51, 51), (7, 40), (30, 2)]:
        assert _spline(coefs, float(i), float(j))[0] == pytest.approx(g[i - 1, j - 1], rel=1e-10, abs=1e-10)
    rng = np.random.default_rng(0)
    for _ in range(20):
        x, y = rng.uniform(0.6, 51.4, 2)     # includes the clamp-then-extrapolate rim
        v = _spline(coefs, x, y)
        e = 1e-5
        fx = (_spline(coefs, x + e, y)[0] - _spline(coefs, x - e, y)[0]) / (2 * e)
        fy = (_spline(coefs, x, y + e)[0] - _spline(coefs, x, y - e)[0]) / (2 * e)
        fxx = (_spline(coefs, x + e, y)[1] - _spline(coefs, x - e, y)[1]) / (2 * e)
        fxy = (_spline(coefs, x, y + e)[1] - _spline(coefs, x, y - e)[1]) / (2 * e)
        fyy = (_spline(coefs, x, y + e)[2] - _spline(coefs, x, y - e)[2]) / (2 * e)
        scale = max(1.0, abs(v[0]))
        assert abs(v[1] - fx) < 1e-6 * scale and abs(v[2] - fy) < 1e-6 * scale
        assert abs(v[3] - fxx) < 1e-5 * scale and abs(v[4] - fxy) < 1e-5 * scale and abs(v[5] - fyy) < 1e-5 * scale


def _check_vs_ad(images, patches, vp, active, hessian):
    v, d, h, _ = oracle_lib.oracle_elbo(images, patches, vp, active, mode=2 if hessian else 1)
    va, da, ha = ad_model.elbo_ad(images, patches, vp, active, hessian=hessian)
    assert abs(v - va) <= 1e-12 * abs(va)
    assert np.all(np.abs(d - da) <= 1e-9 * np.maximum(np.abs(da), np.abs(da).max() * 1e-6))
    if hessian:
        assert np.all(np.abs(h - ha) <= 1e-9 * np.maximum(np.abs(ha), np.abs(ha).max() * 1e-6))
        assert np.array_equal(h, h.T)
        k0 = 28
        assert not h[k0:44, :].any() and not d[k0:44].any()        # ids.k never gets likelihood derivatives


def test_manual_gradient_and_hessian_match_autodiff_star():
    """test/test_elbo.jl:223-270 on the 1-star fixture: every gradient and Hessian entry."""
    images, patches, vp, _ = synthetic.gen_sample_star_dataset(bands=(3,), H=20, W=20)
    _check_vs_ad(images, patches, vp, [1], hessian=True)


def test_manual_gradient_and_hessian_match_autodiff_two_body():
    """test/test_elbo.jl:223-270 on the two-body fixture (active galaxy-like source with a neighbour)."""
    images, patches, vp, _ = synthetic.gen_two_body_dataset()
    _check_vs_ad(images, patches, vp, [2], hessian=True)


def test_manual_gradient_matches_autodiff_rotated_wcs_two_active():
    """Gradient for Sa = 2 under a non-identity wcs_jacobian (the -J chain of transform_bvn_ux_derivs!)."""
    images, patches, vp, _ = synthetic.gen_config2_dataset(rotated_wcs=True)
    _check_vs_ad(images, patches, vp, [3, 1], hessian=False)


def _flat(images, patches):
    from celeste_jl_b200.flatten import FlatImages, FlatPatches
    return FlatImages(images), FlatPatches(patches)


def test_calculate_G_s_overwrites():
    """test/test_elbo.jl:13-42: calculate_G_s! overwrites E_G_s / E_G2_s / var_G_s instead of accumulating."""
    images, patches, vp, _ = synthetic.gen_two_body_dataset()
    fi, fp = _flat(images, patches)
    vpm = np.asfortranarray(np.stack(vp, axis=1)).ravel(order="F")
    act = np.array([1, 2], dtype=np.int32)
    n_out = 3 * (1 + 44 + 44 * 44)

    def probe(calls):
        s = np.array([c[0] for c in calls], dtype=np.int32)
        b = np.array([c[1] for c in calls], dtype=np.int32)
        out = np.zeros(n_out)
        st = oracle_lib.load().oracle_calculate_G_s_probe(fi.N, C.addressof(fi.arr), fp.S_tot, C.addressof(fp.arr), 2,
                                                          act.ctypes.data, vpm.ctypes.data, len(calls),
                                                          s.ctypes.data, b.ctypes.data, out.ctypes.data)
        assert st == 0
        return out
    cleared = probe([(2, 4)])
    dirty = probe([(2, 4), (1, 2), (2, 3), (2, 4)])
    assert np.array_equal(cleared, dirty)
    assert np.abs(cleared).max() > 0


def test_active_sources_partition():
    """test/test_elbo.jl:64-130."""
    images, patches, vp, _ = synthetic.gen_two_body_dataset()
    P = 44
    for n in range(5):
        p = patches[0, n]
        assert tuple(p.bitmap_offset) == (0, 0) and p.active_pixel_bitmap.shape == images[n].pixels.shape
        assert p.active_pixel_bitmap.all()
    for n in range(5):
        patches[1, n].active_pixel_bitmap[:] = False
    _, d, _, _ = oracle_lib.oracle_elbo(images, patches, vp, [1, 2])
    assert not d[:, 1].any()
    patches[1, 4].active_pixel_bitmap[9:11, 9:11] = True
    v12, d12, h12, _ = oracle_lib.oracle_elbo(images, patches, vp, [1, 2])
    v21, d21, h21, _ = oracle_lib.oracle_elbo(images, patches, vp, [2, 1])
    assert v12 == pytest.approx(v21, rel=1e-13)
    assert np.allclose(d12[:, 0], d21[:, 1], rtol=1e-10) and np.allclose(d12[:, 1], d21[:, 0], rtol=1e-10)
    v1, d1, h1, _ = oracle_lib.oracle_elbo(images, patches, vp, [1])
    assert v1 == pytest.approx(v12, rel=1e-13)
    v2, d2, h2, _ = oracle_lib.oracle_elbo(images, patches, vp, [2])
    assert np.allclose(d12[:, 0], d1[:, 0], rtol=1e-10) and np.allclose(d12[:, 1], d2[:, 0], rtol=1e-10)
    assert np.allclose(h12[:P, :P], h1, rtol=1e-10) and np.allclose(h12[P:, P:], h2, rtol=1e-10)


def _val(images, patches, vp):
    return oracle_lib.oracle_elbo(images, patches, vp, [1], mode=0)[0]


def test_star_truth_is_most_likely():
    """test/test_elbo.jl:132-170 with true_star_init (SampleData.jl:264-274)."""
    images, patches, vp, _ = synthetic.gen_sample_star_dataset(perturb=False)
    vp[0][ids.is_star] = [1.0 - 1e-4, 1e-4]
    vp[0][ids.flux_scale] = 1e-4
    vp[0][ids.flux_loc] = math.log(synthetic.sample_star_fluxes[2]) - 0.5 * vp[0][ids.flux_scale]
    vp[0][ids.color_var] = 1e-4
    best = _val(images, patches, vp)
    for bad_a in (.3, .5, .9):
        q = [v.copy() for v in vp]
        q[0][ids.is_star] = [1.0 - bad_a, bad_a]
        assert best > _val(images, patches, q)
    for h2 in range(-2, 3):
        for w2 in range(-2, 3):
            if h2 or w2:
                q = [v.copy() for v in vp]
                q[0][ids.pos] += [h2 * .5, w2 * .5]
                assert best > _val(images, patches, q)
    for delta in (.7, .9, 1.1, 1.3):
        q = [v.copy() for v in vp]
        q[0][ids.flux_loc] += math.log(delta)
        assert best > _val(images, patches, q)
    for b in range(4):
        for delta in (-.3, .3):
            q = [v.copy() for v in vp]
            q[0][ids.color_mean[b, 0]] += delta
            assert best > _val(images, patches, q)


def test_galaxy_truth_is_most_likely():
    """test/test_elbo.jl:173-220."""
    images, patches, vp, _ = synthetic.gen_sample_galaxy_dataset(perturb=False)
    vp[0][ids.is_star] = [0.01, .99]
    best = _val(images, patches, vp)
    for bad_a in (.3, .5, .9):
        q = [v.copy() for v in vp]
        q[0][ids.is_star] = [1.0 - bad_a, bad_a]
        assert best > _val(images, patches, q)
    for h2 in range(-2, 3):
        for w2 in range(-2, 3):
            if h2 or w2:
                q = [v.copy() for v in vp]
                q[0][ids.pos] += [h2 * .5, w2 * .5]
                assert best > _val(images, patches, q)
    for bad_scale in (.8, 1.2):
        q = [v.copy() for v in vp]
        q[0][ids.flux_loc] += 2 * math.log(bad_scale)
        assert best > _val(images, patches, q)
    for name in ("gal_axis_ratio", "gal_angle", "gal_radius_px"):
        for bad_scale in (.8, 1.2):
            q = [v.copy() for v in vp]
            q[0][getattr(ids, name)] *= bad_scale
            assert best > _val(images, patches, q)
    for b in range(4):
        for delta in (-.3, .3):
            q = [v.copy() for v in vp]
            q[0][ids.color_mean[b, 1]] += delta
            assert best > _val(images, patches, q)


def test_hessian_vector_product_matches_finite_difference():
    """test/test_elbo.jl:273-301 (1 % tolerance on the first 20 entries)."""
    images, patches, vp, _ = synthetic.gen_two_body_dataset()
    _, d0, h, _ = oracle_lib.oracle_elbo(images, patches, vp, [1, 2])
    eps = 1e-5
    vp1 = [v + eps for v in vp]
    _, d1, _, _ = oracle_lib.oracle_elbo(images, patches, vp1, [1, 2], mode=1)
    hv_fd = (d1 - d0).ravel(order="F") / eps
    hv = h @ np.ones(88)
    for i in range(20):
        assert abs(hv_fd[i] - hv[i]) <= 0.01 * abs(hv[i])


def test_oracle_threads_agree():
    images, patches, tasks = cases.get("small_field")
    of = oracle_lib.OracleField(images, patches)
    a = of.elbo_batch(tasks, mode=2, n_threads=1)
    b = of.elbo_batch(tasks, mode=2, n_threads=4)
    for k in ("v", "d", "h", "counters", "flags"):
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("name", ["two_body", "clipped_and_empty", "config2_rotated_wcs", "psf_k3"])
def test_oracle_render_matches_independent_restatement(name):
    """Row f.4 checker: the oracle's fill_celeste_expectation! (add_pixel_term! on every pixel) against the torch
    restatement of the model (tests/ad_model.py) -- and it is additive over sources."""
    import ad_model
    images, patches, tasks = cases.get(name)
    vp = cases.all_vp(patches, tasks)
    S = patches.shape[0]
    rows = np.arange(1, S + 1)
    got = oracle_lib.oracle_render_expectation(images, patches, rows, vp, n_threads=4)
    ref = ad_model.render_value(images, patches, [vp[:, s] for s in range(S)])
    cases.assert_render_parity(ref, got, name, rtol=1e-11)
    parts = [oracle_lib.oracle_render_expectation(images, patches, rows[s:s + 1], vp[:, s:s + 1]) for s in range(S)]
    for n in range(len(images)):
        total = sum(p[n] for p in parts)
        assert np.allclose(total, got[n], rtol=1e-12, atol=1e-12 * max(np.abs(got[n]).max(), 1.0))
