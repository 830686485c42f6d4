"""Named parity cases shared by the CPU (emulation), GPU and golden-vector tests.

Each case returns (images, patches, tasks) with tasks = [(rows_1based, active_local_1based, vp 44 x S)].
They map onto BASELINE.json `configs` and the edge cases the reference tests exercise
(test/test_elbo.jl, test/test_images.jl): masked/NaN pixels, empty and clipped patches, neighbours,
non-identity WCS, K != 2.
"""
import math

import numpy as np

import celeste_jl_b200 as cj
from celeste_jl_b200 import synthetic
from celeste_jl_b200.model import AffineWCS, ImagePatch, PsfComponent, get_sky_patches


def _vpm(vp, rows):
    return np.stack([vp[r - 1] for r in rows], axis=1)


def _all_tasks(vp, neighbors=None):
    S = len(vp)
    tasks = []
    for t in range(S):
        nb = [s for s in range(S) if s != t] if neighbors is None else neighbors[t]
        rows = [t + 1] + [s + 1 for s in nb]
        tasks.append((rows, [1], _vpm(vp, rows)))
    return tasks


def case_star_1band():
    """configs[0]: SampleData.gen_sample_star_dataset, 1 star, 1 band (r), 20 x 20 tile."""
    images, patches, vp, _ = synthetic.gen_sample_star_dataset(bands=(3,), H=20, W=20)
    return images, patches, _all_tasks(vp)


def case_star_5band():
    """test/SampleData.jl:161-173: the 20 x 23, 5-band star fixture of test_elbo.jl."""
    images, patches, vp, _ = synthetic.gen_sample_star_dataset()
    return images, patches, _all_tasks(vp)


def case_galaxy():
    images, patches, vp, _ = synthetic.gen_sample_galaxy_dataset()
    return images, patches, _all_tasks(vp)


def case_two_body():
    """test/SampleData.jl:193-208 (two overlapping sources; each is the other's neighbour)."""
    images, patches, vp, _ = synthetic.gen_two_body_dataset()
    return images, patches, _all_tasks(vp)


def case_config2():
    """configs[1]: 3 sources (2 stars + 1 galaxy), 5 SDSS bands, 50 x 50 tiles."""
    images, patches, vp, _ = synthetic.gen_config2_dataset()
    return images, patches, _all_tasks(vp)


def case_config2_rotated_wcs():
    """configs[1] with a rotated / anisotropic wcs_jacobian (second WCS variant)."""
    images, patches, vp, _ = synthetic.gen_config2_dataset(rotated_wcs=True)
    return images, patches, _all_tasks(vp)


def case_masked():
    """NaN pixels (masked in SDSS frames, imaged_sources.jl:94-95 + elbo_objective.jl:459) and a
    hand-edited bitmap whose NaN pixel stays 'active' (the second check must catch it)."""
    images, patches, vp, catalog = synthetic.gen_two_body_dataset()
    rng = np.random.default_rng(5)
    for img in images:
        m = rng.random(img.pixels.shape) < 0.08
        img.pixels[m] = np.nan
    patches = get_sky_patches(images, catalog)
    # re-activate some NaN pixels in the bitmap of source 1, image 3
    bm = patches[0, 2].active_pixel_bitmap
    idx = np.argwhere(~bm)
    for h2, w2 in idx[:5]:
        bm[h2, w2] = True
    # and knock out a block of good pixels
    patches[1, 4].active_pixel_bitmap[3:6, 2:9] = False
    return images, patches, _all_tasks(vp)


def case_clipped_and_empty():
    """Patches clipped by the image border and one source whose patch is entirely off one image
    (empty 1:0 box, imaged_sources.jl:81-84); plus a 1-column patch (only the strict last column)."""
    images = synthetic.blank_images(40, 36)
    catalog = [synthetic.sample_ce([3.2, 4.1], False), synthetic.sample_ce([38.6, 33.9], True),
               synthetic.sample_ce([20.3, 60.0], True), synthetic.sample_ce([19.0, 17.5], False)]
    synthetic.gen_images(images, catalog, seed=3, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=9.0)
    # source 3 is off-image in w: its box clamps to empty
    assert patches[2, 0].active_pixel_bitmap.size == 0
    # a degenerate single-column patch for source 4 in image 2
    patches[3, 1] = ImagePatch(images[1], ((10, 28), (17, 17)))
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    return images, patches, _all_tasks(vp)


def _case_psf(psf, bands, seed):
    from celeste_jl_b200.model import render_psf
    images = synthetic.blank_images(30, 30, bands=bands)
    for im in images:
        im.psf = psf
        im.psf_stamp = render_psf(im.psf, (51, 51))
        im._coefs_cache = None
    catalog = [synthetic.sample_ce([14.2, 15.7], False), synthetic.sample_ce([18.9, 11.3], True)]
    synthetic.gen_images(images, catalog, seed=seed, device="cpu")
    patches = get_sky_patches(images, catalog)
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    return images, patches, _all_tasks(vp)


def case_psf_k1():
    """K = 1 PSF mixture (psf_K is a free parameter of ElboArgs, elbo_args.jl:197)."""
    return _case_psf([PsfComponent(1.0, np.array([0.1, -0.05]), np.array([[2.0, 0.1], [0.1, 1.8]]))], (2, 4), 4)


def case_psf_k3():
    """K = 3 PSF mixture."""
    return _case_psf([PsfComponent(0.6, np.array([0.0, 0.0]), np.array([[1.5, 0.0], [0.0, 1.5]])),
                      PsfComponent(0.3, np.array([0.1, 0.1]), np.array([[4.0, 0.3], [0.3, 5.0]])),
                      PsfComponent(0.1, np.array([-0.2, 0.0]), np.array([[12.0, -1.0], [-1.0, 10.0]]))], (1, 5), 7)


def case_crowded(n_sources=90, seed=11):
    """> 64 overlapping neighbours for one target (ParallelRun.jl:475 warns above 100): exercises the
    neighbour-list overflow path of the pixel kernel."""
    images = synthetic.blank_images(48, 48, bands=(3,))
    rng = np.random.default_rng(seed)
    catalog = []
    for k in range(n_sources):
        pos = [24 + rng.normal(0, 6), 24 + rng.normal(0, 6)]
        catalog.append(synthetic.sample_ce(pos, bool(k % 3)))
        catalog[-1].star_fluxes = catalog[-1].star_fluxes * 0.02
        catalog[-1].gal_fluxes = catalog[-1].gal_fluxes * 0.02
    synthetic.gen_images(images, catalog, seed=6, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=8.0)
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    tasks = _all_tasks(vp)
    return images, patches, [tasks[0], tasks[37]]


def case_small_field(n_sources=40, H=160, W=140, seed=42):
    """A miniature of configs[2]: prior-drawn catalog, catalog-sized patches, find_neighbors tasks."""
    ds = synthetic.FieldDataset(n_sources, H=H, W=W, seed=seed, device="cpu")
    rows, act = ds.tasks()
    tasks = [(r, a, _vpm(ds.vp, r)) for r, a in zip(rows, act)]
    return ds.images, ds.patches, tasks


def case_wide_patch():
    """Patches wider than the 51 x 51 PSF stamp (un-capped SEP boxes, detection.jl:153-158): the spline is evaluated
    in its clamped edge cells (fsm_util.jl:236 does no bounds handling), rows need several exact restarts of the
    march kernel's walks, and the neighbour overlaps only part of the rows."""
    images = synthetic.blank_images(90, 84, bands=(2, 4))
    catalog = [synthetic.sample_ce([44.3, 40.8], True), synthetic.sample_ce([61.9, 57.2], False)]
    synthetic.gen_images(images, catalog, seed=9, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=38.0)
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    return images, patches, _all_tasks(vp)


def case_sharp_psf():
    """A PSF with a sub-pixel core (variance 0.2 px^2) on patches of radius 30: at the patch edge the Gaussian
    components of the core are below the smallest double (exp(-q/2), q > 1490).  Direct evaluation does not care; a
    recurrence that starts a row there must not carry the underflowed value towards the centre."""
    psf = [PsfComponent(0.7, np.array([0.05, -0.1]), np.array([[0.2, 0.02], [0.02, 0.25]])),
           PsfComponent(0.3, np.array([0.0, 0.0]), np.array([[3.0, 0.0], [0.0, 3.2]]))]
    from celeste_jl_b200.model import render_psf
    images = synthetic.blank_images(72, 70, bands=(2, 3))
    for im in images:
        im.psf = psf
        im.psf_stamp = render_psf(im.psf, (51, 51))
        im._coefs_cache = None
    catalog = [synthetic.sample_ce([36.4, 34.7], True), synthetic.sample_ce([41.2, 30.1], False)]
    synthetic.gen_images(images, catalog, seed=21, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=30.0)
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    for v in vp:
        v[ids_radius()] = 0.05          # a nearly unresolved galaxy: its narrowest components are the PSF core itself
    return images, patches, _all_tasks(vp)


def ids_radius():
    from celeste_jl_b200.model import ids
    return ids.gal_radius_px - 1 if isinstance(ids.gal_radius_px, int) else ids.gal_radius_px[0] - 1


def case_seven_images():
    """N = 7 images (two exposures of bands 3 and 4 on top of the five SDSS bands, as a multi-epoch box has): the
    march kernel splits a source into two image groups whose partial sums meet in the epilogue, and two images share
    a band's accumulators."""
    images = synthetic.blank_images(44, 40, bands=(1, 2, 3, 4, 5, 3, 4))
    catalog = [synthetic.sample_ce([20.7, 18.2], False), synthetic.sample_ce([25.1, 23.3], True),
               synthetic.sample_ce([12.4, 27.9], True)]
    synthetic.gen_images(images, catalog, seed=13, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=10.0)
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    return images, patches, _all_tasks(vp)


def random_extreme_scene(seed):
    """A small random scene whose sources sit at the corners of the parameter box (ElboMaximize.jl:63-93) and well off
    their patch centres: radius 0.02 .. 60 px, axis ratio down to 0.02, gal_frac_dev and is_star at 0 / 1, patches of
    radius 4 .. 28 px, 1-2 bands, 1-3 sources that are each other's neighbours.  Not in CASES (stress tests draw it)."""
    from celeste_jl_b200.model import ids
    rng = np.random.default_rng(seed)
    H, W = int(rng.integers(30, 60)), int(rng.integers(30, 60))
    bands = tuple(sorted(rng.choice([1, 2, 3, 4, 5], size=int(rng.integers(1, 3)), replace=False).tolist()))
    images = synthetic.blank_images(H, W, bands=bands)
    S = int(rng.integers(1, 4))
    catalog = [synthetic.sample_ce([rng.uniform(3, H - 3), rng.uniform(3, W - 3)], bool(rng.integers(0, 2))) for _ in range(S)]
    synthetic.gen_images(images, catalog, seed=seed, device="cpu")
    patches = get_sky_patches(images, catalog, radius_override_pix=float(rng.uniform(4, 28)))
    vp = [cj.catalog_init_source(ce) for ce in catalog]
    synthetic.perturb_params(vp)
    for v in vp:
        v[ids.gal_radius_px] = float(np.exp(rng.uniform(np.log(0.02), np.log(60.0))))
        v[ids.gal_axis_ratio] = float(rng.choice([0.02, 0.3, 0.99, rng.uniform(0.05, 1.0)]))
        v[ids.gal_angle] = float(rng.uniform(-4, 8))
        v[ids.gal_frac_dev] = float(rng.choice([0.0, 1.0, 0.01, 0.99, rng.uniform()]))
        a = float(rng.choice([0.0, 1.0, 1e-4, 0.9999, rng.uniform()]))
        v[ids.is_star[0]], v[ids.is_star[1]] = a, 1 - a
        v[ids.pos[0]] += rng.normal(0, 6)
        v[ids.pos[1]] += rng.normal(0, 6)
    return images, patches, _all_tasks(vp)


CASES = {
    "star_1band": case_star_1band,
    "star_5band": case_star_5band,
    "galaxy": case_galaxy,
    "two_body": case_two_body,
    "config2": case_config2,
    "config2_rotated_wcs": case_config2_rotated_wcs,
    "masked": case_masked,
    "clipped_and_empty": case_clipped_and_empty,
    "psf_k1": case_psf_k1,
    "psf_k3": case_psf_k3,
    "crowded": case_crowded,
    "small_field": case_small_field,
    "wide_patch": case_wide_patch,
    "seven_images": case_seven_images,
    "sharp_psf": case_sharp_psf,
}

_cache = {}


def get(name):
    if name not in _cache:
        _cache[name] = CASES[name]()
    return _cache[name]


def assert_parity(ref, got, mode, label=""):
    """The parity statement of SURVEY.md 8c: value |d|/|ref| <= 1e-8; gradient / Hessian component-wise
    |d| <= 1e-8 * max(|ref_ij|, ||ref||_inf * 1e-6); counters and flags exact."""
    assert np.array_equal(ref["counters"], got["counters"]), (label, ref["counters"], got["counters"])
    assert np.array_equal(ref["flags"], got["flags"]), label
    rv, gv = ref["v"], got["v"]
    assert np.all(np.abs(rv - gv) <= 1e-8 * np.abs(rv)), (label, rv, gv)
    if mode >= 1:
        n = len(rv)
        rd, gd = ref["d"].reshape(n, -1), got["d"].reshape(n, -1)
        sc = np.abs(rd).max(axis=1, keepdims=True)
        assert np.all(np.abs(rd - gd) <= 1e-8 * np.maximum(np.abs(rd), sc * 1e-6)), (label, np.abs(rd - gd).max())
    if mode >= 2:
        rh, gh = ref["h"].reshape(n, -1), got["h"].reshape(n, -1)
        sc = np.abs(rh).max(axis=1, keepdims=True)
        assert np.all(np.abs(rh - gh) <= 1e-8 * np.maximum(np.abs(rh), sc * 1e-6)), (label, np.abs(rh - gh).max())


def all_vp(patches, tasks):
    """44 x S matrix of the variational parameters of every source of a case (rows in patch order)."""
    S = patches.shape[0]
    vp = [None] * S
    for rows, act, v in tasks:
        for j, r in enumerate(rows):
            vp[r - 1] = np.asarray(v)[:, j]
    assert all(x is not None for x in vp)
    return np.stack(vp, axis=1)


def assert_render_parity(ref, got, what="", rtol=1e-12):
    """E_G - sky images: |delta| <= rtol * max|ref| per image (the subtraction of the Float32 sky leaves an
    absolute error of a few ulp of the sky level)."""
    assert len(ref) == len(got)
    for n, (r, g) in enumerate(zip(ref, got)):
        assert r.shape == g.shape
        scale = max(np.abs(r).max(), 1.0)
        err = np.abs(r - g).max()
        assert err <= rtol * scale, f"{what} image {n}: max |delta| {err:.3e} vs scale {scale:.3e}"
        assert np.array_equal(r == 0.0, g == 0.0) or np.abs(r[(r == 0.0) != (g == 0.0)]).max() <= rtol * scale
