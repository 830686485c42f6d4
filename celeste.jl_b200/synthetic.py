"""Synthetic SDSS-shaped inputs for tests and the benchmark.

Host-side analogue of src/Synthetic.jl (gen_image!:30-47), test/SampleData.jl and the
catalog sampler of src/AccuracyBenchmark.jl:395-470.  All physical constants are the
reference's own "SDSS-like" values:
  sky  (nmgy)   u,g,r,i,z = 0.2696, 0.3425, 0.7748, 1.6903, 4.9176   benchmark/galsim/galsim_field.py:15
  iota (e-/nmgy)          = 146.9, 838.1, 829.8, 597.2, 129.8        benchmark/galsim/galsim_field.py:16
  PSF sigma 2.29 px (galsim_field.py:14); K = 2 (elbo_args.jl:197)
The reference's own fixtures need SDSS field 3900-6-269 from the network
(test/SampleData.jl:144-158) and cannot be materialised offline; these are the
synthetic analogues.  The renderer is plain torch tensor code (plumbing, CPU or GPU);
it is input generation, not the product hot path.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .deterministic_vi import catalog_init_source
from .model import (AffineWCS, CatalogEntry, Image, PsfComponent, box_around_point, find_all_neighbors,
                    galaxy_prototypes, get_sky_patches, ids, load_prior)

SDSS_SKY_NMGY = [0.2696, 0.3425, 0.7748, 1.6903, 4.9176]
SDSS_IOTA = [146.9, 838.1, 829.8, 597.2, 129.8]
SDSS_PSF_SIGMA_PX = 2.29

# test/SampleData.jl:23-27
sample_star_fluxes = np.array([4.451805E+03, 1.491065E+03, 2.264545E+03, 2.027004E+03, 1.846822E+04])
sample_galaxy_fluxes = np.array([1.377666E+01, 5.635334E+01, 1.258656E+02, 1.884264E+02, 2.351820E+02]) * 100


def make_simple_psf(psf_sigma_px: float) -> List[PsfComponent]:
    """AccuracyBenchmark.jl:504-516: weights [1, 0], one sigma."""
    t = np.array([[psf_sigma_px ** 2, 0.0], [0.0, psf_sigma_px ** 2]])
    return [PsfComponent(a, np.zeros(2), t.copy()) for a in (1.0, 0.0)]


def make_two_component_psf(band: int = 3) -> List[PsfComponent]:
    """A non-degenerate K = 2 PSF: weights (0.8, 0.2), sigma (1.2, 2.9) px,
    plus a small band-dependent offset/shear so xiBar and the off-diagonal of tauBar are exercised."""
    e = 0.01 * (band - 3)
    return [PsfComponent(0.8, np.array([0.02 + e, -0.015]), np.array([[1.44, 0.05 + e], [0.05 + e, 1.5]])),
            PsfComponent(0.2, np.array([-0.06, 0.045 - e]), np.array([[8.41, -0.2], [-0.2, 8.0]]))]


def blank_images(H: int, W: int, psf_kind: str = "two", wcs: Optional[AffineWCS] = None,
                 bands: Sequence[int] = (1, 2, 3, 4, 5)) -> List[Image]:
    """Five (or fewer) empty SDSS-shaped images; make_image of AccuracyBenchmark.jl:553-571."""
    out = []
    for b in bands:
        psf = make_two_component_psf(b) if psf_kind == "two" else make_simple_psf(SDSS_PSF_SIGMA_PX)
        out.append(Image(np.zeros((H, W), dtype=np.float32), b, wcs or AffineWCS(), psf,
                         np.full((H, W), SDSS_SKY_NMGY[b - 1], dtype=np.float32),
                         np.full((H,), SDSS_IOTA[b - 1], dtype=np.float32)))
    return out


# ------------------------------------------------------------------ renderer (Synthetic.jl:17-47)
def _spline_value(coefs: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Cubic B-spline value (Interpolations.jl rule, see model.cubic_bspline_prefilter) at 1-based (x, y)."""
    n1, n2 = coefs.shape
    ix = torch.clamp(torch.floor(x), 1, n1 - 3).to(torch.long)
    iy = torch.clamp(torch.floor(y), 1, n2 - 3).to(torch.long)
    fx = x - ix
    fy = y - iy

    def w(f):
        omf = 1 - f
        return [omf ** 3 / 6, 2 / 3 - f * f + 0.5 * f ** 3, 2 / 3 - omf * omf + 0.5 * omf ** 3, f ** 3 / 6]
    wx, wy = w(fx), w(fy)
    val = torch.zeros_like(x)
    for b in range(4):
        r = torch.zeros_like(x)
        for a in range(4):
            r = r + wx[a] * coefs[ix - 1 + a, iy - 1 + b]
        val = val + wy[b] * r
    return val


def _star_density(img: Image, world_pos: torch.Tensor, hh: torch.Tensor, ww: torch.Tensor, dev) -> torch.Tensor:
    """fsm_util.jl:225-240 value path for a batch: world_pos (B,2); hh, ww (B,P) 1-based pixel coords."""
    A = torch.as_tensor(img.wcs.A, dtype=torch.float64, device=dev)
    w0 = torch.as_tensor(img.wcs.world0, dtype=torch.float64, device=dev)
    p0 = torch.as_tensor(img.wcs.pix0, dtype=torch.float64, device=dev)
    m = (world_pos - w0) @ A.T + p0        # == linear_world_to_pix for an affine wcs
    coefs = torch.as_tensor(np.ascontiguousarray(img.spline_coefs()), dtype=torch.float64, device=dev)
    y = _spline_value(coefs, hh - m[:, 0:1] + 26, ww - m[:, 1:2] + 26)
    return torch.where(y < 0, 1e-3 * torch.exp(y), 1e-3 * (y + 1))


def _galaxy_density(img: Image, world_pos, frac_dev, axis_ratio, angle, radius, hh, ww, dev) -> torch.Tensor:
    """fsm_util.jl:194-219 value path (load_bvn_mixtures!:111-169) for a batch."""
    A = torch.as_tensor(img.wcs.A, dtype=torch.float64, device=dev)
    w0 = torch.as_tensor(img.wcs.world0, dtype=torch.float64, device=dev)
    p0 = torch.as_tensor(img.wcs.pix0, dtype=torch.float64, device=dev)
    m = (world_pos - w0) @ A.T + p0
    cp, sp = torch.cos(angle), torch.sin(angle)
    ab_term = axis_ratio ** 2 - 1
    ss = radius ** 2
    x11 = ss * (1 + ab_term * sp * sp)
    x12 = -ss * cp * sp * ab_term
    x22 = ss * (1 + ab_term * cp * cp)
    out = torch.zeros_like(hh)
    for i, (eta, nu) in enumerate(galaxy_prototypes):
        theta = frac_dev if i == 0 else 1 - frac_dev
        for j in range(len(eta)):
            for pc in img.psf:
                t = np.asarray(pc.tauBar, dtype=np.float64)
                s11 = t[0, 0] + nu[j] * x11
                s12 = t[0, 1] + nu[j] * x12
                s22 = t[1, 1] + nu[j] * x22
                det = s11 * s22 - s12 * s12
                z = pc.alphaBar * eta[j] / (2 * math.pi * torch.sqrt(det))
                dx = hh - (pc.xiBar[0] + m[:, 0:1])
                dy = ww - (pc.xiBar[1] + m[:, 1:2])
                q = (s22[:, None] * dx * dx - 2 * s12[:, None] * dx * dy + s11[:, None] * dy * dy) / det[:, None]
                out = out + (theta * z)[:, None] * torch.exp(-0.5 * q)
    return out


def gen_images(images: Sequence[Image], catalog: Sequence[CatalogEntry], seed: int = 1,
               expectation: bool = False, device: Optional[str] = None, batch: int = 512):
    """Synthetic.gen_images! (Synthetic.jl:30-58): sky + every body rendered on its radius-25
    box, times iota, then Poisson.  In place on img.pixels."""
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    gen = torch.Generator(device="cpu").manual_seed(seed)
    off = torch.arange(51, device=dev)
    for img in images:
        H, W = img.H, img.W
        acc = torch.as_tensor(np.ascontiguousarray(img.sky), dtype=torch.float64, device=dev).clone().reshape(-1)
        for kind in (True, False):
            idx = [k for k, ce in enumerate(catalog) if ce.is_star == kind]
            for b0 in range(0, len(idx), batch):
                sel = [catalog[k] for k in idx[b0:b0 + batch]]
                pos = torch.tensor(np.array([ce.pos for ce in sel]), dtype=torch.float64, device=dev)
                boxes = [box_around_point(img.wcs, ce.pos, 25) for ce in sel]
                h0 = torch.tensor([bx[0][0] for bx in boxes], device=dev)
                w0 = torch.tensor([bx[1][0] for bx in boxes], device=dev)
                hh = (h0[:, None, None] + off[None, :, None]).expand(-1, 51, 51).reshape(len(sel), -1)
                ww = (w0[:, None, None] + off[None, None, :]).expand(-1, 51, 51).reshape(len(sel), -1)
                ok = (hh >= 1) & (hh <= H) & (ww >= 1) & (ww <= W)
                hf, wf = hh.to(torch.float64), ww.to(torch.float64)
                if kind:
                    flux = torch.tensor([ce.star_fluxes[img.b - 1] for ce in sel], dtype=torch.float64, device=dev)
                    dens = _star_density(img, pos, hf, wf, dev)
                else:
                    flux = torch.tensor([ce.gal_fluxes[img.b - 1] for ce in sel], dtype=torch.float64, device=dev)
                    g = lambda f: torch.tensor([getattr(ce, f) for ce in sel], dtype=torch.float64, device=dev)
                    dens = _galaxy_density(img, pos, g("gal_frac_dev"), g("gal_axis_ratio"), g("gal_angle"),
                                           g("gal_radius_px"), hf, wf, dev)
                lin = ((hh - 1) * W + (ww - 1))[ok]            # row-major flat index
                acc.index_add_(0, lin, (dens * flux[:, None])[ok])
        iota = torch.as_tensor(img.nelec_per_nmgy, dtype=torch.float64, device=dev)
        lam = (acc.reshape(H, W) * iota[:, None]).cpu()
        if not expectation:
            lam = torch.poisson(lam.clamp_min(0), generator=gen)
        img.pixels = np.asfortranarray(lam.numpy().astype(np.float32))


def catalog_truth_vp(ce: CatalogEntry) -> np.ndarray:
    """Variational parameters under which the model's expected flux IS the catalog entry's (Synthetic.jl:17-27
    write_star / write_galaxy use ce.star_fluxes / ce.gal_fluxes directly): is_star 0 / 1, flux_scale = color_var = 0,
    flux_loc = log r-band flux, colours = log flux ratios, so E_l[b] = exp(kappa_b . beta) = flux[b] exactly; the galaxy
    shape as the catalog has it (no clamping)."""
    vs = np.zeros(44)
    vs[ids.pos] = ce.pos
    vs[ids.is_star] = [1.0, 0.0] if ce.is_star else [0.0, 1.0]
    for i, fl in enumerate((ce.star_fluxes, ce.gal_fluxes)):
        fl = np.maximum(np.asarray(fl, dtype=np.float64), 1e-300)
        vs[ids.flux_loc[i]] = math.log(fl[2])
        vs[ids.color_mean[:, i]] = np.log(fl[1:] / fl[:-1])
    vs[ids.gal_frac_dev] = ce.gal_frac_dev
    vs[ids.gal_axis_ratio] = ce.gal_axis_ratio
    vs[ids.gal_angle] = ce.gal_angle
    vs[ids.gal_radius_px] = ce.gal_radius_px
    return vs


def gen_images_device(images: Sequence[Image], catalog: Sequence[CatalogEntry], seed: int = 1, expectation: bool = False,
                      device: int = -1):
    """Synthetic.gen_images! (Synthetic.jl:30-58) with the per-body render on the GPU: every body's expected flux on
    its radius-25 box through celeste_render_boxes (the value-only render kernel over whole boxes) at the catalog's own
    fluxes (catalog_truth_vp), then + sky, x iota and the Poisson draw on the host.  In place on img.pixels (Float32,
    like the reference; the reference also ACCUMULATES in Float32 -- here the sum over bodies is Float64 and rounded
    once, a difference of a few Float32 ulp of the pixel value)."""
    from .deterministic_vi import DeviceField
    from .model import ImagePatch
    S, N = len(catalog), len(images)
    for img in images:
        img.pixels = np.zeros((img.H, img.W), dtype=np.float32, order="F")     # no masked pixels: full bitmaps
    patches = np.empty((S, N), dtype=object)
    for n, img in enumerate(images):
        for s, ce in enumerate(catalog):
            patches[s, n] = ImagePatch(img, box_around_point(img.wcs, ce.pos, 25))
    field = DeviceField(images, patches, device=device)
    vp = np.stack([catalog_truth_vp(ce) for ce in catalog], axis=1) if S else np.zeros((44, 0))
    add = field.render_expectation(np.arange(1, S + 1), vp, full_box=True)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    for img, a in zip(images, add):
        lam = (a + np.asarray(img.sky, dtype=np.float64)) * np.asarray(img.nelec_per_nmgy, dtype=np.float64)[:, None]
        if not expectation:
            lam = torch.poisson(torch.from_numpy(np.ascontiguousarray(lam)).clamp_min(0), generator=gen).numpy()
        img.pixels = np.asfortranarray(lam.astype(np.float32))
    return images


# ------------------------------------------------------------------ test/SampleData.jl analogues
def sample_ce(pos, is_star: bool) -> CatalogEntry:
    """SampleData.jl:119-122."""
    return CatalogEntry(np.array(pos, dtype=np.float64), is_star, sample_star_fluxes.copy(),
                        sample_galaxy_fluxes.copy(), 0.1, 0.7, math.pi / 4, 4.0)


def perturb_params(vp):
    """SampleData.jl:126-141."""
    for vs in vp:
        vs[ids.is_star] = [0.4, 0.6]
        vs[ids.pos[0]] += .8
        vs[ids.pos[1]] -= .7
        vs[ids.flux_loc] -= math.log(10)
        vs[ids.flux_scale] *= 25.
        vs[ids.gal_frac_dev] += 0.05
        vs[ids.gal_axis_ratio] += 0.05
        vs[ids.gal_angle] += math.pi / 10
        vs[ids.gal_radius_px] *= 1.2
        vs[ids.color_mean] += 0.5
        vs[ids.color_var] = 1e-1


def make_elbo_inputs(images, catalog, patch_radius_pix=float("nan"), perturb=True):
    """make_elbo_args (SampleData.jl:96-108) minus the ElboArgs construction:
    returns (patches, vp)."""
    patches = get_sky_patches(images, catalog, radius_override_pix=patch_radius_pix)
    vp = [catalog_init_source(ce) for ce in catalog]
    if perturb:
        perturb_params(vp)
    return patches, vp


def gen_sample_star_dataset(perturb=True, bands=(1, 2, 3, 4, 5), H=20, W=23, seed=1):
    """SampleData.jl:161-173 (20 x 23 crops, one star at [10.1, 12.2])."""
    images = blank_images(H, W, bands=bands)
    catalog = [sample_ce([10.1, 12.2], True)]
    gen_images(images, catalog, seed=seed, device="cpu")
    patches, vp = make_elbo_inputs(images, catalog, perturb=perturb)
    return images, patches, vp, catalog


def gen_sample_galaxy_dataset(perturb=True, seed=1):
    """SampleData.jl:176-188."""
    images = blank_images(20, 23)
    catalog = [sample_ce([8.5, 9.6], False)]
    gen_images(images, catalog, seed=seed, device="cpu")
    patches, vp = make_elbo_inputs(images, catalog, perturb=perturb)
    return images, patches, vp, catalog


def gen_two_body_dataset(perturb=True, seed=1):
    """SampleData.jl:193-208."""
    images = blank_images(20, 23)
    catalog = [sample_ce([4.5, 3.6], False), sample_ce([10.1, 12.1], True)]
    gen_images(images, catalog, seed=seed, device="cpu")
    patches, vp = make_elbo_inputs(images, catalog, perturb=perturb)
    return images, patches, vp, catalog


def gen_three_body_dataset(perturb=True, seed=1, H=112, W=238):
    """SampleData.jl:211-227."""
    images = blank_images(H, W)
    catalog = [sample_ce([4.5, 3.6], False), sample_ce([60.1, 82.2], True), sample_ce([71.3, 100.4], False)]
    gen_images(images, catalog, seed=seed, device="cpu")
    patches, vp = make_elbo_inputs(images, catalog, perturb=perturb)
    return images, patches, vp, catalog


def gen_config2_dataset(seed=1, rotated_wcs=False):
    """BASELINE.json configs[1]: 3 sources (2 stars + 1 galaxy), 5 bands, 50 x 50 tiles,
    patches via radius_override_pix = 25 (SURVEY.md 8d config mapping)."""
    wcs = None
    if rotated_wcs:
        c, s = math.cos(0.3), math.sin(0.3)
        wcs = AffineWCS(A=np.array([[1.1 * c, -0.9 * s], [1.1 * s, 0.9 * c]]), world0=np.array([3.0, -2.0]),
                        pix0=np.array([1.5, 0.5]))
    images = blank_images(50, 50, wcs=wcs)
    pix = [[24.5, 23.6], [14.1, 32.2], [33.3, 18.4]]
    world = [(wcs.pix_to_world(p) if wcs else p) for p in pix]
    catalog = [sample_ce(world[0], False), sample_ce(world[1], True), sample_ce(world[2], True)]
    gen_images(images, catalog, seed=seed, device="cpu")
    patches, vp = make_elbo_inputs(images, catalog, patch_radius_pix=25.0, perturb=True)
    return images, patches, vp, catalog


# ------------------------------------------------------------------ prior catalog (AccuracyBenchmark.jl:395-445)
PRIOR_PROBABILITY_OF_STAR = 0.28


def _fluxes_from_colors(r_flux, colors):
    """Synthetic.jl:66-77 (sample_fluxes)."""
    l = np.zeros(5)
    l[2] = r_flux
    l[3] = l[2] * math.exp(colors[2])
    l[4] = l[3] * math.exp(colors[3])
    l[1] = l[2] / math.exp(colors[1])
    l[0] = l[1] / math.exp(colors[0])
    return l


def draw_catalog(n: int, H: int, W: int, seed: int = 42, margin: float = 0.0,
                 wcs: Optional[AffineWCS] = None) -> List[CatalogEntry]:
    """draw_source_params (AccuracyBenchmark.jl:400-445) with positions uniform over the image."""
    rng = np.random.default_rng(seed)
    prior = load_prior()
    out = []
    for _ in range(n):
        is_star = rng.random() < PRIOR_PROBABILITY_OF_STAR
        t = 0 if is_star else 1
        flux_r = math.exp(rng.normal(prior.flux_mean[t], math.sqrt(prior.flux_var[t])))
        k = rng.choice(8, p=prior.k[:, t])
        colors = rng.multivariate_normal(prior.color_mean[:, k, t], prior.color_cov[:, :, k, t])
        if not is_star:
            radius = math.exp(rng.normal(prior.gal_radius_px_mean, math.sqrt(prior.gal_radius_px_var)))
            angle = rng.uniform(0, math.pi)
            axis_ratio = rng.beta(2, 2)
            frac_dev = rng.beta(0.5, 0.5)
        else:
            radius, angle, axis_ratio, frac_dev = 1.0, 0.0, 0.8, 0.5
        pix = np.array([rng.uniform(0.5 + margin, H + 0.5 - margin), rng.uniform(0.5 + margin, W + 0.5 - margin)])
        pos = wcs.pix_to_world(pix) if wcs else pix
        fl = _fluxes_from_colors(flux_r, colors)
        # both flux vectors are filled (the init code reads star_fluxes and gal_fluxes of every entry)
        out.append(CatalogEntry(pos, bool(is_star), fl.copy(), fl.copy(), float(frac_dev), float(axis_ratio),
                                float(angle), float(radius)))
    return out


class FieldDataset:
    """One synthetic SDSS-shaped field: images, catalog, S x N patches, neighbour lists,
    initial variational parameters -- the inputs of ParallelRun._infer_box (ParallelRun.jl:610-637)."""

    def __init__(self, n_sources: int, H: int = 2048, W: int = 1489, seed: int = 42, pixel_seed: int = 1,
                 psf_kind: str = "two", perturb: bool = True, device: Optional[str] = None):
        self.images = blank_images(H, W, psf_kind=psf_kind)
        self.catalog = draw_catalog(n_sources, H, W, seed=seed)
        gen_images(self.images, self.catalog, seed=pixel_seed, device=device)
        # catalog path: box_from_catalog(width_scale = 1.2, max_radius = 25), imaged_sources.jl:173-176
        self.patches = get_sky_patches(self.images, self.catalog)
        self.neighbors = find_all_neighbors(self.patches)
        self.vp = [catalog_init_source(ce) for ce in self.catalog]
        if perturb:
            perturb_params(self.vp)

    def tasks(self, targets: Optional[Sequence[int]] = None):
        """One (ElboArgs, vp) task per target: [target, neighbours...], active_sources = [1]
        (ParallelRun.jl:236-253, 483-489).  Rows are 1-based."""
        targets = range(len(self.catalog)) if targets is None else targets
        rows, act = [], []
        for t in targets:
            loc = [t] + self.neighbors[t]
            rows.append([r + 1 for r in loc])
            act.append([1])
        return rows, act

    def vp_flat(self, rows) -> np.ndarray:
        return np.concatenate([np.concatenate([self.vp[r - 1] for r in rr]) for rr in rows])
