"""Host-side data model: the reference's `Model` types that the ELBO path consumes.

Mirrors (names, field meaning, 1-based pixel conventions) of
  src/model/param_set.jl        -- parameter index sets (`ids`, `bids`, ...)
  src/model/light_source_model.jl -- CatalogEntry, galaxy prototypes, prior constants
  src/model/psf_model.jl        -- PsfComponent, render_psf, get_psf_width
  src/model/image_model.jl      -- Image
  src/model/imaged_sources.jl   -- ImagePatch, box helpers, get_sky_patches, find_neighbors
  src/model/wcs_utils.jl        -- linear_world_to_pix
The host builds these once per box; the CUDA library consumes flattened copies
(include/celeste_cuda.h).  Nothing here is on the per-pixel hot path.

All matrices are numpy arrays indexed [h-1, w-1]; they are handed to the C ABI in
column-major (Fortran) order, matching Julia.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

NUM_BANDS = 5
NUM_SOURCE_TYPES = 2          # light_source_model.jl:7
NUM_COLOR_COMPONENTS = 8      # light_source_model.jl:4


class _CanonicalParams:
    """param_set.jl:76-103, as 0-based numpy index arrays (reference id - 1)."""
    pos = np.array([0, 1])
    gal_frac_dev = 2
    gal_axis_ratio = 3
    gal_angle = 4
    gal_radius_px = 5
    flux_loc = np.array([6, 7])
    flux_scale = np.array([8, 9])
    color_mean = np.arange(10, 18).reshape(2, 4).T   # [m, i] -> 10 + m + 4 i
    color_var = np.arange(18, 26).reshape(2, 4).T
    is_star = np.array([26, 27])
    k = np.arange(28, 44).reshape(2, 8).T            # [c, i]

    def __len__(self):
        return 44


ids = _CanonicalParams()
NUM_PARAMS = 44


def bright_ids(i: int) -> np.ndarray:
    """param_set.jl:163 (0-based)."""
    return np.concatenate([[ids.flux_loc[i]], [ids.flux_scale[i]], ids.color_mean[:, i], ids.color_var[:, i]])


@dataclass
class CatalogEntry:
    """light_source_model.jl:10-19."""
    pos: np.ndarray
    is_star: bool
    star_fluxes: np.ndarray
    gal_fluxes: np.ndarray
    gal_frac_dev: float
    gal_axis_ratio: float
    gal_angle: float
    gal_radius_px: float


def get_galaxy_prototypes():
    """light_source_model.jl:45-72 -> ((eta_dev, nu_dev), (eta_exp, nu_exp))."""
    dev_amp = np.array([4.26347652e-2, 2.40127183e-1, 6.85907632e-1, 1.51937350,
                        2.83627243, 4.46467501, 5.72440830, 5.60989349])
    dev_amp = dev_amp / dev_amp.sum()
    dev_var = np.array([2.23759216e-4, 1.00220099e-3, 4.18731126e-3, 1.69432589e-2,
                        6.84850479e-2, 2.87207080e-1, 1.33320254, 8.40215071])
    exp_amp = np.array([2.34853813e-3, 3.07995260e-2, 2.23364214e-1,
                        1.17949102, 4.33873750, 5.99820770])
    exp_amp = exp_amp / exp_amp.sum()
    exp_var = np.array([1.20078965e-3, 8.84526493e-3, 3.91463084e-2,
                        1.39976817e-1, 4.60962500e-1, 1.50159566])
    effective_radii = [1.078031, 0.928896]
    dev_var = dev_var / effective_radii[0] ** 2
    exp_var = exp_var / effective_radii[1] ** 2
    return (dev_amp, dev_var), (exp_amp, exp_var)


galaxy_prototypes = get_galaxy_prototypes()


@dataclass
class PriorParams:
    """light_source_model.jl:78-133 (`load_prior_init`)."""
    is_star: np.ndarray
    flux_mean: np.ndarray
    flux_var: np.ndarray
    k: np.ndarray            # 8 x 2
    color_mean: np.ndarray   # 4 x 8 x 2
    color_cov: np.ndarray    # 4 x 4 x 8 x 2
    gal_radius_px_mean: float
    gal_radius_px_var: float


def load_prior() -> PriorParams:
    fn = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "celeste_priors.json")
    raw = json.load(open(fn))
    k = np.zeros((8, 2))
    cm = np.zeros((4, 8, 2))
    cc = np.zeros((4, 4, 8, 2))
    for i, name in enumerate(("star", "gal")):
        k[:, i] = raw[name]["c_weights"]
        cm[:, :, i] = np.array(raw[name]["c_means"]).T
        cc[:, :, :, i] = np.transpose(np.array(raw[name]["c_covs"]), (2, 1, 0))
    return PriorParams(np.array([0.95, 0.05]), np.array([1.5035546, 1.07431]),
                       np.array([1.9039063 ** 2, 1.1177502 ** 2]), k, cm, cc,
                       0.5015693, 0.8590007 ** 2)


# --------------------------------------------------------------------------- PSF
@dataclass
class PsfComponent:
    """psf_model.jl:17-29 (tauBarInv / tauBarLd are not used by the ELBO path)."""
    alphaBar: float
    xiBar: np.ndarray     # (2,)
    tauBar: np.ndarray    # (2, 2)

    def flat7(self) -> np.ndarray:
        t = np.asarray(self.tauBar, dtype=np.float64)
        return np.array([self.alphaBar, self.xiBar[0], self.xiBar[1], t[0, 0], t[1, 0], t[0, 1], t[1, 1]])


def get_psf_width(psf: Sequence[PsfComponent], width_scale: float = 1.0) -> float:
    """psf_model.jl:32-52."""
    alpha_norm = sum(pc.alphaBar for pc in psf)
    cov_est = np.zeros((2, 2))
    for pc in psf:
        xi = np.asarray(pc.xiBar, dtype=np.float64)
        cov_est += pc.alphaBar * (np.outer(xi, xi) + np.asarray(pc.tauBar)) / alpha_norm
    return width_scale * math.sqrt(np.linalg.eigvalsh(cov_est)[-1]) * alpha_norm


def render_psf(psf: Sequence[PsfComponent], dims: Tuple[int, int]) -> np.ndarray:
    """psf_model.jl:61-75: the mixture rasterised on a grid centred at (dims+1)/2."""
    c0, c1 = (dims[0] + 1) / 2, (dims[1] + 1) / 2
    ii = np.arange(1, dims[0] + 1)[:, None] - c0
    jj = np.arange(1, dims[1] + 1)[None, :] - c1
    stamp = np.zeros(dims)
    for pc in psf:
        t = np.asarray(pc.tauBar, dtype=np.float64)
        det = t[0, 0] * t[1, 1] - t[0, 1] * t[1, 0]
        inv = np.array([[t[1, 1], -t[0, 1]], [-t[1, 0], t[0, 0]]]) / det
        dx = ii - pc.xiBar[0]
        dy = jj - pc.xiBar[1]
        q = inv[0, 0] * dx * dx + (inv[0, 1] + inv[1, 0]) * dx * dy + inv[1, 1] * dy * dy
        stamp += pc.alphaBar * np.exp(-0.5 * q) / (2 * math.pi * math.sqrt(det))
    return stamp


def softpluslike(x: np.ndarray) -> np.ndarray:
    """fsm_util.jl:221."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(1000 * x > 1, 1000 * x - 1, np.log(1000 * x))


def cubic_bspline_prefilter(grid: np.ndarray) -> np.ndarray:
    """Coefficients of Interpolations.jl `interpolate(grid, BSpline(Cubic(Line())), OnGrid())`.

    Un-vendored dependency (REQUIRE:21), restated: coefficient array padded by one per
    side; separable tridiagonal system with interior rows (1/6, 2/3, 1/6) and "Line"
    boundary rows (1, -2, 1) = 0, i.e. zero second derivative at the edge grid points.
    The host (Julia) passes `itp.coefs` itself across the ABI, so this restatement only
    feeds the synthetic inputs.  Self-check in tests: the interpolant reproduces `grid`.
    """
    def solve_axis(a: np.ndarray) -> np.ndarray:   # along axis 0
        n = a.shape[0]
        M = np.zeros((n + 2, n + 2))
        rhs = np.zeros((n + 2,) + a.shape[1:])
        M[0, 0:3] = [1.0, -2.0, 1.0]
        M[n + 1, n - 1:n + 2] = [1.0, -2.0, 1.0]
        for i in range(n):
            M[i + 1, i:i + 3] = [1 / 6, 2 / 3, 1 / 6]
            rhs[i + 1] = a[i]
        return np.linalg.solve(M, rhs.reshape(n + 2, -1)).reshape((n + 2,) + a.shape[1:])
    c = solve_axis(np.asarray(grid, dtype=np.float64))
    c = solve_axis(c.T).T
    return np.ascontiguousarray(c)


def psf_spline_coefs(stamp: np.ndarray) -> np.ndarray:
    """imaged_sources.jl:97-107: clamp, +1e-6, normalise, softpluslike, cubic prefilter."""
    g = np.maximum(np.asarray(stamp, dtype=np.float64), 0.0)
    g = g + 1e-6
    g = g / g.sum()
    return cubic_bspline_prefilter(softpluslike(g))


# --------------------------------------------------------------------------- WCS
@dataclass
class AffineWCS:
    """pix = A (world - world0) + pix0.  The reference uses WCS.jl; its tests use the
    identity transform (test/SampleData.jl:30-34).  `pixel_world_jacobian`
    (wcs_utils.jl:36-51) of an affine map is A exactly."""
    A: np.ndarray = field(default_factory=lambda: np.eye(2))
    world0: np.ndarray = field(default_factory=lambda: np.zeros(2))
    pix0: np.ndarray = field(default_factory=lambda: np.zeros(2))

    def world_to_pix(self, world):
        return self.A @ (np.asarray(world, dtype=np.float64) - self.world0) + self.pix0

    def pix_to_world(self, pix):
        return np.linalg.solve(self.A, np.asarray(pix, dtype=np.float64) - self.pix0) + self.world0


def linear_world_to_pix(wcs_jacobian, world_offset, pix_offset, worldcoords):
    """wcs_utils.jl:14-18."""
    return wcs_jacobian @ (np.asarray(worldcoords) - world_offset) + pix_offset


# --------------------------------------------------------------------------- Image
class Image:
    """image_model.jl:6-38.  `sky` is the dense H x W Float32 matrix (for SDSS the host
    materialises SDSSBackground, SDSSIO.jl:56-99); `psf_stamp` stands for a
    ConstantPSFMap (psf_model.jl:92-95)."""

    def __init__(self, pixels, b: int, wcs: AffineWCS, psf: List[PsfComponent], sky, nelec_per_nmgy,
                 psf_stamp: Optional[np.ndarray] = None):
        self.pixels = np.asfortranarray(pixels, dtype=np.float32)
        self.H, self.W = self.pixels.shape
        self.b = int(b)                      # band id 1..5
        self.wcs = wcs
        self.psf = list(psf)
        self.sky = np.asfortranarray(sky, dtype=np.float32)
        self.nelec_per_nmgy = np.ascontiguousarray(nelec_per_nmgy, dtype=np.float32)
        assert self.sky.shape == (self.H, self.W) and self.nelec_per_nmgy.shape == (self.H,)
        self.psf_stamp = render_psf(self.psf, (51, 51)) if psf_stamp is None else np.asarray(psf_stamp)
        self._coefs_cache = None

    def psfmap(self, x, y) -> np.ndarray:
        return self.psf_stamp.copy()

    def spline_coefs(self) -> np.ndarray:
        # one prefilter per image for a ConstantPSFMap; patches share the array
        if self._coefs_cache is None:
            self._coefs_cache = np.asfortranarray(psf_spline_coefs(self.psf_stamp))
        return self._coefs_cache

    def log_iota(self) -> np.ndarray:
        """Float64(log(iota::Float32)), elbo_objective.jl:292."""
        return np.log(self.nelec_per_nmgy).astype(np.float64)


# --------------------------------------------------------------------------- patches
Box = Tuple[Tuple[int, int], Tuple[int, int]]   # ((first1, last1), (first2, last2)) inclusive, 1-based


def _clamp(x, lo, hi):
    return max(lo, min(hi, x))


def clamp_box(box: Box, dims) -> Box:
    """imaged_sources.jl:10-14."""
    return ((_clamp(box[0][0], 1, dims[0] + 1), _clamp(box[0][1], 0, dims[0])),
            (_clamp(box[1][0], 1, dims[1] + 1), _clamp(box[1][1], 0, dims[1])))


def boxes_overlap(b1: Box, b2: Box) -> bool:
    """imaged_sources.jl:36-40."""
    def ov(r1, r2):
        return r1[0] <= r2[1] and r2[0] <= r1[1]
    return ov(b1[0], b2[0]) and ov(b1[1], b2[1])


def _round_half_even(x: float) -> int:
    # Julia's round(Int, x) rounds half to even, as does Python's round()
    return int(round(x))


class ImagePatch:
    """imaged_sources.jl:60-117."""

    def __init__(self, img: Image, box: Box):
        box = clamp_box(box, (img.H, img.W))
        self.box = box
        self.pixel_center = np.array([(box[0][0] + box[0][1]) / 2, (box[1][0] + box[1][1]) / 2])
        self.world_center = img.wcs.pix_to_world(self.pixel_center)
        self.wcs_jacobian = np.array(img.wcs.A, dtype=np.float64)       # pixel_world_jacobian
        self.psf = img.psf
        self.bitmap_offset = np.array([box[0][0] - 1, box[1][0] - 1], dtype=np.int64)
        h0, h1, w0, w1 = box[0][0], box[0][1], box[1][0], box[1][1]
        sub = img.pixels[max(h0 - 1, 0):max(h1, 0), max(w0 - 1, 0):max(w1, 0)]
        self.active_pixel_bitmap = np.asfortranarray(~np.isnan(sub))
        self.itp_coefs = img.spline_coefs()     # itp_psf.coefs (53 x 53)

    @property
    def shape(self):
        return self.active_pixel_bitmap.shape


class PatchSpec:
    """What the host contributes to ImagePatch(img, box) when the device builds the rest
    (celeste_patches_build): the clamped box, the WCS linearisation at its centre (imaged_sources.jl:83-90) and
    where the PSF stamp comes from -- `grid_psf` = img.psfmap(centre) raw, or None to rasterise img.psf on the
    device (render_psf).  No bitmap, no spline coefficients."""

    def __init__(self, img: Image, box: Box, raw_stamp: bool = True):
        box = clamp_box(box, (img.H, img.W))
        self.box = box
        self.pixel_center = np.array([(box[0][0] + box[0][1]) / 2, (box[1][0] + box[1][1]) / 2])
        self.world_center = img.wcs.pix_to_world(self.pixel_center)
        self.wcs_jacobian = np.array(img.wcs.A, dtype=np.float64)
        self.psf = img.psf
        self.bitmap_offset = np.array([box[0][0] - 1, box[1][0] - 1], dtype=np.int64)
        self.H2 = max(box[0][1] - box[0][0] + 1, 0)
        self.W2 = max(box[1][1] - box[1][0] + 1, 0)
        self.grid_psf = img.psf_stamp if raw_stamp else None       # ConstantPSFMap: one array per image, shared
        self.grid_n = img.psf_stamp.shape[0]


def get_sky_patch_specs(images: Sequence[Image], catalog: Sequence[CatalogEntry], radius_override_pix=float("nan"),
                        raw_stamp: bool = True):
    """get_sky_patches (imaged_sources.jl:165-183) up to the point where pixels / the PSF stamp are touched:
    S x N object array of PatchSpec for DeviceField.build_patches."""
    S, N = len(catalog), len(images)
    specs = np.empty((S, N), dtype=object)
    for n in range(N):
        for s in range(S):
            if math.isnan(radius_override_pix):
                box = box_from_catalog(images[n], catalog[s], width_scale=1.2)
            else:
                box = box_around_point(images[n].wcs, catalog[s].pos, radius_override_pix)
            specs[s, n] = PatchSpec(images[n], box, raw_stamp)
    return specs


def box_around_point(wcs: AffineWCS, world_center, pixel_radius) -> Box:
    """imaged_sources.jl:126-136."""
    pc = wcs.world_to_pix(world_center)
    return ((_round_half_even(pc[0] - pixel_radius), _round_half_even(pc[0] + pixel_radius)),
            (_round_half_even(pc[1] - pixel_radius), _round_half_even(pc[1] + pixel_radius)))


def choose_patch_radius(ce: CatalogEntry, img: Image, width_scale=1.0, max_radius=25) -> float:
    """imaged_sources.jl:197-223."""
    psf_width = get_psf_width(img.psf, width_scale=width_scale)
    obj_width = 0.0 if ce.is_star else width_scale * ce.gal_radius_px / 0.67
    obj_width += psf_width
    flux = ce.star_fluxes[img.b - 1] if ce.is_star else ce.gal_fluxes[img.b - 1]
    assert flux > 0.0
    epsilon = float(img.sky[img.H // 2 - 1, img.W // 2 - 1])
    pdf_90 = math.exp(-0.5 * 1.64 ** 2) / (math.sqrt(2 * math.pi) * obj_width)
    pdf_target = min(pdf_90, epsilon / (20 * flux))
    rhs = math.log(pdf_target) + 0.5 * math.log(2 * math.pi) + math.log(obj_width)
    radius_req = math.sqrt(-2 * obj_width ** 2 * rhs)
    return min(radius_req, max_radius)


def box_from_catalog(img: Image, ce: CatalogEntry, width_scale=1.0, max_radius=25) -> Box:
    """imaged_sources.jl:147-159."""
    r = choose_patch_radius(ce, img, width_scale=width_scale, max_radius=max_radius)
    return box_around_point(img.wcs, ce.pos, r)


def get_sky_patches(images: Sequence[Image], catalog: Sequence[CatalogEntry], radius_override_pix=float("nan")):
    """imaged_sources.jl:165-183 -> S x N object array of ImagePatch."""
    S, N = len(catalog), len(images)
    patches = np.empty((S, N), dtype=object)
    for n in range(N):
        for s in range(S):
            if math.isnan(radius_override_pix):
                box = box_from_catalog(images[n], catalog[s], width_scale=1.2)
            else:
                box = box_around_point(images[n].wcs, catalog[s].pos, radius_override_pix)
            patches[s, n] = ImagePatch(images[n], box)
    return patches


def find_neighbors(patches: np.ndarray, target: int) -> List[int]:
    """imaged_sources.jl:232-244 (0-based indices in and out)."""
    out = []
    S, N = patches.shape
    for i in range(S):
        if i == target:
            continue
        for j in range(N):
            if boxes_overlap(patches[target, j].box, patches[i, j].box):
                out.append(i)
                break
    return out


def find_all_neighbors(patches: np.ndarray) -> List[List[int]]:
    """find_neighbors for every source at once (vectorised box test; same result)."""
    S, N = patches.shape
    lo1 = np.array([[patches[s, n].box[0][0] for n in range(N)] for s in range(S)])
    hi1 = np.array([[patches[s, n].box[0][1] for n in range(N)] for s in range(S)])
    lo2 = np.array([[patches[s, n].box[1][0] for n in range(N)] for s in range(S)])
    hi2 = np.array([[patches[s, n].box[1][1] for n in range(N)] for s in range(S)])
    out = []
    for t in range(S):
        ov = ((lo1[t] <= hi1) & (lo1 <= hi1[t]) & (lo2[t] <= hi2) & (lo2 <= hi2[t])).any(axis=1)
        ov[t] = False
        out.append(np.nonzero(ov)[0].tolist())
    return out
