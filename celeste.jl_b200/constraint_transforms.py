"""Bound <-> free reparameterisation of the variational parameters, batched over sources.

Mirror of src/deterministic_vi/ConstraintTransforms.jl for the constraint table of
ElboMaximize.elbo_constraints (ElboMaximize.jl:63-93):
  * 26 box constraints on canonical ids 1..26 (pos +- loc_width, gal_frac_dev, gal_axis_ratio, gal_angle,
    gal_radius_px, flux_loc, flux_scale, color_mean, color_var), each a scaled logistic
    (to_bound :67-70, to_free :76-79);
  * 3 simplex constraints: is_star (n = 2), k[:, 1], k[:, 2] (n = 8): softmax with an implicit last
    logit 0 and a floor `lower` (to_bound! :89-111, to_free! :114-123).
44 bound parameters <-> 41 free parameters; free index order = boxes in table order, then simplexes
(to_bound!/to_free! over a ConstraintBatch, :189-216).

The reference differentiates `to_bound!` with nested ForwardDiff Jacobians (:360-367) and pushes a
SensitiveFloat through it (`propagate_derivatives!` :373-396: J' g, J' H J + sum_i g_i d2b_i, symmetrize).
Here the Jacobian and the contracted second derivative are closed forms evaluated for all sources at once
with torch tensors (float64, CPU or CUDA) -- row (f.1) of SURVEY.md section 8.
"""
from __future__ import annotations

import numpy as np
import torch

N_BOUND = 44
N_FREE = 41
N_BOX = 26
# simplexes: (first bound index 0-based, n, lower)
SIMPLEXES = ((26, 2, 0.005), (28, 8, 0.01 / 8), (36, 8, 0.01 / 8))
SIMPLEX_SCALE = 1.0


def box_bounds(vp_bound: torch.Tensor, loc_width: float = 1e-4):
    """lower/upper (B x 26) of elbo_constraints (ElboMaximize.jl:70-85); the position box is centred on the
    CURRENT position of each source and therefore must be built once per maximize! and then kept
    (ParallelRun.jl:99-101)."""
    B = vp_bound.shape[0]
    lo = torch.empty((B, N_BOX), dtype=vp_bound.dtype, device=vp_bound.device)
    hi = torch.empty_like(lo)
    lo[:, 0:2] = vp_bound[:, 0:2] - loc_width
    hi[:, 0:2] = vp_bound[:, 0:2] + loc_width
    table = [(2, 3, 1e-2, 0.99), (3, 4, 1e-2, 0.99), (4, 5, -10.0, 10.0), (5, 6, 0.10, 70.0),
             (6, 8, -1.0, 10.0), (8, 10, 1e-4, 0.10), (10, 18, -10.0, 10.0), (18, 26, 1e-4, 1.0)]
    for a, b, l, u in table:
        lo[:, a:b] = l
        hi[:, a:b] = u
    return lo, hi


def enforce(vp: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """enforce! (ConstraintTransforms.jl:222-290): clip into the open boxes / simplexes."""
    out = vp.clone()
    b = out[:, :N_BOX]
    up = torch.nextafter(hi, lo)
    dn = torch.nextafter(lo, hi)
    out[:, :N_BOX] = torch.where((b > lo) & (b < hi), b, torch.maximum(torch.minimum(b, up), dn))
    one_m = float(np.nextafter(1.0, 0.0))
    for first, n, lower in SIMPLEXES:
        s = out[:, first:first + n]
        lower_p = float(np.nextafter(lower, 1.0))
        s = torch.where((s > lower) & (s < 1.0), s, s.clamp(min=lower_p, max=one_m))
        tot = s.sum(dim=1, keepdim=True)
        bad = ~torch.isclose(tot, torch.ones_like(tot), rtol=float(np.sqrt(np.finfo(np.float64).eps)), atol=0.0)
        rescale = (1 - n * lower) / (tot - n * lower)
        s = torch.where(bad, lower_p + rescale * (s - lower), s)
        out[:, first:first + n] = s
    return out


def to_free(vp: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """to_free! (ConstraintTransforms.jl:76-79, 114-123, 199-216); box scale = 1 (ElboMaximize.jl:66)."""
    B = vp.shape[0]
    free = torch.empty((B, N_FREE), dtype=vp.dtype, device=vp.device)
    u = (vp[:, :N_BOX] - lo) / (hi - lo)
    free[:, :N_BOX] = -torch.log(1.0 / u - 1.0)
    f = N_BOX
    for first, n, lower in SIMPLEXES:
        un = (vp[:, first:first + n] - lower) / (1 - n * lower)
        lg = torch.log(un)
        free[:, f:f + n - 1] = SIMPLEX_SCALE * (lg[:, :n - 1] - lg[:, n - 1:n])
        f += n - 1
    return free


def _softmax_last0(z: torch.Tensor) -> torch.Tensor:
    """p (B x n) from the n-1 free logits (implicit last logit 0), max-shifted like to_bound! :97-110."""
    zz = torch.cat([z, torch.zeros_like(z[:, :1])], dim=1)
    m = zz[:, :-1].max(dim=1, keepdim=True).values     # the reference's max runs over the free entries only
    e = torch.exp(zz - m)
    return e / e.sum(dim=1, keepdim=True)


def to_bound(free: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """to_bound! (ConstraintTransforms.jl:67-70, 89-111, 189-196)."""
    B = free.shape[0]
    vp = torch.empty((B, N_BOUND), dtype=free.dtype, device=free.device)
    sig = 1.0 / (1.0 + torch.exp(-free[:, :N_BOX]))
    vp[:, :N_BOX] = sig * (hi - lo) + lo
    f = N_BOX
    for first, n, lower in SIMPLEXES:
        p = _softmax_last0(free[:, f:f + n - 1] / SIMPLEX_SCALE)
        vp[:, first:first + n] = (1 - n * lower) * p + lower
        f += n - 1
    return vp


def propagate_derivatives(free: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, g: torch.Tensor,
                          H: torch.Tensor | None):
    """propagate_derivatives! (ConstraintTransforms.jl:373-396) for Sa = 1, batched.

    g: B x 44 bound gradient, H: B x 44 x 44 bound Hessian (or None).  Returns (g_free B x 41,
    H_free B x 41 x 41 symmetrised).  J is block diagonal (26 scalars + three softmax blocks), so J' H J is
    formed block-wise without materialising a dense 44 x 41 Jacobian product for the box part."""
    B = free.shape[0]
    dt, dev = free.dtype, free.device
    J = torch.zeros((B, N_BOUND, N_FREE), dtype=dt, device=dev)
    sig = 1.0 / (1.0 + torch.exp(-free[:, :N_BOX]))
    d1 = sig * (1 - sig) * (hi - lo)
    d2 = d1 * (1 - 2 * sig)
    idx = torch.arange(N_BOX, device=dev)
    J[:, idx, idx] = d1
    C = torch.zeros((B, N_FREE, N_FREE), dtype=dt, device=dev)      # sum_i g_i d2 b_i / dfree dfree
    C[:, idx, idx] = g[:, :N_BOX] * d2
    f = N_BOX
    for first, n, lower in SIMPLEXES:
        m = n - 1
        p = _softmax_last0(free[:, f:f + m] / SIMPLEX_SCALE)               # B x n
        a = (1 - n * lower)
        eye = torch.eye(n, m, dtype=dt, device=dev)                          # delta_ij for i in bound, j in free
        # dp_i/dz_j = p_i (delta_ij - p_j)
        Jp = p[:, :, None] * (eye[None] - p[:, None, :m])
        J[:, first:first + n, f:f + m] = a * Jp / SIMPLEX_SCALE
        # d2p_i/dz_j dz_k = p_i [(d_ij - p_j)(d_ik - p_k) - p_j (d_jk - p_k)]
        gi = g[:, first:first + n] * a / SIMPLEX_SCALE ** 2                  # B x n
        A = eye[None] - p[:, None, :m]                                       # B x n x m: (d_ij - p_j)
        t1 = torch.einsum("bi,bi,bij,bik->bjk", gi, p, A, A)
        s = (gi * p).sum(dim=1)                                              # sum_i g_i p_i
        pj = p[:, :m]
        t2 = s[:, None, None] * (torch.diag_embed(pj) - pj[:, :, None] * pj[:, None, :])
        C[:, f:f + m, f:f + m] = t1 - t2
        f += m
    g_free = torch.einsum("bij,bi->bj", J, g)
    if H is None:
        return g_free, None
    H_free = torch.einsum("bij,bik,bkl->bjl", J, H, J) + C
    H_free = 0.5 * (H_free + H_free.transpose(1, 2))                         # symmetrize!, :452-457
    return g_free, H_free
