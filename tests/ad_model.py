"""Independent automatic-differentiation model of elbo_likelihood (test infrastructure).

The reference pins its hand-written derivatives by comparing them with ForwardDiff duals
pushed through the same value code (test/test_elbo.jl:223-301).  Here the same role is
played by torch.float64 autograd through a VALUE-ONLY, vectorised restatement of the
model written from the math, sharing no code with the oracle or
the CUDA library.  Gradient and Hessian of the oracle must match autograd of this.
"""
import math

import numpy as np
import torch

from celeste_jl_b200.model import galaxy_prototypes, ids

T = torch.float64


def _spline(coefs, x, y):
    n1, n2 = coefs.shape
    ix = torch.clamp(torch.floor(x.detach()), 1, n1 - 3).to(torch.long)
    iy = torch.clamp(torch.floor(y.detach()), 1, n2 - 3).to(torch.long)
    fx, fy = x - ix, y - iy

    def w(f):
        o = 1 - f
        return [o ** 3 / 6, 2 / 3 - f ** 2 + f ** 3 / 2, 2 / 3 - o ** 2 + o ** 3 / 2, f ** 3 / 6]
    wx, wy = w(fx), w(fy)
    out = 0
    for a in range(4):
        for b in range(4):
            out = out + wx[a] * wy[b] * coefs[ix - 1 + a, iy - 1 + b]
    return out


def _brightness(vs):
    """E_l[b, i], E_ll[b, i] as 5 x 2 tensors."""
    El = [[None] * 2 for _ in range(5)]
    Ell = [[None] * 2 for _ in range(5)]
    for i in range(2):
        r, s = vs[ids.flux_loc[i]], vs[ids.flux_scale[i]]
        c = [vs[ids.color_mean[m, i]] for m in range(4)]
        v = [vs[ids.color_var[m, i]] for m in range(4)]
        El[2][i] = torch.exp(r + s / 2)
        El[3][i] = El[2][i] * torch.exp(c[2] + v[2] / 2)
        El[4][i] = El[3][i] * torch.exp(c[3] + v[3] / 2)
        El[1][i] = El[2][i] * torch.exp(-c[1] + v[1] / 2)
        El[0][i] = El[1][i] * torch.exp(-c[0] + v[0] / 2)
        Ell[2][i] = torch.exp(2 * r + 2 * s)
        Ell[3][i] = Ell[2][i] * torch.exp(2 * c[2] + 2 * v[2])
        Ell[4][i] = Ell[3][i] * torch.exp(2 * c[3] + 2 * v[3])
        Ell[1][i] = Ell[2][i] * torch.exp(-2 * c[1] + 2 * v[1])
        Ell[0][i] = Ell[1][i] * torch.exp(-2 * c[0] + 2 * v[0])
    return El, Ell


def _densities(patch, vs, hh, ww):
    """(star density, galaxy density) of one source at float pixel coords hh, ww."""
    J = torch.tensor(patch.wcs_jacobian, dtype=T)
    wc = torch.tensor(patch.world_center, dtype=T)
    pc = torch.tensor(patch.pixel_center, dtype=T)
    m = J @ (vs[0:2] - wc) + pc
    coefs = torch.tensor(np.ascontiguousarray(patch.itp_coefs), dtype=T)
    y = _spline(coefs, hh - m[0] + 26, ww - m[1] + 26)
    f0 = torch.where(y < 0, 1e-3 * torch.exp(y), 1e-3 * (y + 1))
    theta, rho, phi, sig = vs[2], vs[3], vs[4], vs[5]
    cp, sp = torch.cos(phi), torch.sin(phi)
    x11 = sig ** 2 * (1 + (rho ** 2 - 1) * sp ** 2)
    x12 = -sig ** 2 * cp * sp * (rho ** 2 - 1)
    x22 = sig ** 2 * (1 + (rho ** 2 - 1) * cp ** 2)
    f1 = 0
    for i, (eta, nu) in enumerate(galaxy_prototypes):
        th = theta if i == 0 else 1 - theta
        for j in range(len(eta)):
            for k in patch.psf:
                t = torch.tensor(np.asarray(k.tauBar), dtype=T)
                S = torch.stack([torch.stack([t[0, 0] + nu[j] * x11, t[0, 1] + nu[j] * x12]),
                                 torch.stack([t[1, 0] + nu[j] * x12, t[1, 1] + nu[j] * x22])])
                det = S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]
                L = torch.linalg.inv(S)
                dx = hh - (k.xiBar[0] + m[0])
                dy = ww - (k.xiBar[1] + m[1])
                q = L[0, 0] * dx * dx + (L[0, 1] + L[1, 0]) * dx * dy + L[1, 1] * dy * dy
                f1 = f1 + th * k.alphaBar * eta[j] / (2 * math.pi * torch.sqrt(det)) * torch.exp(-0.5 * q)
    return f0, f1


def elbo_value(images, patches, vp_list, active_sources):
    """Scalar torch value of elbo_likelihood; vp_list: list of 44-tensors (float64)."""
    S, N = patches.shape
    total = torch.zeros((), dtype=T)
    bright = [_brightness(vs) for vs in vp_list]
    for n in range(N):
        img = images[n]
        H, W = img.H, img.W
        visit = np.zeros((H, W), dtype=bool)
        for s1 in active_sources:
            p = patches[s1 - 1, n]
            H2, W2 = p.active_pixel_bitmap.shape
            if H2 == 0 or W2 == 0:
                continue
            o = p.bitmap_offset
            visit[o[0]:o[0] + H2, o[1]:o[1] + W2] |= p.active_pixel_bitmap
        visit &= ~np.isnan(img.pixels)
        hs, ws = np.nonzero(visit)
        if len(hs) == 0:
            continue
        hh = torch.tensor(hs + 1.0, dtype=T)
        ww = torch.tensor(ws + 1.0, dtype=T)
        E = torch.tensor(img.sky[hs, ws].astype(np.float64))
        V = torch.zeros_like(E)
        b = img.b - 1
        for s in range(S):
            p = patches[s, n]
            H2, W2 = p.active_pixel_bitmap.shape
            o = p.bitmap_offset
            h2, w2 = hs - o[0], ws - o[1]                     # 0-based local
            inside = (h2 >= 0) & (h2 < H2) & (w2 >= 0) & (w2 < W2 - 1)   # strict last column (elbo_objective.jl:349)
            cover = np.zeros(len(hs), dtype=bool)
            cover[inside] = p.active_pixel_bitmap[h2[inside], w2[inside]]
            if not cover.any():
                continue
            cm = torch.tensor(cover)
            vs = vp_list[s]
            f0, f1 = _densities(p, vs, hh[cm], ww[cm])
            El, Ell = bright[s]
            a = [vs[ids.is_star[0]], vs[ids.is_star[1]]]
            Es = a[0] * El[b][0] * f0 + a[1] * El[b][1] * f1
            E2s = a[0] * Ell[b][0] * f0 ** 2 + a[1] * Ell[b][1] * f1 ** 2
            E = E.index_put((torch.nonzero(cm)[:, 0],), Es, accumulate=True)
            V = V.index_put((torch.nonzero(cm)[:, 0],), E2s - Es ** 2, accumulate=True)
        x = torch.tensor(img.pixels[hs, ws].astype(np.float64))
        iota = torch.tensor(img.nelec_per_nmgy[hs].astype(np.float64))
        logiota = torch.tensor(np.log(img.nelec_per_nmgy[hs]).astype(np.float64))    # Float32 log
        total = total + (x * (logiota + torch.log(E) - V / (2 * E ** 2)) - iota * E - torch.lgamma(x + 1)).sum()
    return total


def elbo_ad(images, patches, vp, active_sources, hessian=True):
    """value, gradient (44 x Sa), Hessian (44 Sa x 44 Sa) by autograd."""
    Sa = len(active_sources)
    x0 = torch.tensor(np.concatenate([vp[s - 1] for s in active_sources]), dtype=T, requires_grad=True)

    def f(x):
        vl = [torch.tensor(v, dtype=T) for v in vp]
        for k, s in enumerate(active_sources):
            vl[s - 1] = x[44 * k:44 * (k + 1)]
        return elbo_value(images, patches, vl, active_sources)
    val = f(x0)
    g, = torch.autograd.grad(val, x0, create_graph=hessian)
    Hm = None
    if hessian:
        rows = []
        for i in range(44 * Sa):
            if g[i].requires_grad:
                gi, = torch.autograd.grad(g[i], x0, retain_graph=True, allow_unused=True)
                rows.append(torch.zeros(44 * Sa, dtype=T) if gi is None else gi)
            else:
                rows.append(torch.zeros(44 * Sa, dtype=T))
        Hm = torch.stack(rows).detach().numpy()
    return float(val.detach()), g.detach().numpy().reshape(Sa, 44).T, Hm


def render_value(images, patches, vp_list):
    """Independent restatement of fill_celeste_expectation! (bin/write_celeste_expectation.jl:111-156): per image
    the H x W array of E_G - sky, i.e. the summed expected source flux (nmgy) with the ELBO's in-patch test."""
    S, N = patches.shape
    bright = [_brightness(torch.tensor(vs, dtype=T)) for vs in vp_list]
    outs = []
    for n in range(N):
        img = images[n]
        out = np.zeros((img.H, img.W))
        b = img.b - 1
        for s in range(S):
            p = patches[s, n]
            H2, W2 = p.active_pixel_bitmap.shape
            if H2 == 0 or W2 <= 1:
                continue
            o = p.bitmap_offset
            h2, w2 = np.nonzero(p.active_pixel_bitmap[:, :W2 - 1])          # strict last column
            hs, ws = h2 + o[0], w2 + o[1]
            ok = (hs >= 0) & (hs < img.H) & (ws >= 0) & (ws < img.W)
            hs, ws = hs[ok], ws[ok]
            if len(hs) == 0:
                continue
            vs = torch.tensor(vp_list[s], dtype=T)
            f0, f1 = _densities(p, vs, torch.tensor(hs + 1.0, dtype=T), torch.tensor(ws + 1.0, dtype=T))
            El, _ = bright[s]
            Es = vs[ids.is_star[0]] * El[b][0] * f0 + vs[ids.is_star[1]] * El[b][1] * f1
            np.add.at(out, (hs, ws), Es.numpy())
        outs.append(out)
    return outs
