"""CPU checker for the optimiser rows (f.1-f.3): per-source, plain restatements (test infrastructure).

  * kl_value / kl_ad            src/deterministic_vi/elbo_kl.jl:25-154, derivatives by torch autograd
                                (the reference uses ReverseDiff + ForwardDiff, :163-193)
  * to_bound_single / transform_ad   ConstraintTransforms.jl:67-123,189-216, derivatives by autograd
                                (the reference uses nested ForwardDiff Jacobians, :360-367)
  * maximize_single             ElboMaximize.maximize! (:228-242) with the Newton trust-region of Optim.jl
                                restated per source in numpy (Nocedal & Wright Alg. 4.1 / exact subproblem)
"""
import math

import numpy as np
import torch

from celeste_jl_b200.model import load_prior, ids

T = torch.float64
_prior = load_prior()


def kl_value(vs):
    """subtract_kl (elbo_kl.jl:143-154) of one 44-vector (torch)."""
    p = _prior
    kl = torch.zeros((), dtype=T)
    a = vs[ids.is_star]
    kl = kl - sum(a[i] * (torch.log(a[i]) - math.log(p.is_star[i])) for i in range(2))        # kl_source_a
    for i in range(2):                                                                       # kl_source_k
        k = vs[ids.k[:, i]]
        kl = kl - a[i] * sum(k[d] * (torch.log(k[d]) - math.log(p.k[d, i])) for d in range(8))
    for i in range(2):                                                                       # kl_source_r
        mu1, var1 = vs[ids.flux_loc[i]], vs[ids.flux_scale[i]]
        mu2, var2 = p.flux_mean[i], p.flux_var[i]
        kl = kl - a[i] * .5 * (math.log(var2) - torch.log(var1) + (var1 + (mu1 - mu2) ** 2) / var2 - 1)
    for i in range(2):                                                                       # kl_source_c
        mu1, var1 = vs[ids.color_mean[:, i]], vs[ids.color_var[:, i]]
        for d in range(8):
            mu2 = torch.tensor(p.color_mean[:, d, i], dtype=T)
            S2 = torch.tensor(p.color_cov[:, :, d, i], dtype=T)
            inv = torch.linalg.inv(S2)
            diff = mu2 - mu1
            t = (torch.diagonal(inv) * var1).sum() - 4 + diff @ inv @ diff + torch.logdet(S2) - torch.log(var1).sum()
            kl = kl - a[i] * vs[ids.k[d, i]] * 0.5 * t
    x = vs[ids.gal_radius_px]                                                                # source_e_log_prob
    kl = kl + -0.5 * (math.log(2 * math.pi) + math.log(p.gal_radius_px_var) + (x - p.gal_radius_px_mean) ** 2 / p.gal_radius_px_var)
    return kl


def kl_ad(vs_np):
    x = torch.tensor(vs_np, dtype=T, requires_grad=True)
    v = kl_value(x)
    g, = torch.autograd.grad(v, x, create_graph=True)
    H = torch.stack([torch.autograd.grad(g[i], x, retain_graph=True)[0] for i in range(44)])
    return float(v.detach()), g.detach().numpy(), H.detach().numpy()


BOXES = [(0, 1, None), (1, 2, None), (2, 3, (1e-2, 0.99)), (3, 4, (1e-2, 0.99)), (4, 5, (-10.0, 10.0)),
         (5, 6, (0.10, 70.0)), (6, 8, (-1.0, 10.0)), (8, 10, (1e-4, 0.10)), (10, 14, (-10.0, 10.0)),
         (14, 18, (-10.0, 10.0)), (18, 22, (1e-4, 1.0)), (22, 26, (1e-4, 1.0))]
SIMPLEXES = [(26, 2, 0.005), (28, 8, 0.01 / 8), (36, 8, 0.01 / 8)]


def bounds_single(vs0, loc_width):
    lo, hi = np.zeros(26), np.zeros(26)
    for a, b, lu in BOXES:
        if lu is None:
            lo[a:b], hi[a:b] = vs0[a:b] - loc_width, vs0[a:b] + loc_width
        else:
            lo[a:b], hi[a:b] = lu
    return lo, hi


def to_bound_single(free, lo, hi):
    """to_bound! for one source (torch, differentiable)."""
    lo_t, hi_t = torch.as_tensor(lo, dtype=T), torch.as_tensor(hi, dtype=T)
    out = [1.0 / (1.0 + torch.exp(-free[:26])) * (hi_t - lo_t) + lo_t]
    f = 26
    for first, n, lower in SIMPLEXES:
        z = free[f:f + n - 1]
        m = z.max().detach()
        e = torch.exp(z - m)
        tot = torch.exp(-m) + e.sum()
        out.append((1 - n * lower) * (e / tot) + lower)
        out.append(((1 - n * lower) * (torch.exp(-m) / tot) + lower).reshape(1))
        f += n - 1
    return torch.cat(out)


def to_free_single(vs, lo, hi):
    free = np.zeros(41)
    u = (vs[:26] - lo) / (hi - lo)
    free[:26] = -np.log(1.0 / u - 1)
    f = 26
    for first, n, lower in SIMPLEXES:
        un = (vs[first:first + n] - lower) / (1 - n * lower)
        free[f:f + n - 1] = np.log(un[:-1]) - np.log(un[-1])
        f += n - 1
    return free


def propagate_ad(free_np, lo, hi, g, H):
    """propagate_derivatives! by autograd: J' g, J' H J + sum_i g_i d2 b_i, symmetrised."""
    x = torch.tensor(free_np, dtype=T, requires_grad=True)
    J = torch.autograd.functional.jacobian(lambda t: to_bound_single(t, lo, hi), x)            # 44 x 41
    gt = torch.tensor(g, dtype=T)
    C = torch.autograd.functional.hessian(lambda t: (to_bound_single(t, lo, hi) * gt).sum(), x)
    J = J.numpy()
    gf = J.T @ g
    Hf = J.T @ H @ J + C.numpy()
    return gf, 0.5 * (Hf + Hf.T)


def tr_subproblem_single(g, H, delta):
    """Exact trust-region subproblem (eigenbasis, bisection on the secular equation; hard case included)."""
    ev, Q = np.linalg.eigh(H)
    qg = Q.T @ g
    if ev[0] >= 1e-8:
        s = -Q @ (qg / ev)
        if s @ s <= delta * delta:
            return s, g @ s + 0.5 * s @ H @ s, True
    lam_lb = max(0.0, -ev[0])

    def pn(lam):
        return np.sqrt(((qg / (ev + lam)) ** 2).sum())
    at_min = np.abs(ev - ev[0]) <= 1e-12 * (1 + np.abs(ev))
    if ev[0] <= 1e-8 and np.all(np.abs(qg[at_min]) <= 1e-12 * (np.abs(qg).max() + 1e-300)):
        coef = np.where(at_min, 0.0, -qg / np.where(at_min, 1.0, ev - ev[0]))
        if coef @ coef <= delta * delta:
            coef[0] = np.sqrt(delta * delta - coef @ coef)
            s = Q @ coef
            return s, g @ s + 0.5 * s @ H @ s, False
    lo_l = lam_lb + 1e-12 * (1 + np.abs(ev).max())
    hi_l = lo_l + 1.0
    while pn(hi_l) > delta:
        hi_l = lo_l + 2 * (hi_l - lo_l)
    for _ in range(200):
        mid = 0.5 * (lo_l + hi_l)
        if pn(mid) > delta:
            lo_l = mid
        else:
            hi_l = mid
    lam = 0.5 * (lo_l + hi_l)
    s = -Q @ (qg / (ev + lam))
    return s, g @ s + 0.5 * s @ H @ s, False


def maximize_single(elbo_fn, vs0, include_kl=True, loc_width=1e-4, max_iters=50):
    """maximize! for one source. elbo_fn(vs44) -> (v, g44, H44) of the LIKELIHOOD (oracle)."""
    lo, hi = bounds_single(vs0, loc_width)

    def evaluate(x):
        b = to_bound_single(torch.tensor(x, dtype=T), lo, hi).numpy()
        v, g, H = elbo_fn(b)
        if include_kl:
            kv, kg, kH = kl_ad(b)
            v, g, H = v + kv, g + kg, H + kH
        gf, Hf = propagate_ad(x, lo, hi, g, H)
        return -v, -gf, -Hf, b
    eps = np.finfo(float).eps
    vs = vs0.copy()
    vs[:26] = np.where((vs[:26] > lo) & (vs[:26] < hi), vs[:26],
                       np.maximum(np.minimum(vs[:26], np.nextafter(hi, lo)), np.nextafter(lo, hi)))
    x = to_free_single(vs, lo, hi)
    f, g, H, b = evaluate(x)
    delta, it, calls = 1.0, 0, 1
    if np.abs(g).max() < 1e-8:
        return b, -f, it, calls, True
    converged = False
    while it < max_iters:
        it += 1
        s, m, interior = tr_subproblem_single(g, H, delta)
        fn, gn, Hn, bn = evaluate(x + s)
        calls += 1
        if abs(m) <= eps:
            rho = 1.0
        elif m > 0 or not np.isfinite(fn):
            rho = 0.25 - 1.0
        else:
            rho = (f - fn) / (-m)
        if rho < 0.25:
            delta *= 0.25
        elif rho > 0.75 and not interior:
            delta = min(2 * delta, 1e9)
        if rho > 0.1:
            xc = np.abs(s).max() < 1e-7
            fc = abs(fn - f) <= 1e-6 * abs(fn)
            gc = np.abs(gn).max() < 1e-8
            x, f, g, H, b = x + s, fn, gn, Hn, bn
            if xc or fc or gc:
                converged = True
                break
        if delta < 1e-14:
            break
    return b, -f, it, calls, converged
