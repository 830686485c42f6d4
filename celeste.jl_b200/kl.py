"""-KL(q || p) per source with gradient and Hessian, batched over sources -- row (f.3) of SURVEY.md section 8.

Mirror of src/deterministic_vi/elbo_kl.jl: `subtract_kl(vs)` (:143-154) =
  - kl_source_a (:94)  categorical KL of is_star vs prior.is_star
  - kl_source_k (:106) sum_i a_i * categorical KL of k[:, i] vs prior.k[:, i]
  - kl_source_r (:96)  sum_i a_i * gaussian_kl(flux_loc_i, flux_scale_i; prior.flux_mean_i, prior.flux_var_i)
  - kl_source_c (:116) sum_i a_i sum_d k[d, i] * diagmvn_mvn_kl(color_mean[:, i], color_var[:, i]; prior colour comp d)
  + source_e_log_prob (:132) log N(gal_radius_px; prior.gal_radius_px_mean, prior.gal_radius_px_var)
The reference differentiates this with ReverseDiff tapes and a ForwardDiff Jacobian (:163-193); here the
derivatives are closed forms evaluated for all sources at once (torch float64, CPU or CUDA).  It is added to
the likelihood SensitiveFloat exactly like subtract_kl_all_sources! (:214-225).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .model import load_prior


class KLTerm:
    def __init__(self, device="cpu", prior=None):
        p = prior or load_prior()
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64, device=device)
        self.log_is_star = torch.log(t(p.is_star))                   # (2,)
        self.flux_mean, self.flux_var = t(p.flux_mean), t(p.flux_var)
        self.log_k = torch.log(t(p.k)).T.contiguous()                # (2, 8)  [i, d]
        mu = t(np.transpose(p.color_mean, (2, 1, 0)))                # (2, 8, 4) [i, d, :]
        cov = t(np.transpose(p.color_cov, (3, 2, 0, 1)))             # (2, 8, 4, 4)
        self.mu2 = mu
        self.prec = torch.linalg.inv(cov)
        self.logdet = torch.logdet(cov)                              # (2, 8)
        self.rad_mean, self.rad_var = float(p.gal_radius_px_mean), float(p.gal_radius_px_var)
        self.device = device

    def __call__(self, vp: torch.Tensor, order: int = 2):
        """vp: B x 44.  Returns (v [B], g [B x 44] or None, H [B x 44 x 44] or None) of subtract_kl."""
        B = vp.shape[0]
        dev, dt = vp.device, vp.dtype
        v = torch.zeros(B, dtype=dt, device=dev)
        g = torch.zeros((B, 44), dtype=dt, device=dev) if order >= 1 else None
        H = torch.zeros((B, 44, 44), dtype=dt, device=dev) if order >= 2 else None

        def addH(i, j, val):          # symmetric scatter, i/j index tensors or ints broadcastable
            H[:, i, j] += val
            if not (isinstance(i, int) and isinstance(j, int) and i == j):
                H[:, j, i] += val

        for i in range(2):
            ia = 26 + i
            a = vp[:, ia]
            # --- kl_source_a
            la = torch.log(a) - self.log_is_star[i]
            v -= a * la
            if order >= 1:
                g[:, ia] -= la + 1
            if order >= 2:
                H[:, ia, ia] -= 1 / a
            # --- kl_source_k
            ik = torch.arange(28 + 8 * i, 36 + 8 * i, device=dev)
            k = vp[:, ik]
            lk = torch.log(k) - self.log_k[i]
            Q = (k * lk).sum(dim=1)
            v -= a * Q
            if order >= 1:
                g[:, ia] -= Q
                g[:, ik] -= a[:, None] * (lk + 1)
            if order >= 2:
                H[:, ia, ik] -= lk + 1
                H[:, ik, ia] -= lk + 1
                H[:, ik, ik] -= a[:, None] / k
            # --- kl_source_r
            ir, isc = 6 + i, 8 + i
            r, s = vp[:, ir], vp[:, isc]
            M, V = self.flux_mean[i], self.flux_var[i]
            R = 0.5 * (torch.log(V) - torch.log(s) + (s + (r - M) ** 2) / V - 1)
            dRr, dRs = (r - M) / V, 0.5 * (1 / V - 1 / s)
            v -= a * R
            if order >= 1:
                g[:, ia] -= R
                g[:, ir] -= a * dRr
                g[:, isc] -= a * dRs
            if order >= 2:
                addH(ia, ir, -dRr)
                addH(ia, isc, -dRs)
                H[:, ir, ir] -= a / V
                H[:, isc, isc] -= a * 0.5 / s ** 2
            # --- kl_source_c
            ic = torch.arange(10 + 4 * i, 14 + 4 * i, device=dev)
            iv = torch.arange(18 + 4 * i, 22 + 4 * i, device=dev)
            c, var = vp[:, ic], vp[:, iv]
            P = self.prec[i]                                              # (8, 4, 4)
            delta = self.mu2[i][None] - c[:, None, :]                     # (B, 8, 4)
            Pd = torch.einsum("djk,bdk->bdj", P, delta)                   # Lambda delta
            diagP = torch.diagonal(P, dim1=1, dim2=2)                     # (8, 4)
            D = 0.5 * ((diagP[None] * var[:, None, :]).sum(-1) - 4 + (delta * Pd).sum(-1) + self.logdet[i][None]
                       - torch.log(var).sum(-1, keepdim=True))           # (B, 8)
            dDc = -Pd                                                     # (B, 8, 4)
            dDv = 0.5 * (diagP[None] - 1 / var[:, None, :])               # (B, 8, 4)
            kD = (k * D).sum(dim=1)
            v -= a * kD
            if order >= 1:
                kdc = torch.einsum("bd,bdj->bj", k, dDc)
                kdv = torch.einsum("bd,bdj->bj", k, dDv)
                g[:, ia] -= kD
                g[:, ik] -= a[:, None] * D
                g[:, ic] -= a[:, None] * kdc
                g[:, iv] -= a[:, None] * kdv
            if order >= 2:
                H[:, ia, ik] -= D
                H[:, ik, ia] -= D
                H[:, ia, ic] -= kdc
                H[:, ic, ia] -= kdc
                H[:, ia, iv] -= kdv
                H[:, iv, ia] -= kdv
                kc = a[:, None, None] * dDc                               # (B, 8, 4): d2/dk dc
                kv = a[:, None, None] * dDv
                H[:, ik[:, None], ic[None, :]] -= kc
                H[:, ic[:, None], ik[None, :]] -= kc.transpose(1, 2)
                H[:, ik[:, None], iv[None, :]] -= kv
                H[:, iv[:, None], ik[None, :]] -= kv.transpose(1, 2)
                H[:, ic[:, None], ic[None, :]] -= a[:, None, None] * torch.einsum("bd,djk->bjk", k, P)
                H[:, iv, iv] -= a[:, None] * k.sum(dim=1, keepdim=True) * 0.5 / var ** 2
        # --- source_e_log_prob
        x = vp[:, 5]
        v += -0.5 * (math.log(2 * math.pi) + math.log(self.rad_var) + (x - self.rad_mean) ** 2 / self.rad_var)
        if order >= 1:
            g[:, 5] += -(x - self.rad_mean) / self.rad_var
        if order >= 2:
            H[:, 5, 5] += -1 / self.rad_var
        return v, g, H
