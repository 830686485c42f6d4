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
        self._idx = {}

    def packed(self) -> torch.Tensor:
        """The prior as the flat 360-double array newton_step_kernel reads (csrc/maximize_kernels.cuh, PR_*):
        log pi_a[2], flux_mean[2], flux_var[2], log pi_k[2][8], mu[2][8][4], precision[2][8][4][4], logdet[2][8],
        radius mean, radius var."""
        rad = torch.tensor([self.rad_mean, self.rad_var], dtype=torch.float64, device=self.device)
        return torch.cat([self.log_is_star.reshape(-1), self.flux_mean.reshape(-1), self.flux_var.reshape(-1),
                          self.log_k.reshape(-1), self.mu2.reshape(-1), self.prec.reshape(-1), self.logdet.reshape(-1),
                          rad]).contiguous()

    def __call__(self, vp: torch.Tensor, order: int = 2):
        """vp: B x 44.  Returns (v [B], g [B x 44] or None, H [B x 44 x 44] or None) of subtract_kl.

        Per source type i the KL terms are  T_i = a_i (log a_i - log pi_i) + a_i G_i(r, s, c, v, k)  with
        G_i = Q(k) + R(r, s) + sum_d k_d D_d(c, v); its 19 x 19 Hessian over the local ordering
        (a, r, s, c[4], v[4], k[8]) is assembled from dense sub-blocks and scattered once."""
        B = vp.shape[0]
        dev, dt = vp.device, vp.dtype
        v = torch.zeros(B, dtype=dt, device=dev)
        g = torch.zeros((B, 44), dtype=dt, device=dev) if order >= 1 else None
        H = torch.zeros((B, 44, 44), dtype=dt, device=dev) if order >= 2 else None
        for i in range(2):
            idx = self._local_index(i, dev)
            th = vp[:, idx]                                   # B x 19
            a, r, s = th[:, 0], th[:, 1], th[:, 2]
            c, var, k = th[:, 3:7], th[:, 7:11], th[:, 11:19]
            la = torch.log(a) - self.log_is_star[i]
            lk = torch.log(k) - self.log_k[i]
            M, V = self.flux_mean[i], self.flux_var[i]
            R = 0.5 * (torch.log(V) - torch.log(s) + (s + (r - M) ** 2) / V - 1)
            P = self.prec[i]                                   # (8, 4, 4)
            delta = self.mu2[i][None] - c[:, None, :]          # (B, 8, 4)
            Pd = torch.einsum("djk,bdk->bdj", P, delta)
            diagP = torch.diagonal(P, dim1=1, dim2=2)          # (8, 4)
            D = 0.5 * ((diagP[None] * var[:, None, :]).sum(-1) - 4 + (delta * Pd).sum(-1) + self.logdet[i][None]
                       - torch.log(var).sum(-1, keepdim=True))   # (B, 8)
            G = (k * lk).sum(dim=1) + R + (k * D).sum(dim=1)
            v -= a * (la + G)
            if order >= 1:
                dDc, dDv = -Pd, 0.5 * (diagP[None] - 1 / var[:, None, :])      # (B, 8, 4)
                Gg = torch.empty((B, 18), dtype=dt, device=dev)
                Gg[:, 0] = (r - M) / V
                Gg[:, 1] = 0.5 * (1 / V - 1 / s)
                Gg[:, 2:6] = torch.einsum("bd,bdj->bj", k, dDc)
                Gg[:, 6:10] = torch.einsum("bd,bdj->bj", k, dDv)
                Gg[:, 10:18] = lk + 1 + D
                gl = torch.empty((B, 19), dtype=dt, device=dev)
                gl[:, 0] = la + 1 + G
                gl[:, 1:] = a[:, None] * Gg
                g[:, idx] -= gl
            if order >= 2:
                GH = torch.zeros((B, 18, 18), dtype=dt, device=dev)
                GH[:, 0, 0] = 1 / V
                GH[:, 1, 1] = 0.5 / s ** 2
                GH[:, 2:6, 2:6] = torch.einsum("bd,djk->bjk", k, P)
                GH[:, 6:10, 6:10] = torch.diag_embed(k.sum(dim=1, keepdim=True) * 0.5 / var ** 2)
                GH[:, 10:18, 10:18] = torch.diag_embed(1 / k)
                GH[:, 10:18, 2:6] = dDc
                GH[:, 2:6, 10:18] = dDc.transpose(1, 2)
                GH[:, 10:18, 6:10] = dDv
                GH[:, 6:10, 10:18] = dDv.transpose(1, 2)
                Hl = torch.empty((B, 19, 19), dtype=dt, device=dev)
                Hl[:, 0, 0] = 1 / a
                Hl[:, 0, 1:] = Gg
                Hl[:, 1:, 0] = Gg
                Hl[:, 1:, 1:] = a[:, None, None] * GH
                H[:, idx[:, None], idx[None, :]] -= Hl
        # --- source_e_log_prob
        x = vp[:, 5]
        v += -0.5 * (math.log(2 * math.pi) + math.log(self.rad_var) + (x - self.rad_mean) ** 2 / self.rad_var)
        if order >= 1:
            g[:, 5] += -(x - self.rad_mean) / self.rad_var
        if order >= 2:
            H[:, 5, 5] += -1 / self.rad_var
        return v, g, H

    def _local_index(self, i, dev):
        """canonical (0-based) ids of (a_i, flux_loc_i, flux_scale_i, color_mean[:, i], color_var[:, i], k[:, i])."""
        key = (i, str(dev))
        if key not in self._idx:
            ids = [26 + i, 6 + i, 8 + i] + list(range(10 + 4 * i, 14 + 4 * i)) + list(range(18 + 4 * i, 22 + 4 * i)) \
                + list(range(28 + 8 * i, 36 + 8 * i))
            self._idx[key] = torch.tensor(ids, dtype=torch.long, device=dev)
        return self._idx[key]
