"""Diagnostics for the config-5 validation test: GPU BatchMaximizer vs the oracle-driven CPU run on 200 sources."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import celeste_jl_b200 as cj
from celeste_jl_b200 import deterministic_vi as dvi, elbo_maximize as em, synthetic
from celeste_jl_b200.model import ids
from test_maximize import OracleRunner, PlanLike
ds = synthetic.FieldDataset(1000, H=2048, W=1489, seed=42, pixel_seed=1)
targets = list(range(0, 1000, 5))
rows, act = ds.tasks(targets)
vps = []
for r in rows:
    vps.append(dvi.generic_init_source(ds.catalog[r[0] - 1].pos))
    vps += [dvi.catalog_init_source(ds.catalog[k - 1]) for k in r[1:]]
vp = np.concatenate(vps)
field = cj.DeviceField(ds.images, ds.patches)
plan = cj.Plan(field, rows, act)
gpu = em.BatchMaximizer(plan, vp, include_kl=True).run()
unf = em.BatchMaximizer(plan, vp, include_kl=True, fused=False).run()
pl = PlanLike(rows, act)
cpu = em.BatchMaximizer(pl, vp, include_kl=True, device="cpu", runner=OracleRunner(ds.images, ds.patches, pl)).run()
for name, a, b in (("gpu fused vs cpu", gpu, cpu), ("gpu unfused vs cpu", unf, cpu), ("gpu fused vs unfused", gpu, unf)):
    di = np.abs(a.iterations - b.iterations)
    conv = a.converged & b.converged
    rv = np.abs(a.value - b.value) / np.abs(b.value)
    dvp = np.abs(a.vp - b.vp) / np.maximum(np.abs(b.vp), 1e-3)
    print(name, "conv", a.converged.mean(), b.converged.mean(), "same iters", (di == 0).mean(), "|diters|<=2", (di <= 2).mean(),
          "max diters", di.max(), "rel value: median", np.median(rv[conv]), "p95", np.percentile(rv[conv], 95), "max", rv[conv].max(),
          "vp rel: p95", np.percentile(dvp[conv].max(axis=1), 95), "max", dvp[conv].max())
n_ok = 0
for k, t in enumerate(targets):
    ce = ds.catalog[t]
    flux_r = (ce.star_fluxes if ce.is_star else ce.gal_fluxes)[2]
    if len(rows[k]) > 1 or flux_r < 30.0 or not gpu.converged[k]:
        continue
    vs = gpu.vp[k]
    a_true = 0 if ce.is_star else 1
    bright = np.exp(vs[ids.flux_loc[a_true]] + 0.5 * vs[ids.flux_scale[a_true]])
    print("src", t, "star" if ce.is_star else "gal", "flux", round(flux_r, 1), "a", vs[ids.is_star].round(3), "dpos", (vs[:2] - ce.pos).round(3),
          "bright ratio", round(bright / flux_r, 3))
    n_ok += 1
print("checked", n_ok)
