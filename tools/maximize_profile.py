"""Where does a lock-step Newton iteration spend its time?  (diagnostics)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import celeste_jl_b200 as cj
from celeste_jl_b200 import synthetic, elbo_maximize as em, deterministic_vi as dvi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ds = synthetic.FieldDataset(n, H=2048, W=1489, seed=42, pixel_seed=1)
field = cj.DeviceField(ds.images, ds.patches)
rows, act = ds.tasks()
plan = cj.Plan(field, rows, act)
vps = []
for r in rows:
    vps.append(dvi.generic_init_source(ds.catalog[r[0] - 1].pos))
    vps += [dvi.catalog_init_source(ds.catalog[k - 1]) for k in r[1:]]
vp = np.concatenate(vps)
for prof, fused in ((None, True), ({}, True), (None, False), ({}, False)):
    bm = em.BatchMaximizer(plan, vp, include_kl=True, fused=fused)
    bm.profile = prof
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = bm.run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n={n} total {dt:.3f}s steps {res.total_steps} mean iters {res.iterations.mean():.1f} converged {res.converged.mean():.3f} "
          f"src/s {n/dt:.1f} fused={fused}", prof)
