"""Per-iteration trace of the device-resident Newton loop: active sources, plan time, step time (diagnostics)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import celeste_jl_b200 as cj
import bench
from celeste_jl_b200 import elbo_maximize as em

n_fields = int(sys.argv[1]) if len(sys.argv) > 1 else 10
stripe = bench.build_stripe(n_fields, 1000, device=None)
fields, rows, act, tf, vps = [], [], [], [], []
for fi, ds in enumerate(stripe):
    fields.append(cj.DeviceField(ds.images, ds.patches))
    r, a = ds.tasks()
    rows += r; act += a; tf += [fi] * len(r)
    vps.append(ds.vp_flat(r))
plan = cj.Plan(fields, rows, act, task_field=tf)
vp = np.concatenate(vps)
em.BatchMaximizer(plan, vp, include_kl=True, max_iters=2).run()
bm = em.BatchMaximizer(plan, vp, include_kl=True)
trace = []
orig_eval, orig_step = bm._evaluate_plan, bm._step
def ev():
    torch.cuda.synchronize(); t0 = time.perf_counter(); orig_eval(); torch.cuda.synchronize()
    trace.append(["plan", int(bm.mask.sum()), (time.perf_counter() - t0) * 1e3])
def stp(ph):
    torch.cuda.synchronize(); t0 = time.perf_counter(); orig_step(ph); torch.cuda.synchronize()
    trace.append(["step%d" % ph, int(bm.mask.sum()), (time.perf_counter() - t0) * 1e3])
bm._evaluate_plan, bm._step = ev, stp
t0 = time.perf_counter(); res = bm.run(); dt = time.perf_counter() - t0
print("total", dt, "steps", res.total_steps)
for t in trace:
    print(f"{t[0]:6s} active={t[1]:6d} {t[2]:8.3f} ms")
