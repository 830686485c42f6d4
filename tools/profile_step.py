"""Driver for ncu captures: the SAME workload as bench.py (multi-field plan over the synthetic stripe),
a few evaluations per mode, no torch kernels between them.
    python tools/profile_step.py <fields> <modes e.g. 1,2> <reps>"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celeste_jl_b200 as cj  # noqa: E402
import bench  # noqa: E402

n_fields = int(sys.argv[1]) if len(sys.argv) > 1 else 10
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
stripe = bench.build_stripe(n_fields, 1000, device=None)
fields, rows, act, tf, vps = [], [], [], [], []
for fi, ds in enumerate(stripe):
    fields.append(cj.DeviceField(ds.images, ds.patches))
    r, a = ds.tasks()
    rows += r
    act += a
    tf += [fi] * len(r)
    vps.append(ds.vp_flat(r))
plan = cj.Plan(fields, rows, act, task_field=tf)
vp = np.concatenate(vps)
for mode in modes:
    for _ in range(reps):
        out = plan.run_host(vp, mode)
    print("mode", mode, "tasks", plan.n_tasks, "v[0]", out["v"][0], "visits", out["counters"].reshape(-1, 2).sum(axis=0))
