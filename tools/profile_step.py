"""Minimal driver for ncu: one 1000-source field, a few evaluations per mode (no torch kernels in between)."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import celeste_jl_b200 as cj  # noqa: E402
from celeste_jl_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ds = synthetic.FieldDataset(n, H=2048, W=1489, seed=42, pixel_seed=1)
field = cj.DeviceField(ds.images, ds.patches)
rows, act = ds.tasks()
plan = field.make_plan(rows, act)
vp = ds.vp_flat(rows)
for mode in modes:
    for _ in range(reps):
        out = plan.run_host(vp, mode)
    print("mode", mode, "v[0]", out["v"][0], "visits", out["counters"].sum(axis=0) if out["counters"].ndim > 1 else out["counters"].reshape(-1, 2).sum(axis=0))
