"""Which tasks / entries carry the largest Hessian error of the bench's parity sample?  (diagnostics)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import celeste_jl_b200 as cj
import oracle_lib
from celeste_jl_b200 import synthetic
ds = synthetic.FieldDataset(1000, H=2048, W=1489, seed=42, pixel_seed=1)
field = cj.DeviceField(ds.images, ds.patches)
rows, act = ds.tasks()
tasks = [(r, a, np.stack([ds.vp[i - 1] for i in r], axis=1)) for r, a in zip(rows, act)]
pick = np.random.default_rng(5).choice(len(tasks), 200, replace=False)
sub = [tasks[i] for i in pick]
ref = oracle_lib.OracleField(ds.images, ds.patches).elbo_batch(sub, mode=2, n_threads=16)
got = field.elbo_batch(sub, mode=2)
n = len(sub)
r, g = ref["h"].reshape(n, 44, 44), got["h"].reshape(n, 44, 44)
sc = np.abs(r).reshape(n, -1).max(axis=1)[:, None, None]
err6 = np.abs(r - g) / np.maximum(np.abs(r), sc * 1e-6)
err3 = np.abs(r - g) / np.maximum(np.abs(r), sc * 1e-3)
print(os.environ.get("CELESTE_CUDA_LIB", "default lib"), os.environ.get("CELESTE_EPILOGUE", ""), os.environ.get("CELESTE_HESS_KERNEL", ""),
      "max err (1e-6 floor)", err6.max(), "(1e-3 floor)", err3.max(), "median of per-task max", np.median(err6.reshape(n, -1).max(axis=1)))
t, i, j = np.unravel_index(np.argmax(err6), err6.shape)
print("  worst: task", pick[t], "entry", (i, j), "ref", r[t, i, j], "got", g[t, i, j], "row max", sc[t, 0, 0], "abs err / row max", abs(r[t, i, j] - g[t, i, j]) / sc[t, 0, 0])
