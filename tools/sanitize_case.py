"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel of the library once per mode."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import celeste_jl_b200 as cj
from celeste_jl_b200 import elbo_maximize as em
import cases
for name in ("two_body", "clipped_and_empty", "psf_k3", "crowded", "masked", "wide_patch", "seven_images", "sharp_psf"):
    images, patches, tasks = cases.get(name)
    f = cj.DeviceField(images, patches)
    for mode in (0, 1, 2):
        out = f.elbo_batch(tasks, mode=mode)
    print(name, out["v"][:2])
images, patches, tasks = cases.get("config2")
vpm = np.stack([tasks[0][2][:, 0], tasks[1][2][:, 0], tasks[2][2][:, 0]], axis=1)
out = cj.DeviceField(images, patches).elbo_batch([([1, 2, 3], [3, 1, 2], vpm)], mode=2)
print("Sa=3", out["v"], out["counters"])
g = torch.randn(4, 41, dtype=torch.float64, device="cuda")
A = torch.randn(4, 41, 41, dtype=torch.float64, device="cuda")
s, m, interior = em.solve_tr_subproblem(g, A + A.transpose(1, 2), torch.ones(4, dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
print("tr", m.cpu().numpy())
# row f.4 render kernel and the fused Newton step (f.1-f.3)
images, patches, tasks = cases.get("clipped_and_empty")
vp = cases.all_vp(patches, tasks)
r = cj.DeviceField(images, patches).render_expectation(np.arange(1, vp.shape[1] + 1), vp)
print("render", [float(a.sum()) for a in r])
from celeste_jl_b200 import synthetic
ds = synthetic.FieldDataset(6, H=90, W=80, seed=3, device="cpu")
rows, act = ds.tasks()
plan = cj.Plan(cj.DeviceField(ds.images, ds.patches), rows, act)
res = em.BatchMaximizer(plan, ds.vp_flat(rows), include_kl=True, max_iters=3).run()
print("newton", res.value[:3], res.iterations)
# one full-size 1000-source plan (configs[2]); the dataset is rendered on the CPU so that only this library's kernels are instrumented
ds = synthetic.FieldDataset(1000, H=2048, W=1489, seed=42, device="cpu")
rows, act = ds.tasks()
plan = cj.Plan(cj.DeviceField(ds.images, ds.patches), rows, act)
vpf = ds.vp_flat(rows)
for mode in (0, 1, 2):
    out = plan.run_host(vpf, mode)
    print("field1000 mode", mode, float(out["v"].sum()), int(out["counters"].sum()), int(out["flags"].sum()))
