"""How many host cores does the box really give us?  Times the oracle (CPU baseline) at several thread counts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from celeste_jl_b200 import synthetic
from celeste_jl_b200.flatten import csr_tasks
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
oracle_lib, lib, build = bench.load_oracle_for_timing()
ds = synthetic.FieldDataset(1000, H=2048, W=1489, seed=42, pixel_seed=1, device="cpu")
rows, act = ds.tasks()
tasks = [(r, a, np.stack([ds.vp[i - 1] for i in r], axis=1)) for r, a in zip(rows, act)]
csr = csr_tasks(tasks)
of = oracle_lib.OracleField(ds.images, ds.patches, lib=lib)
for nt in (1, 4, 8, 16, 32, 64, 128):
    n = min(len(tasks), max(64, nt * 8))
    sub = csr_tasks(tasks[:n])
    of.elbo_csr(*sub, mode=1, n_threads=nt)
    t0 = time.perf_counter(); of.elbo_csr(*sub, mode=1, n_threads=nt); dt = time.perf_counter() - t0
    print(f"threads {nt:4d}: {n} tasks in {dt*1e3:8.1f} ms -> {n/dt:9.1f} src/s", flush=True)
