#!/bin/bash
# quick check of a kernel change: march parity subset + gradient-only bench at three plan sizes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march or cuda_matches_oracle or deterministic or plan_device or task_mask or multi_field" > gpurun_out/pytest_march.log 2>&1; tail -2 gpurun_out/pytest_march.log
B="--steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
for f in 1 10; do timeout 600 python bench.py $B --fields $f > gpurun_out/q_f$f.json 2> gpurun_out/q_f$f.err; echo "f$f rc=$?"; done
python - <<'PY'
import json
for f in (1, 10):
    d = json.load(open(f"gpurun_out/q_f{f}.json")); r = d["roofline"]
    print(f"fields {f:2d}: {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel']} {r['kernel_ms_per_step']:.3f} ms  share {r['kernel_share_of_step']:.3f}  e2e {d['e2e']['value']/1e6:.3f} M  launches {d['gpu_launches']}")
PY
echo done
