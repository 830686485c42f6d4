#!/bin/bash
# quick check of a kernel change: unit-kernel parity subset + bench legs at the stripe and at one field
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unit_kernel or full_size_field or two_field or deterministic or task_mask or cuda_matches_oracle" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-render --no-single"
timeout 600 python bench.py $B > gpurun_out/q_f10.json 2> gpurun_out/q_f10.err; python tools/show_bench.py gpurun_out/q_f10.json | grep -v parity
timeout 600 python bench.py $B --fields 1 --sources-per-field 1250 --no-maximize > gpurun_out/q_f1.json 2> gpurun_out/q_f1.err; python tools/show_bench.py gpurun_out/q_f1.json | grep -v parity
python - <<'PY'
import json
for f in ("gpurun_out/q_f10.json", "gpurun_out/q_f1.json"):
    d = json.load(open(f))
    for nm, leg in (("grad", d), ("hess", d["hessian"])):
        r = leg["roofline"]
        print(f, nm, "step", round(leg["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in r["kernels"].items()}, "rest", round(leg["ms_per_step"] - r["pixel_kernels_ms_per_step"], 3))
PY
echo done
