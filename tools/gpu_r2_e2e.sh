#!/bin/bash
# e2e of a 1250-source plan (one rank's share of an 8-GPU run) and of the stripe, then the GPU test-suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-single"
timeout 600 python bench.py $B --fields 1 --sources-per-field 1250 > gpurun_out/e2e_small.json 2> gpurun_out/e2e_small.err
timeout 600 python bench.py $B > gpurun_out/e2e_big.json 2> gpurun_out/e2e_big.err
python - <<'PY'
import json
for n in ["small", "big"]:
    d = json.loads(open(f"gpurun_out/e2e_{n}.json").read().strip().splitlines()[-1]); h = d["hessian"]
    print(n, "grad", round(d["value"] / 1e6, 3), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"] / 1e6, 3), "| hess", round(h["value"] / 1e6, 3), "e2e", {k: (round(v / 1e6, 3) if isinstance(v, float) and v > 1e4 else v) for k, v in h["e2e"].items() if "value" in k})
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
