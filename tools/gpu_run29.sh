#!/bin/bash
# table-driven log in the pixel term; loads hoisted above the exchange (variant)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
V=$PWD/celeste.jl_b200/variants
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march or cuda_matches_oracle or deterministic or full_size" > gpurun_out/pytest_march.log 2>&1; tail -2 gpurun_out/pytest_march.log
run fastlog X=1
run hoist CELESTE_CUDA_LIB=$V/libceleste_cuda_hoist.so
python - <<'PY'
import json
for n in ("fastlog", "hoist"):
    try:
        d = json.load(open(f"gpurun_out/ab_{n}.json"))
        r = d["roofline"]
        print(f"{n:9s} {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel']} {r['kernel_ms_per_step']:.3f} ms  frac {r['frac']:.3f}  e2e {d['e2e']['value']/1e6:.3f} M  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
echo done
