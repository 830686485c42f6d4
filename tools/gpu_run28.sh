#!/bin/bash
# march_kernel with loads after the mixture sums: A/B (168 vs 128 vs 254 registers) on 1 GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
V=$PWD/celeste.jl_b200/variants
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march or cuda_matches_oracle or deterministic" > gpurun_out/pytest_march.log 2>&1; tail -2 gpurun_out/pytest_march.log
run march X=1
run b4 CELESTE_CUDA_LIB=$V/libceleste_cuda_b4.so
run b2 CELESTE_CUDA_LIB=$V/libceleste_cuda_b2.so
python - <<'PY'
import json
for n in ("march", "b4", "b2"):
    try:
        d = json.load(open(f"gpurun_out/ab_{n}.json"))
        r = d["roofline"]
        print(f"{n:9s} {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel']} {r['kernel_ms_per_step']:.3f} ms  frac {r['frac']:.3f}  e2e {d['e2e']['value']/1e6:.3f} M  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/prof_march python tools/profile_step.py 10 1 3 > gpurun_out/ncu_march.log 2>&1; echo "ncu march rc=$?"
echo done
