#!/bin/bash
# march_kernel v2b (host-built block headers, block-level segment choice, task-space accumulation): A/B + ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march or cuda_matches_oracle or deterministic or full_size or plan_device or task_mask" > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_march.log; tail -4 gpurun_out/pytest_march.log
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
timeout 600 python bench.py $B > gpurun_out/ab_march.json 2> gpurun_out/ab_march.err; echo "march rc=$?"
CELESTE_MARCH_SPLIT=6000 timeout 600 python bench.py $B > gpurun_out/ab_split6000.json 2> gpurun_out/ab_split6000.err; echo "split rc=$?"
for v in b2 b4 t64b6 t64b8 seg24; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/variants/libceleste_cuda_$v.so timeout 600 python bench.py $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo "$v rc=$?"
done
python - <<'PY'
import json
for n in ("march", "split6000", "b2", "b4", "t64b6", "t64b8", "seg24"):
    try:
        d = json.load(open(f"gpurun_out/ab_{n}.json"))
        r = d["roofline"]
        print(f"{n:9s} {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel']} {r['kernel_ms_per_step']:.3f} ms  frac {r['frac']:.3f}  e2e {d['e2e']['value']/1e6:.3f} M  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/prof_march python tools/profile_step.py 10 1 3 > gpurun_out/ncu_march.log 2>&1; echo "ncu rc=$?"
echo done
