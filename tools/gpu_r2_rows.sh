#!/bin/bash
# fixed per-patch row cut of the units: sweep CELESTE_UNIT_ROWS; config-5 validation test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 10 --warmup 3 --no-cpu-baseline --no-render"
for r in 0 8 12 16 24 32; do
  CELESTE_UNIT_ROWS=$r timeout 600 python bench.py $B > gpurun_out/rows_$r.json 2> gpurun_out/rows_$r.err
  echo "rows $r"; python tools/show_bench.py gpurun_out/rows_$r.json | grep -E "grad|hess |maximize|single"
done
for f in 1; do
  for r in 0 16; do
  CELESTE_UNIT_ROWS=$r timeout 600 python bench.py $B --fields 1 --no-maximize --no-single > gpurun_out/rows_f1_$r.json 2> gpurun_out/rows_f1_$r.err
  echo "fields 1 rows $r"; python tools/show_bench.py gpurun_out/rows_f1_$r.json | grep -E "grad|hess "
  done
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config5 or full_size or deterministic" > gpurun_out/pytest_c5.log 2>&1; tail -5 gpurun_out/pytest_c5.log
echo done
