#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:celeste -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --fields 2 --no-cpu-baseline --no-hessian > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_grad_r01 python tools/profile_step.py 1000 1 3 > gpurun_out/ncu_grad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_hess_r01 python tools/profile_step.py 1000 2 3 > gpurun_out/ncu_hess.log 2>&1
echo done
