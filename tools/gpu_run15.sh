#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_task.json 2> gpurun_out/bench_task.err
CELESTE_TASK_KERNEL=0 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_pix.json 2> gpurun_out/bench_pix.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 2 -c 1 -o gpurun_out/prof_task python tools/profile_step.py 10 1 3 > gpurun_out/ncu_task.log 2>&1
echo done
