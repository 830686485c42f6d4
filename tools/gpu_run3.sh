#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
for V in b53 b62 b31; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/libceleste_cuda_$V.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_$V.json 2> gpurun_out/bench_v2_$V.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_grad_v2 python tools/profile_step.py 1000 1 3 > gpurun_out/ncu_grad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_hess_v2 python tools/profile_step.py 1000 2 3 > gpurun_out/ncu_hess.log 2>&1
echo done
