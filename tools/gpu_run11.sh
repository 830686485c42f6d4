#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/libceleste_cuda_g5.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize > gpurun_out/bench_v7_g5.json 2> gpurun_out/bench_v7_g5.err
echo done
