#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle" > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_task2.json 2> gpurun_out/bench_task2.err
CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/libceleste_cuda_g5.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_task2_g5.json 2> gpurun_out/bench_task2_g5.err
echo done
