#!/bin/bash
# first GPU pass: smoke, parity tests, bench (default + thread variants), ncu launch list
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_t128.json 2> gpurun_out/bench_t128.err; echo "bench rc=$?"
for T in 32 64; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/libceleste_cuda_t$T.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t$T.json 2> gpurun_out/bench_t$T.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --fields 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo done
