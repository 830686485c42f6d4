#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_v11.err
echo done
