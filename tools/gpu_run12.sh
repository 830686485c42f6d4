#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck_r01.txt 2>&1; echo "memcheck rc=$?"
echo done
