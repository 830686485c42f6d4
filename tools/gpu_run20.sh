#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/maximize_profile.py 10000 > gpurun_out/maxprof.log 2>&1; grep "^n=" gpurun_out/maxprof.log
for v in v1; do
CELESTE_CUDA_LIB=build_variants/$v.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
python tools/show_bench.py gpurun_out/bench_$v.json | head -1
done
echo done
