#!/bin/bash
# round 2, first call: HEAD (round-1 kernels) + the new full-size parity tests, bench, sanitizers on HEAD
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_1gpu.err
python tools/show_bench.py gpurun_out/bench_1gpu.json
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck.txt
echo done
