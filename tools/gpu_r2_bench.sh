#!/bin/bash
# full GPU suite + bench with the reworked report
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_1gpu.err
python tools/show_bench.py gpurun_out/bench_1gpu.json
echo done
