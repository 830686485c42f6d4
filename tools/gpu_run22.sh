#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_v13.err
python tools/show_bench.py gpurun_out/bench_v13.json
python -c "
import json; d=json.load(open('gpurun_out/bench_v13.json')); print(d.get('maximize')); print(d.get('render'))"
echo done
