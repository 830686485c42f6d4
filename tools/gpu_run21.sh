#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_v12.err
python tools/show_bench.py gpurun_out/bench_v12.json
python -c "
import json; d=json.load(open('gpurun_out/bench_v12.json')); print(d.get('maximize')); print(d.get('render'))"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck.txt
echo done
