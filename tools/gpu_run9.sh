#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "joint" > gpurun_out/pytest_joint.log 2>&1; tail -3 gpurun_out/pytest_joint.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck_r01.txt 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_case.py > gpurun_out/sanitizer_racecheck_r01.txt 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitizer_memcheck_r01.txt gpurun_out/sanitizer_racecheck_r01.txt
echo done
