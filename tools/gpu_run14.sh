#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pixel_kernel|setup_kernel|epilogue_kernel" -c 30 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-maximize > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_grad_final python tools/profile_step.py 10 1 3 > gpurun_out/ncu_grad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -o gpurun_out/prof_hess_final python tools/profile_step.py 10 2 3 > gpurun_out/ncu_hess.log 2>&1
echo done
