#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 2 -c 1 -f -o gpurun_out/prof_task python tools/profile_step.py 10 1 3 > gpurun_out/ncu_task.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -f -o gpurun_out/prof_hess python tools/profile_step.py 10 2 3 > gpurun_out/ncu_hess.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:task_kernel|pixel_kernel|setup_kernel|epilogue_kernel|prep_image|pair_kernel" -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-maximize > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 python tools/maximize_profile.py 10000 > gpurun_out/maxprof.log 2>&1; tail -3 gpurun_out/maxprof.log
echo done
