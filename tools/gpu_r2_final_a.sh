#!/bin/bash
# round 2, evidence pass A (same commit as the final bench): ncu --set full of every kernel of both modes on the bench
# workload, the ncu launch list of a bench run, compute-sanitizer memcheck + racecheck
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unit_|epilogue|slotbr" -s 5 -c 14 -f -o gpurun_out/prof_units python tools/profile_step.py 10 1,2 2 > gpurun_out/ncu_units.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/prof_units.ncu-rep 2>&1 | grep -E "==|time_duration" | head -40
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unit_|epilogue|slotbr|newton|prep_image|render|setup" -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-render --no-single > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -2 gpurun_out/sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -2 gpurun_out/sanitizer_racecheck.txt
echo done
