#!/bin/bash
# full verification of the march_kernel build: whole GPU suite, smoke, full bench (both arms), launch list, ncu captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_1gpu.err
python tools/show_bench.py gpurun_out/bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:march|task_kernel|pixel_kernel|setup_kernel|epilogue_kernel|prep_image|pair_kernel" -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-maximize --no-render > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/prof_march python tools/profile_step.py 10 1 3 > gpurun_out/ncu_march.log 2>&1; echo "ncu march rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pixel_kernel -s 2 -c 1 -f -o gpurun_out/prof_hess python tools/profile_step.py 10 2 3 > gpurun_out/ncu_hess.log 2>&1; echo "ncu hess rc=$?"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck.txt 2>&1; tail -2 gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/sanitizer_racecheck.txt 2>&1; tail -2 gpurun_out/sanitizer_racecheck.txt
echo done
