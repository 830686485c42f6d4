#!/bin/bash
# ncu --set full of every unit kernel of both modes on the bench workload + launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"unit_|epilogue" -s 4 -c 11 -f -o gpurun_out/prof_units python tools/profile_step.py 10 1,2 2 > gpurun_out/ncu_units.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_units.ncu-rep 2>&1 | grep -E "==|time_duration" | head -30
