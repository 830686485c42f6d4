#!/bin/bash
# unit_kernel first light: parity subset, then bench A/B (unit vs pixel for the Hessian; unit vs march for the gradient)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unit_kernel or full_size_field or two_field or plan_device or deterministic or task_mask" > gpurun_out/pytest_unit.log 2>&1; tail -4 gpurun_out/pytest_unit.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-maximize --no-render"
timeout 600 python bench.py $B > gpurun_out/u_default.json 2> gpurun_out/u_default.err; echo "default rc=$?"
CELESTE_GRAD_KERNEL=unit timeout 600 python bench.py $B > gpurun_out/u_gradunit.json 2> gpurun_out/u_gradunit.err; echo "gradunit rc=$?"
python tools/show_bench.py gpurun_out/u_default.json gpurun_out/u_gradunit.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:unit_ -s 6 -c 3 -f -o gpurun_out/prof_unit python tools/profile_step.py 10 2 3 > gpurun_out/ncu_unit.log 2>&1; echo "ncu unit rc=$?"
echo done
