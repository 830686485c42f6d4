#!/bin/bash
# maximize loop: tests + bench leg; A/B of newton_step_kernel occupancy variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "maximizer or infer or joint or kl" > gpurun_out/pytest_max.log 2>&1; tail -4 gpurun_out/pytest_max.log
B="--steps 5 --warmup 3 --no-cpu-baseline --no-render --no-single"
timeout 900 python bench.py $B > gpurun_out/bench_max.json 2> gpurun_out/bench_max.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_max.err
python tools/show_bench.py gpurun_out/bench_max.json | grep -E "maximize|hess "
for v in nminb3 nminb4 jacobi; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/variants/libceleste_cuda_$v.so timeout 600 python bench.py $B > gpurun_out/bench_max_$v.json 2> gpurun_out/bench_max_$v.err
  echo $v; python tools/show_bench.py gpurun_out/bench_max_$v.json | grep -E "maximize"
done
timeout 600 python tools/maximize_profile.py 10000 > gpurun_out/maximize_profile.txt 2>&1; tail -6 gpurun_out/maximize_profile.txt
echo done
