#!/bin/bash
# A/B: default vs variants listed in $VARIANTS (libs under celeste.jl_b200/variants), gradient + Hessian legs only
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 20 --warmup 3 --no-cpu-baseline --no-render --no-maximize --no-single"
timeout 600 python bench.py $B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
echo default; python tools/show_bench.py gpurun_out/ab_default.json | grep -E "grad|hess |kernels"
for v in $VARIANTS; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/variants/libceleste_cuda_$v.so timeout 600 python bench.py $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo $v; python tools/show_bench.py gpurun_out/ab_$v.json | grep -E "grad|hess |kernels"
done
echo done
