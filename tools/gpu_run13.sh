#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/$2 CELESTE_CHUNK_PIXELS=$3 timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-maximize $4 > gpurun_out/bench_v10_$1.json 2> gpurun_out/bench_v10_$1.err; }
run smem6 libceleste_cuda_smem6.so 512 ""
run reg6 libceleste_cuda_reg6.so 512 --no-hessian
run reg5 libceleste_cuda_reg5.so 512 --no-hessian
run reg4 libceleste_cuda_reg4.so 512 --no-hessian
run t32r20_c256 libceleste_cuda_t32r20.so 256 ""
run t32r20_c128 libceleste_cuda_t32r20.so 128 --no-hessian
run t64r10_c256 libceleste_cuda_t64r10.so 256 ""
echo done
