#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or deterministic or multi_field or plan_device" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
run() { # name lib chunk
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/$2 CELESTE_CHUNK_PIXELS=$3 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4_$1.json 2> gpurun_out/bench_v4_$1.err
}
run t128_c512 libceleste_cuda.so 512
run t128b_c512 libceleste_cuda_t128b.so 512
run t128b_c256 libceleste_cuda_t128b.so 256
run t64a_c256 libceleste_cuda_t64a.so 256
run t64b_c256 libceleste_cuda_t64b.so 256
run t64b_c128 libceleste_cuda_t64b.so 128
run t32a_c128 libceleste_cuda_t32a.so 128
run t32a_c256 libceleste_cuda_t32a.so 256
run t32a_c512 libceleste_cuda_t32a.so 512
echo done
