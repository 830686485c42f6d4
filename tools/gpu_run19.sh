#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parity or golden" > gpurun_out/pytest_gpu_quick.log 2>&1; tail -2 gpurun_out/pytest_gpu_quick.log
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_cpasync6.json 2> gpurun_out/bench_cpasync6.err
CELESTE_CUDA_LIB=build_variants/libceleste_minb5.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_cpasync5.json 2> gpurun_out/bench_cpasync5.err
CELESTE_CUDA_LIB=build_variants/libceleste_minb4.so timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-hessian > gpurun_out/bench_cpasync4.json 2> gpurun_out/bench_cpasync4.err
for f in 6 5 4; do python tools/show_bench.py gpurun_out/bench_cpasync$f.json; done
echo done
