#!/bin/bash
# last evidence pass of round 2: the fast-reciprocal build of newton_step_kernel (variant A: its Newton tests + the bench
# line) and the shipped build (B: whole GPU suite, smoke, reference arm, bench line) on ONE box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
V=$PWD/celeste.jl_b200/variants/libceleste_cuda_fastrcp.so
if [ -f "$V" ]; then
CELESTE_CUDA_LIB=$V timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_maximize.py -m gpu -x -q -k "maxim or newton or config5 or infer or tr_" > gpurun_out/pytest_fastrcp.log 2>&1; tail -2 gpurun_out/pytest_fastrcp.log
CELESTE_CUDA_LIB=$V timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_1gpu_fastrcp.json 2> gpurun_out/bench_1gpu_fastrcp.err; echo "bench A rc=$?"
python tools/show_bench.py gpurun_out/bench_1gpu_fastrcp.json | grep -E "grad|maximize" | cut -c1-330
fi
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench B rc=$?"
python tools/show_bench.py gpurun_out/bench_1gpu.json | grep -E "grad|hess|maximize|single" | cut -c1-330
echo done
