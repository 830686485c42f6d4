#!/bin/bash
# N GPUs of one box:  gpurun --gpus N -- bash tools/gpu_multi.sh N
# the two-rank value test (N >= 2), then bench.py launched exactly as the driver does (torchrun, one rank per GPU)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" -le 2 ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_rank" > gpurun_out/pytest_two_rank.log 2>&1; tail -3 gpurun_out/pytest_two_rank.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "$N gpu rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python tools/show_bench.py gpurun_out/bench_${N}gpu.json
