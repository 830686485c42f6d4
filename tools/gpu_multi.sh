#!/bin/bash
# bench.py on N GPUs of one box, launched exactly as the driver does (torchrun, one rank per GPU):  gpurun --gpus N -- bash tools/gpu_multi.sh N
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "$N gpu rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python tools/show_bench.py gpurun_out/bench_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print('maximize', d.get('maximize')); print('scaling', d.get('scaling'), 'n_gpus', d.get('n_gpus'))"
