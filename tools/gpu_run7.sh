#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:pixel_kernel|setup_kernel|epilogue_kernel|prep_image_kernel|dfma_peak_kernel" -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "2gpu ref rc=$?"
tail -3 gpurun_out/bench_2gpu.err
echo done
