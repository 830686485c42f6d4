#!/bin/bash
# march_kernel tuning sweep: segment cap, block size, L1 carve-out
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
V=$PWD/celeste.jl_b200/variants
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"; }
run base CELESTE_CUDA_LIB=$V/libceleste_cuda_base.so
run seg32 CELESTE_CUDA_LIB=$V/libceleste_cuda_seg32.so
run seg51 CELESTE_CUDA_LIB=$V/libceleste_cuda_seg51.so
run t192b2 CELESTE_CUDA_LIB=$V/libceleste_cuda_t192b2.so
run t96b4 CELESTE_CUDA_LIB=$V/libceleste_cuda_t96b4.so
run t256b1 CELESTE_CUDA_LIB=$V/libceleste_cuda_t256b1.so
run carve25 CELESTE_CUDA_LIB=$V/libceleste_cuda_base.so CELESTE_MARCH_CARVEOUT=25
run carve50 CELESTE_CUDA_LIB=$V/libceleste_cuda_base.so CELESTE_MARCH_CARVEOUT=50
run carve75 CELESTE_CUDA_LIB=$V/libceleste_cuda_base.so CELESTE_MARCH_CARVEOUT=75
run carve100 CELESTE_CUDA_LIB=$V/libceleste_cuda_base.so CELESTE_MARCH_CARVEOUT=100
python - <<'PY'
import json
for n in ("base", "seg32", "seg51", "t192b2", "t96b4", "t256b1", "carve25", "carve50", "carve75", "carve100"):
    try:
        d = json.load(open(f"gpurun_out/ab_{n}.json"))
        r = d["roofline"]
        print(f"{n:9s} {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel']} {r['kernel_ms_per_step']:.3f} ms  frac {r['frac']:.3f}  e2e {d['e2e']['value']/1e6:.3f} M  sm {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(n, "failed", e)
PY
echo done
