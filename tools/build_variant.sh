#!/bin/bash
# Build a tuning variant of the library next to the shipped one (never loaded unless CELESTE_CUDA_LIB points at it):
#   tools/build_variant.sh seg32 -DCELESTE_MARCH_MAXSEG=32
#   tools/build_variant.sh t64b6 -DCELESTE_MARCH_THREADS=64 -DCELESTE_MARCH_MINB=6
# tools/gpu_tune_sweep.sh times such variants with bench.py (A/B on one box, same cached stripe).
set -e
name=$1; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$root/celeste.jl_b200/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" -Xptxas -v \
  -o "$root/celeste.jl_b200/variants/libceleste_cuda_$name.so" "$root/celeste.jl_b200/csrc/celeste_abi.cu" 2>&1 | grep -A2 "unit_walk_kernelILi1" | tail -2
