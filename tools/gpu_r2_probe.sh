#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python tools/parity_probe.py 2>&1 | tail -2
CELESTE_EPILOGUE=block python tools/parity_probe.py 2>&1 | tail -2
CELESTE_HESS_KERNEL=pixel python tools/parity_probe.py 2>&1 | tail -2
CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/variants/libceleste_cuda_rb1.so python tools/parity_probe.py 2>&1 | tail -2
