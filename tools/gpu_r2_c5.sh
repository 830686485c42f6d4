#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python tools/config5_diag.py > gpurun_out/config5_diag.txt 2>&1; tail -40 gpurun_out/config5_diag.txt
