#!/bin/bash
# A/B of unit_walk_kernel load variants (cp.async record ring, prefetch distance) on one box, same cached stripe
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-single"
V=$PWD/celeste.jl_b200/variants
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$?"; }
for v in "$@"; do run $v CELESTE_CUDA_LIB=$V/libceleste_cuda_$v.so; done
python - "$@" <<'PY'
import json, sys
for n in sys.argv[1:]:
    try:
        d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
        r = d["roofline"]; h = d.get("hessian", {})
        hk = {k: round(v["ms_per_step"], 3) for k, v in h.get("roofline", {}).get("kernels", {}).items()}
        gk = {k: round(v["ms_per_step"], 3) for k, v in r.get("kernels", {}).items()}
        print(f"{n:9s} grad {d['value']/1e6:.3f} M step {d['ms_per_step']:.3f} ms {gk} | hess {h.get('value', 0)/1e6:.3f} M {h.get('ms_per_step')} ms {hk} parity {d['parity_check'].get('max_rel_d')} {h.get('parity_check', {}).get('max_rel_h')}")
    except Exception as e:
        print(n, "failed", e)
PY
