#!/bin/bash
# A/B: unit_bg_kernel beside the walk of the units without neighbours (CELESTE_BG_OVERLAP = blocks per SM; 0 = in sequence)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 30 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-single"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $EXTRA > gpurun_out/ov_$name.json 2> gpurun_out/ov_$name.err; echo "$name rc=$?"; }
EXTRA=""
for o in 0 1 2; do run big$o CELESTE_BG_OVERLAP=$o; done
EXTRA="--fields 1 --sources-per-field 1250"
for o in 0 1 2 3; do run small$o CELESTE_BG_OVERLAP=$o; done
python - <<'PY'
import json
for n in ["big0", "big1", "big2", "small0", "small1", "small2", "small3"]:
    try:
        d = json.loads(open(f"gpurun_out/ov_{n}.json").read().strip().splitlines()[-1])
        r = d["roofline"]; h = d.get("hessian", {})
        hk = {k: round(v["ms_per_step"], 3) for k, v in h.get("roofline", {}).get("kernels", {}).items()}
        gk = {k: round(v["ms_per_step"], 3) for k, v in r.get("kernels", {}).items()}
        print(f"{n:7s} ov {r.get('bg_overlap_blocks_per_sm')} grad {d['value']/1e6:.3f} M step {d['ms_per_step']:.4f} ms e2e {d['e2e']['value']/1e6:.3f} {gk} | hess {h.get('value', 0)/1e6:.3f} M {h.get('ms_per_step'):.4f} ms {hk} parity {d['parity_check'].get('max_rel_d'):.2e} {h.get('parity_check', {}).get('max_rel_h'):.2e} cnt {d['parity_check'].get('counters_equal')}")
    except Exception as e:
        print(n, "failed", e, open(f"gpurun_out/ov_{n}.err").read()[-600:])
PY
# correctness of the two-stream sequence incl. the captured graph of small plans
CELESTE_BG_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
