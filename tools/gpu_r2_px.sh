#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 20 --warmup 3 --no-cpu-baseline --no-render --no-single --no-maximize"
for px in 0 300 450 600 900 1400; do
  for cfg in "10 1000" "1 1250"; do
    set -- $cfg
    CELESTE_UNIT_PIXELS=$px timeout 600 python bench.py $B --fields $1 --sources-per-field $2 > gpurun_out/px.json 2> gpurun_out/px.err
    python - "$px" "$1x$2" <<'PY'
import json, sys
d = json.load(open("gpurun_out/px.json"))
out = []
for nm, leg in (("grad", d), ("hess", d["hessian"])):
    r = leg["roofline"]
    out.append(f"{nm} {leg['ms_per_step']:.3f} " + " ".join(f"{v['ms_per_step']:.3f}" for v in r["kernels"].values()))
print("px", sys.argv[1], sys.argv[2], " | ".join(out))
PY
  done
done
