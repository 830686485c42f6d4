#!/bin/bash
# A/B: default vs variants listed in $VARIANTS (libs under celeste.jl_b200/variants), gradient + Hessian legs only
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 20 --warmup 3 --no-cpu-baseline --no-render --no-maximize --no-single"
show() {
python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
out = []
for nm, leg in (("grad", d), ("hess", d["hessian"])):
    r = leg["roofline"]
    out.append(f"{nm} {leg['ms_per_step']:.3f} " + " ".join(f"{v['ms_per_step']:.3f}" for v in r["kernels"].values()))
print(sys.argv[1], " | ".join(out))
PY
}
timeout 600 python bench.py $B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; show default gpurun_out/ab_default.json
for v in $VARIANTS; do
  CELESTE_CUDA_LIB=$PWD/celeste.jl_b200/variants/libceleste_cuda_$v.so timeout 600 python bench.py $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; show $v gpurun_out/ab_$v.json
done
echo done
