#!/bin/bash
# A/B at two plan sizes: default vs $VARIANTS
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
for v in default $VARIANTS; do
  lib=$PWD/celeste.jl_b200/variants/libceleste_cuda_$v.so; [ "$v" = default ] && lib=$PWD/celeste.jl_b200/libceleste_cuda.so
  CELESTE_CUDA_LIB=$lib timeout 600 python bench.py $B > gpurun_out/ab2.json 2> gpurun_out/ab2.err; show "$v 10x1000" gpurun_out/ab2.json
  CELESTE_CUDA_LIB=$lib timeout 600 python bench.py $B --fields 1 --sources-per-field 1250 > gpurun_out/ab2.json 2> gpurun_out/ab2.err; show "$v 1x1250" gpurun_out/ab2.json
done
echo done
