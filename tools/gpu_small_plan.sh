#!/bin/bash
# small plans (what one rank of an 8-GPU run sees): tail of the launch vs the split threshold (percent of pixels per block slot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export CELESTE_STRIPE_CACHE=/tmp/celeste_stripe
B="--steps 50 --warmup 3 --no-cpu-baseline --no-maximize --no-render --no-hessian"
run() { name=$1; f=$2; shift; shift; env "$@" timeout 600 python bench.py $B --fields $f > gpurun_out/sp_$name.json 2> gpurun_out/sp_$name.err; echo "$name rc=$?"; }
for f in 1 2 4; do
  for pct in 50 100 200; do run f${f}_pct$pct $f CELESTE_MARCH_SPLIT_PCT=$pct; done
done
run f10_pct100 10 X=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sp_f*_pct*.json")):
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f"{f[14:-5]:13s} {d['value']/1e6:.3f} M src/s  step {d['ms_per_step']:.3f} ms  kernel {r['kernel_ms_per_step']:.3f} ms  e2e {d['e2e']['value']/1e6:.3f} M  sources {d['config']['sources']}")
    except Exception as e:
        print(f, "failed", e)
PY
echo done
