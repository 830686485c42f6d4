import json, sys
def f(x, fmt):
    return "None" if x is None else format(x, fmt)
for fn in sys.argv[1:]:
    try:
        d = json.load(open(fn))
    except Exception as e:
        print(fn, "ERR", e, open(fn.replace(".json", ".err")).read()[-600:] if fn.endswith(".json") else "")
        continue
    r = d["roofline"]
    print(f"{fn}: grad {d['value']:.0f} src/s ({d['ms_per_step']:.3f} ms; runs {['%.3f' % x for x in d['stability']['ms_per_step_runs']]}) "
          f"e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.2f} ms) {r['kernel']} {r['kernel_ms_per_step']:.3f} ms "
          f"frac {f(r['frac'], '.3f')} executed {f(r['frac_executed'], '.3f')} pipe {f(r['fp64_pipe_active_pct'], '.1f')}")
    print("   kernels", {k: round(v["ms_per_step"], 3) for k, v in r["kernels"].items()}, "parity", d["parity_check"])
    h = d.get("hessian")
    if h:
        r = h["roofline"]
        print(f"   hess {h['value']:.0f} src/s ({h['ms_per_step']:.3f} ms) e2e packed {h['e2e']['value']:.0f} ({h['e2e']['ms_per_step']:.2f} ms) "
              f"dense {h['e2e_dense']['value']:.0f} ({h['e2e_dense']['ms_per_step']:.2f} ms) {r['kernel']} {r['kernel_ms_per_step']:.3f} ms "
              f"frac {f(r['frac'], '.3f')} contract_ratio {f(r.get('contract_ratio'), '.2f')} pipe {f(r['fp64_pipe_active_pct'], '.1f')}")
        print("   kernels", {k: round(v["ms_per_step"], 3) for k, v in r["kernels"].items()}, "parity", h["parity_check"])
    if "cpu_baseline" in d:
        print("   cpu grad", round(d["cpu_baseline"]["value"]), "src/s", d["cpu_baseline"]["cores"], "cores;",
              "hess", round(h["cpu_baseline"]["value"]) if h and "cpu_baseline" in h else None)
    for k in ("maximize", "single_call", "render"):
        if d.get(k):
            print("  ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in d[k].items() if a != "what"})
    print("   clocks", d.get("clocks"))
