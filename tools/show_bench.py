import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-400:] if f.endswith(".json") else "")
        continue
    r = d["roofline"]
    print(f"{f}: grad {d['value']:.0f} src/s ({d['ms_per_step']:.3f} ms) e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.2f} ms) "
          f"pix {r['kernel_ms_per_step']:.3f} ms frac {r['frac']:.3f} ach {r['achieved']:.2f} TF peak {r['peak']:.1f}")
    h = d.get("hessian")
    if h:
        r = h["roofline"]
        print(f"   hess {h['value']:.0f} src/s ({h['ms_per_step']:.3f} ms) e2e {h['e2e']['value']:.0f} ({h['e2e']['ms_per_step']:.2f} ms) "
              f"pix {r['kernel_ms_per_step']:.3f} ms frac {r['frac']:.3f} ach {r['achieved']:.2f} TF")
    if "cpu_baseline" in d:
        print("   cpu grad", round(d["cpu_baseline"]["value"]), "src/s", d["cpu_baseline"]["cores"], "cores;",
              "hess", round(h["cpu_baseline"]["value"]) if h and "cpu_baseline" in h else None)
    print("   clocks", d.get("clocks"))
