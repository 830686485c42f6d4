"""The committed evidence under profiles/ is what bench.py and the docs read: keep it loadable and self-consistent
(CPU only; nothing here touches a GPU or the reference tree)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_ncu_traffic_has_the_kernels_the_roofline_looks_up():
    """bench.py's roofline() reads executed flop / DRAM bytes / pipe activity per kernel from ncu_traffic.json, keyed by
    kernel name (unit_bg_kernel by prefix) and valid for the 10 000-source stripe only."""
    t = json.load(open(os.path.join(PROF, "ncu_traffic.json")))
    for name in ("unit_walk_kernel<1>", "unit_walk_kernel<2>", "unit_moment_kernel"):
        assert name in t, name
    assert any(k.startswith("unit_bg_kernel<") for k in t)
    for name, v in t.items():
        if not name.startswith(("unit_", "epilogue", "slotbr")):
            continue
        assert v["sources"] == 10000, name
        assert v["executed_flop_per_launch"] > 0 and v["dram_bytes_per_launch"] > 0, name
        assert 0.0 < v["fp64_pipe_active_pct"] <= 100.0, name


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_keep_the_contract(n):
    """Every committed round-2 bench line carries the keys of the bench contract, a green parity_check in both modes,
    clean clocks, and a roofline whose kernel time fits inside the step."""
    d = _line(os.path.join(PROF, f"bench_r02_{n}gpu.json"))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["dtype"] == "f64" and d["scaling"] == "strong"
    assert d["value"] == pytest.approx(10000 / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert d["e2e"]["value"] < d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "fp64" and 0.0 < r["frac"] <= 1.0
    assert r["kernel_ms_per_step"] <= d["ms_per_step"] and r["achieved"] <= r["peak"]
    for leg in (d, d["hessian"]):
        pc = leg["parity_check"]
        assert pc["ok"] and pc["counters_equal"] and pc["n"] >= 32 * n
        assert pc["max_rel_v"] <= 1e-8 and pc["max_rel_d"] <= 1e-8
    assert d["hessian"]["parity_check"]["max_rel_h"] <= 1e-8
    h = d["hessian"]["roofline"]
    assert 0.0 < h["frac"] <= 1.0 and "contract_ratio" in h          # the contract count is NOT reported as a fraction
    if n == 1:
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
        assert d["maximize"]["catalog_check"]["sources_checked"] > 100


def test_reference_arm_line():
    d = _line(os.path.join(PROF, "bench_r02_reference_arm.json"))
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["metric"] == _line(os.path.join(PROF, "bench_r02_1gpu.json"))["metric"]


def test_sanitizer_logs_are_clean():
    for f in glob.glob(os.path.join(PROF, "sanitizer_*_r02.txt")):
        txt = open(f).read()
        assert "ERROR SUMMARY: 0 errors" in txt or "0 hazards displayed (0 errors, 0 warnings)" in txt, f
