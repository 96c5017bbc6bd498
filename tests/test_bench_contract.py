"""The bench line's contract, checked on the last full line recorded on a B200 (profiles/r2r_bench_full_1gpu_final.json) and on
the recorded 8-GPU line: every key the driver reads is there, with the right type and a sane value."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_recorded_bench_line_carries_the_contract():
    d = _line("r2r_bench_full_1gpu_final.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["metric"] == d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 1000
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and 0 < e["value"] < d["value"] * 1.05
    assert e["d2h_bytes_per_step"] >= 1024 * 1024 * 64 * 16
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    shares = r["phases"]["share_of_wave_kernels"]
    assert abs(sum(shares.values()) - 1.0) < 1e-6 and r["phases"]["largest"] == max(shares, key=shares.get)
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "reference" and c["cores"] >= 1
    k = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(k) and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the other configurations of the metric ride in `configs`
    assert {"C3", "C4", "C5"} <= set(d["configs"])
    for s in d["configs"]["C5"]["sweeps"]:
        assert 0 < s["roofline"]["frac"] < 1 and len(s["ms_all"]) in (1, 7) and s["ms"] == min(s["ms_all"])
    best = max(s["roofline"]["frac"] for s in d["configs"]["C5"]["sweeps"] if s["order"] == "random")
    assert best > 0.55        # independent random rays over 10,000 spheres: the north-star fraction is within reach


def test_recorded_scaling_lines():
    v = {n: _line("r2q_scaling_same_box_%dgpu.json" % n) for n in (1, 2, 4, 8)}
    for n, d in v.items():
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["workload"] == v[1]["config"]["workload"]
    assert v[8]["value"] / v[1]["value"] > 7.0 and v[8]["roofline"]["collective"]["share_of_step"] < 0.02
