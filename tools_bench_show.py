#!/usr/bin/env python
"""one-screen digest of a bench.py JSON line: python tools_bench_show.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("N=%s value %.1f Mrays/s  %.1f ms/step  waves %.0f  launches %s" % (d.get("n_gpus"), d["value"], d["ms_per_step"], d.get("waves_per_step", 0), d.get("gpu_launches")))
for k in ("e2e", "e2e_cabi"):
    e = d.get(k)
    if e:
        print("  %-8s %.1f Mrays/s  h2d %s d2h %s  %s" % (k, e["value"], e.get("h2d_bytes_per_step"), e.get("d2h_bytes_per_step"), (e.get("path") or "")[:70]))
        if e.get("breakdown"):
            print("           breakdown", {a: round(b, 3) for a, b in e["breakdown"].items()})
        rg = e.get("rgb_pipeline")
        if rg:
            print("           rgb pipeline only", ("%.1f Mrays/s" % rg["value"]) if "value" in rg else rg, {a: round(b, 3) for a, b in (rg.get("breakdown") or {}).items()})
        st = e.get("stock_full_frame_sampler")
        if st:
            print("           stock sampler %.1f Mrays/s" % st["value"], {a: round(b, 3) for a, b in st["breakdown"].items()})
if r:
    print("  roofline %s: %.0f GB/s = %.3f of %.0f; kernel %.3f ms x %.0f = share %.3f; whole step frac %.3f" % (
        r.get("kernel"), r["achieved"], r["frac"], r["peak"], r["kernel_ms"], r["launches_per_step"], r["kernel_share_of_step"], r["whole_step"]["frac"]))
    ph = r.get("phases") or {}
    print("  phases", {k: round(v, 3) for k, v in (ph.get("share_of_wave_kernels") or {}).items()}, "largest", ph.get("largest"), "kernels/step %.3f" % (ph.get("kernels_share_of_step") or 0))
for key, c in (d.get("configs") or {}).items():
    if "failed" in c:
        print("  %s FAILED %s" % (key, c["failed"][:300]))
        continue
    if key == "C5":
        for s in c["sweeps"]:
            print("  C5 %-15s %.0e rays: %.1f Mrays/s  frac %.3f  hit %.3f" % (s["order"], s["rays"], s["Mrays_per_s"], s["roofline"]["frac"], s["hit_fraction"]))
    else:
        rf = c.get("roofline") or {}
        print("  %s %.1f Mrays/s  %.2f frames/s  %.0f ms/step  trace frac %s  share %s  setup %s" % (
            key, c["Mrays_per_s"], c["frames_per_s"], c["ms_per_step"], ("%.3f" % rf["frac"]) if rf.get("frac") else None,
            ("%.3f" % rf["kernel_share_of_step"]) if rf else None, c.get("setup_s")))
        if rf.get("phases"):
            print("     phases", {k: round(v, 3) for k, v in rf["phases"]["share_of_wave_kernels"].items()})
    if c.get("cpu_reference"):
        print("     cpu_reference", json.dumps(c["cpu_reference"])[:200])
cb = d.get("cpu_baseline")
if cb:
    print("  cpu_baseline", json.dumps({k: v for k, v in cb.items() if k != "configs"})[:400])
print("  clocks", d.get("clocks"))
