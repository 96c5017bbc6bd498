#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 hot path (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rng mt|philox]

Workload (BASELINE.json configs[1]): Cornell box (demos/cornell_box.py scene), PinholeCamera 1024x1024,
256 samples/pixel, 64 spectral bins, one spectral ray.  One STEP = one full frame (one observe() pass).
Metric: Mrays/s, "ray" = the reference's ray counter (primary rays + daughters spawned,
raysect/optical/ray.pyx:375-378,537-547).  frames/s = 1000 / ms_per_step.

  value     device-resident: scene, tables and frame buffers live in HBM; timed with CUDA events around
            K x rsb_render_dev (+ the NCCL reduce of the frame when N > 1), max over ranks.
  e2e       through the public API with HOST buffers: per step the pixel lists and spectral tables go
            host->device from pinned memory and the reduced frame (mean, variance) comes device->host into
            pinned memory, inside the timed region.
  roofline  algorithmic bytes of the render kernel (SURVEY 8(d) model, from the kernel's own traversal
            counters in an untimed counting pass) / mean kernel time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the compiled reference (oracle/_ref) on the box's host cores, bounded sample.

--impl reference times the reference's own Cython path (MulticoreEngine, all host cores) on a bounded sample
of the same workload and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

WORKLOAD = dict(name="cornell_box 1024x1024 x 256 spp x 64 bins (BASELINE configs[1])", pixels=1024, spp=256, bins=64)
CPU_SAMPLE = dict(pixels=192, spp=8, bins=64)
DEFAULT_PASSES = 8   # 256 spp = 8 accumulated observe() passes of 32 spp (the reference's progressive-render loop)
RAY_CFG = dict(extinction_prob=0.01, extinction_min_depth=3, max_depth=500, importance_sampling=True,
               important_path_weight=0.25)   # demos/cornell_box.py:147-156
MIN_WL, MAX_WL = 375.0, 740.0               # observer defaults, observer.pyx:116-117


# ----------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path (oracle/_ref = the unmodified compiled reference) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import harness
    import scenes
    if not harness.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) not present in this snapshot"}))
        return 0
    api = harness.ref_api()
    from raysect.core.workflow import MulticoreEngine
    cores = os.cpu_count() or 1

    class Counting(MulticoreEngine):
        rays = 0

        def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
            def counted(result, *a, **k):
                Counting.rays += result[2]
                update(result, *a, **k)
            super().run(tasks, render, counted, render_args, render_kwargs, update_args, update_kwargs)

    s = CPU_SAMPLE
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(s["pixels"], s["pixels"]), samples=s["spp"], bins=s["bins"],
                                      path_weight=RAY_CFG["important_path_weight"])
    cam.render_engine = Counting(processes=cores)
    world.build_accelerator()
    times, rays = [], []
    # The reference renders its sample in ONE observe() call whatever --passes says: that is its faster mode
    # (MulticoreEngine forks its workers on every observe(); measured here, 4 passes of the bounded sample run at
    # half the rays/s of one pass), and the frame is statistically the same.
    for i in range(args.warmup + args.steps):
        Counting.rays = 0
        pipe.accumulate = False
        t0 = time.perf_counter()
        cam.observe()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays.append(Counting.rays)
    total_t, total_r = sum(times), sum(rays)
    value = total_r / total_t / 1e6
    sample = "cornell_box %dx%d x %d spp (one observe() pass) x %d bins, MulticoreEngine(%d)" % (
        s["pixels"], s["pixels"], s["spp"], s["bins"], cores)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, len(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_step": total_r / max(1, len(rays)),
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def algorithmic_bytes(c, bins, spp):
    """per ray 72 B (56 in + 16 out) + 16 B per kd branch + 8 B per leaf + 4 B per item id
    + 128 B per analytic primitive test + 48 B per triangle test; per path the spectral table reads
    (surface + volume interactions, bins*8 B each) and the pixel's share of the frame write (bins*20/spp)."""
    ray = trace_algorithmic_bytes(c)
    # World.contains: 24 B point in, 12 B per kd node descended (split + child), 4 B per item id, 128 B per
    # primitive containment test
    contains = 24 * c["contains"] + 12 * c["contains_nodes"] + 4 * c["contains_items"] + 128 * c["contains_prim_tests"]
    spectral = 8 * bins * c.get("table_reads", 0)
    frame = c["paths"] * bins * 20.0 / spp
    return ray + contains + spectral + frame


def trace_algorithmic_bytes(c):
    """World.hit share of the model = what k_wf_trace touches: 56 B ray in + 16 B hit out, 16 B per kd branch,
    8 B per leaf header, 4 B per item id, 128 B per analytic primitive test, 48 B per triangle test."""
    return 72 * c["rays"] + 16 * c["branches"] + 8 * c["leaves"] + 4 * c["items"] + 128 * c["prim_tests"] + 48 * c["tri_tests"]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; source_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.distributed import FrameRenderer
    from source_b200.engine import Device

    w = dict(WORKLOAD)
    if args.pixels:
        w["pixels"] = args.pixels
    if args.spp:
        w["spp"] = args.spp
    mode = cabi.RNG_MT19937_64 if args.rng == "mt" else cabi.RNG_PHILOX

    device = Device(local_rank)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(w["pixels"], w["pixels"]), samples=w["spp"], bins=w["bins"],
                                      path_weight=RAY_CFG["important_path_weight"])
    cam.rng_mode = mode
    pipe.accumulate = False
    world._device = device
    accel = world.build_accelerator()
    if w["spp"] % args.passes:
        raise SystemExit("bench.py: --passes must divide the samples per pixel")
    renderer = FrameRenderer(cam, accel, rank, world_size, tile=16, passes=args.passes)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for i in range(args.warmup):
        renderer.step_device(seed=1 + i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays_t = torch.zeros(1, dtype=torch.int64, device=dev)
    trace_ms = trace_launches = launches = waves = 0
    barrier()
    e0.record()
    for i in range(args.steps):
        rays_t += renderer.step_device(seed=100 + i, time_trace=True)
        rs = device.render_stats()
        trace_ms += rs["trace_ms"]
        trace_launches += rs["trace_launches"]
        launches += rs["launches"]
        waves += rs["waves"]
    e1.record()
    barrier()
    if rank == 0:
        clocks.stop_flag.set()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kernel_ms = trace_ms / max(1, trace_launches)
    if world_size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    total_ms, total_rays = float(ms.item()), int(rays_t.item())
    value = total_rays / total_ms / 1e3   # Mrays/s

    # ---- end to end through the public API (host buffers, pinned) ---------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    renderer.step_host(seed=7)   # warm-up (allocates pinned buffers)
    barrier()
    t_rays = 0
    e0.record()
    for i in range(e2e_steps):
        t_rays += renderer.step_host(seed=200 + i)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    r2 = torch.tensor([t_rays], dtype=torch.int64, device=dev)
    if world_size > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        dist.all_reduce(r2, op=dist.ReduceOp.SUM)
    e2e_value = int(r2.item()) / float(ms2.item()) / 1e3

    # ---- roofline: counting pass (untimed) -----------------------------------------------------------
    counters = renderer.count_pass(seed=100)
    alg_step = algorithmic_bytes(counters, w["bins"], w["spp"])
    # the dominant kernel, k_wf_trace, owns the World.hit part of the model: ray in/out + kd nodes + leaf items +
    # primitive tests of the hit queries (the contains-query and table/frame terms belong to shade/finalize)
    # rank 0's own share of the frame against rank 0's own kernel time: the roofline is a per-GPU figure
    alg = trace_algorithmic_bytes(renderer.local_counters)
    peaks = {}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    trace_ms_step = trace_ms / args.steps
    achieved = alg / (trace_ms_step * 1e-3) / 1e9
    achieved_step = alg_step / (total_ms / args.steps * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "render_traffic.json")
    if os.path.exists(tp):
        # ncu --set full of one mid-frame k_wf_trace launch: (dram__bytes_read.sum + dram__bytes_write.sum) / rays traced
        traffic = json.load(open(tp)).get("dram_bytes_per_ray")
        if traffic is not None:
            traffic = traffic * renderer.local_counters["rays"] / max(1.0, trace_launches / args.steps)

    # ---- CPU baseline (rank 0, N = 1 only): the compiled reference in a clean subprocess --------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                                 capture_output=True, text=True, timeout=900).stdout.strip().splitlines()
            ref = json.loads(out[-1])
            cpu = ref.get("cpu_baseline", {"unavailable": ref.get("unavailable")})
        except Exception as exc:   # noqa: BLE001
            cpu = {"unavailable": repr(exc)}

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"] if not (args.pixels or args.spp) else "cornell_box %dx%d x %d spp x %d bins" % (w["pixels"], w["pixels"], w["spp"], w["bins"]),
                       "passes": "%d spp accumulated as %d observe() passes of %d spp (streams keyed on (pass, pixel)), merged with "
                                 "StatsArray3D.combine_samples" % (w["spp"], args.passes, w["spp"] // args.passes)
                                 if args.passes > 1 else "one observe() pass of %d spp" % w["spp"],
                       "rng": "mt19937_64 per (pass, pixel)" if args.rng == "mt" else "philox4x32-10 per (pass, pixel, sample)",
                       "partition": "16x16 px tiles interleaved over ranks, one NCCL reduce(sum) of the frame" if world_size > 1 else "single GPU",
                       "l2": "frame buffers 1.07 GB per step exceed the 126 MB L2; scene (4.5 KB) is shared-memory resident by design"},
            "frames_per_s": 1e3 * args.steps / total_ms, "rays_per_step": total_rays / args.steps,
            "gpu_launches": launches, "waves_per_step": waves / args.steps,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": renderer.h2d_bytes, "d2h_bytes_per_step": renderer.d2h_bytes,
                    "steps": e2e_steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_wf_trace", "kernel_ms": kernel_ms,
                         "launches_per_step": trace_launches / args.steps, "kernel_ms_per_step": trace_ms_step,
                         "kernel_share_of_step": trace_ms_step / (total_ms / args.steps),
                         "algorithmic_bytes_per_launch": alg / max(1.0, trace_launches / args.steps),
                         "algorithmic_bytes_per_step": alg, "whole_step": {"algorithmic_bytes": alg_step, "achieved": achieved_step,
                                                                           "frac": achieved_step / peak},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650", "counters": counters},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rng", default="mt", choices=["mt", "philox"])
    ap.add_argument("--pixels", type=int, default=0, help="override frame size (development only)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (development only)")
    ap.add_argument("--passes", type=int, default=DEFAULT_PASSES,
                    help="render the frame's samples as this many accumulated observe() passes, concurrently")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
