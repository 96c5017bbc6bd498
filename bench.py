#!/usr/bin/env python
"""bench.py -- measurement of the B200 hot path (contract: the task brief / DESIGN.md section 9).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rng mt|philox] [--configs C3,C4,C5|none]

Headline workload (BASELINE.json configs[1], "C2"): Cornell box (demos/cornell_box.py scene), PinholeCamera 1024x1024,
256 samples/pixel, 64 spectral bins, one spectral ray.  One STEP = one full frame.  Metric: Mrays/s, "ray" = the
reference's ray counter (primary rays + daughters spawned, raysect/optical/ray.pyx:375-378,537-547).

  value     device-resident: scene, tables and frame buffers live in HBM; CUDA events around K x rsb_render_passes_dev
            (+ the NCCL collective that assembles the frame when N > 1), max over ranks.
  e2e       the same frame through the reference-facing seam with HOST buffers: at N = 1 the real drop-in plugin -- a
            raysect PinholeCamera + SpectralPowerPipeline2D with camera.render_engine = CudaRenderEngine(passes=8),
            wall clock around camera.observe() (task list in, pipeline.frame numpy arrays out) -- when the compiled
            reference travelled with the snapshot; else, and at N > 1, FrameRenderer.step_host (pinned host buffers).
  roofline  algorithmic bytes of the traversal kernel (SURVEY 8(d) model, from the kernels' own counters in an
            untimed counting pass) / its device time (CUDA events around every trace phase), against
            MEASURED_PEAKS.json hbm_gbs; `phases` lists the measured share of every kernel of the wave.
  configs   the rest of BASELINE.json's metric under the same clock: C4 = Cornell box + Stanford bunny refined to
            1,000,000 triangles (1024^2, 64 spp; tiles over the N ranks), C5 = ray-batch sweep over the 10,000-sphere
            field (1e7 and 1e9 device-generated rays, incoherent and Morton order; ray ranges over the N ranks),
            C3 = dispersive CSG prism 512^2 x 512 spectral bins / rays (N = 1) -- each with Mrays/s, frames/s and a
            live roofline.
  cpu_baseline  the compiled reference (oracle/_ref) on the box's host cores, bounded samples.

--impl reference times the reference's own Cython path (MulticoreEngine, all host cores) on bounded samples of the
same workloads and prints the same JSON line with "impl": "reference".
"""
import argparse
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

WORKLOAD = dict(name="cornell_box 1024x1024 x 256 spp x 64 bins (BASELINE configs[1])", pixels=1024, spp=256, bins=64)
CPU_SAMPLE = dict(pixels=512, spp=16, bins=64)      # >= 5 s per observe(): MulticoreEngine's per-call fork is amortised
CPU_SERIAL_SAMPLE = dict(pixels=128, spp=8, bins=64)
# The frame's samples are rendered as P accumulated observe() passes (the reference's progressive-render loop,
# demos/cornell_box.py:160-174), concurrently: P x pixels independent streams.  P = 8 per GPU keeps ~8 M streams per rank
# whatever N is, so that every wave of every rank stays as wide as the single-GPU ones (at N = 8 with P = 8 a rank holds
# 1 M streams for 1.2 M slots: its 659 waves ran 42 % full in round 1).
PASSES_PER_GPU = 8


def auto_passes(spp, n_gpus):
    p = PASSES_PER_GPU * n_gpus
    while p > 1 and (spp % p or spp // p < 4):
        p //= 2
    return max(1, p)
RAY_CFG = dict(extinction_prob=0.01, extinction_min_depth=3, max_depth=500, importance_sampling=True,
               important_path_weight=0.25)   # demos/cornell_box.py:147-156
MIN_WL, MAX_WL = 375.0, 740.0               # observer defaults, observer.pyx:116-117
C4 = dict(name="cornell_box + stanford_bunny refined to 1,000,000 triangles, 1024x1024 x 64 spp x 64 bins (BASELINE configs[3])",
          pixels=1024, spp=64, bins=64, triangles=1000000)
C4_CPU = dict(pixels=256, spp=4)
C5 = dict(name="ray-batch sweep over 10,000 spheres from the reference generator after seed(7) (BASELINE configs[4])",
          spheres=10000, rays=(10**7, 10**9), seed=2024)
C5_CPU_RAYS = 200000
C3 = dict(name="dispersive CSG prism 512x512, 512 spectral bins = 512 spectral rays (BASELINE configs[2])",
          pixels=512, bins=512, rays=512, spp=2)
C3_CPU = dict(pixels=64, bins=64, rays=64, spp=2)      # 64 slices: the reference forks its workers once per slice (512 forks take ~40 s whatever the frame)


def trace_algorithmic_bytes(c):
    """World.hit share of the SURVEY 8(d) model = what the traversal kernels touch: 56 B ray in + 16 B hit out, 16 B per
    kd branch, 8 B per leaf header, 4 B per item id, 128 B per analytic primitive test, 48 B per triangle test."""
    return 72 * c["rays"] + 16 * c["branches"] + 8 * c["leaves"] + 4 * c["items"] + 128 * c["prim_tests"] + 48 * c["tri_tests"]


def algorithmic_bytes(c, bins, spp):
    """whole step: + World.contains (24 B point in, 12 B per kd node descended, 4 B per item id, 128 B per containment
    test), the spectral table reads (bins * 8 B per interaction) and the pixel's share of the frame write"""
    contains = 24 * c["contains"] + 12 * c["contains_nodes"] + 4 * c["contains_items"] + 128 * c["contains_prim_tests"]
    return trace_algorithmic_bytes(c) + contains + 8 * bins * c.get("table_reads", 0) + c["paths"] * bins * 20.0 / spp


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("hbm_gbs", 6650.0)), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------
# reference arm
# ----------------------------------------------------------------------------------------------------
def _counting_engine(cores):
    from raysect.core.workflow import MulticoreEngine, SerialEngine

    class Counting(MulticoreEngine if cores > 1 else SerialEngine):
        rays = 0

        def run(self, tasks, render, update, render_args=(), render_kwargs={}, update_args=(), update_kwargs={}):
            def counted(result, *a, **k):
                Counting.rays += result[2]
                update(result, *a, **k)
            super().run(tasks, render, counted, render_args, render_kwargs, update_args, update_kwargs)
    return Counting(processes=cores) if cores > 1 else Counting(), Counting


def _ref_observe(cam, pipe, cores, steps, warmup):
    eng, cls = _counting_engine(cores)
    cam.render_engine = eng
    times, rays = [], []
    for i in range(warmup + steps):
        cls.rays = 0
        pipe.accumulate = False
        t0 = time.perf_counter()
        cam.observe()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            rays.append(cls.rays)
    return sum(rays) / sum(times) / 1e6, sum(times) / len(times), sum(rays) / len(rays)


def reference_configs(api, harness, scenes, cores, which):
    """bounded samples of C4 / C5 / C3 on the reference's own code path"""
    import numpy as np
    out = {}
    if "C4" in which:
        try:
            path = scenes.refined_bunny_rsm(C4["triangles"])
            world = scenes.cornell_mesh_scene(api, path)       # Mesh.from_file: the reference loads mesh + tree from the .rsm
            s = C4_CPU
            cam, pipe = scenes.cornell_camera(api, world, pixels=(s["pixels"], s["pixels"]), samples=s["spp"], bins=C4["bins"],
                                              path_weight=RAY_CFG["important_path_weight"])
            world.build_accelerator()
            v, sec, rays = _ref_observe(cam, pipe, cores, 1, 0)
            out["C4"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                         "sample": "%dx%d x %d spp (one observe()), MulticoreEngine(%d)" % (s["pixels"], s["pixels"], s["spp"], cores)}
        except Exception as exc:   # noqa: BLE001
            out["C4"] = {"unavailable": repr(exc)}
    if "C5" in which:
        from raysect.core import Point3D, Vector3D
        from raysect.core.math.random import seed, uniform
        from raysect.core.ray import Ray as CoreRay
        seed(7)
        world = scenes.sweep_spheres(api, uniform, C5["spheres"])
        world.build_accelerator()
        o, d, _ = scenes.sweep_rays(C5["seed"], 0, C5_CPU_RAYS, scenes.SWEEP_ORIGIN, scenes.SWEEP_TARGET, scenes.SWEEP_HALF, 0)
        rays = [CoreRay(Point3D(*o[i]), Vector3D(*d[i])) for i in range(len(o))]
        t0 = time.perf_counter()
        hits = sum(1 for r in rays if world.hit(r) is not None)
        dt = time.perf_counter() - t0
        out["C5"] = {"value": len(rays) / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "reference", "hit_fraction": hits / len(rays),
                     "sample": "%d rays of the sweep (incoherent order) through the Python world.hit loop of "
                               "demos/core/ray_intersection_hitpoints.py, 1 core" % len(rays)}
    if "C3" in which:
        s = C3_CPU
        world = scenes.prism_scene(api)
        cam, pipe = scenes.cornell_camera(api, world, pixels=(s["pixels"], s["pixels"]), samples=s["spp"], bins=s["bins"],
                                          spectral_rays=s["rays"], path_weight=0.75)
        cam.transform = api.translate(0.3, 0.2, -2.2) * api.rotate(5, -3, 0)
        world.build_accelerator()
        v, sec, rays = _ref_observe(cam, pipe, cores, 1, 0)
        out["C3"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                     "sample": "%dx%d x %d spp x %d spectral rays (one observe() = %d slices, a worker fork per slice), "
                               "MulticoreEngine(%d)" % (s["pixels"], s["pixels"], s["spp"], s["rays"], s["rays"], cores)}
    return out


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
    cores = os.cpu_count() or 1
    s = CPU_SAMPLE
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(s["pixels"], s["pixels"]), samples=s["spp"], bins=s["bins"],
                                      path_weight=RAY_CFG["important_path_weight"])
    world.build_accelerator()
    # The reference renders its sample in ONE observe() call whatever --passes says: that is its faster mode
    # (MulticoreEngine forks its workers on every observe()), and the frame is statistically the same.
    value, sec, rays = _ref_observe(cam, pipe, cores, args.steps, args.warmup)
    sample = "cornell_box %dx%d x %d spp (one observe() pass) x %d bins, MulticoreEngine(%d)" % (
        s["pixels"], s["pixels"], s["spp"], s["bins"], cores)
    # one core: SerialEngine on a smaller sample of the same workload
    q = CPU_SERIAL_SAMPLE
    cam1, pipe1 = scenes.cornell_camera(api, world, pixels=(q["pixels"], q["pixels"]), samples=q["spp"], bins=q["bins"],
                                        path_weight=RAY_CFG["important_path_weight"])
    v1, _, _ = _ref_observe(cam1, pipe1, 1, 1, 0)
    which = [] if args.configs == "none" else args.configs.split(",")
    cfgs = reference_configs(api, harness, scenes, cores, which)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample,
                         "serial_1core": {"value": v1, "unit": "Mrays/s", "cores": 1,
                                          "sample": "cornell_box %dx%d x %d spp, SerialEngine" % (q["pixels"], q["pixels"], q["spp"])},
                         "configs": cfgs},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_step": rays,
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


class Ctx:
    """what every workload needs: ranks, device, barrier, max-over-ranks timing"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; source_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world_size > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from source_b200.engine import Device
        self.device = Device(self.local_rank)
        self.peak, self.peak_source = peak_hbm()

    def barrier(self):
        if self.world_size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, value, op="max", dtype=None):
        torch = self.torch
        t = torch.tensor([value], dtype=dtype or torch.float64, device=self.dev)
        if self.world_size > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.item()


def phase_shares(stats_sum, total_ms):
    ph = {k: stats_sum[k + "_ms"] for k in ("trace", "shade", "finalize", "regen")}
    tot = sum(ph.values())
    return {"ms_per_step": ph, "share_of_wave_kernels": {k: (v / tot if tot else None) for k, v in ph.items()},
            "largest": max(ph, key=ph.get) if tot else None, "kernels_share_of_step": tot / total_ms if total_ms else None}


def timed_frames(cx, renderer, steps, warmup, seed0):
    """K frames device-resident: (total_ms max over ranks, rays summed over ranks, per-step kernel times of rank 0).
    The K timed frames run without any instrumentation; the per-kernel device times come from ONE extra frame (same
    seed as the first timed one) whose waves are bracketed with CUDA events on the launch stream."""
    torch = cx.torch
    for i in range(warmup):
        renderer.step_device(seed=seed0 + i)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays_t = torch.zeros(1, dtype=torch.int64, device=cx.dev)
    launches = waves = 0
    cx.barrier()
    e0.record()
    for i in range(steps):
        rays_t += renderer.step_device(seed=seed0 + 100 + i)
        rs = cx.device.render_stats()
        launches += rs["launches"]
        waves += rs["waves"]
    e1.record()
    cx.barrier()
    total_ms = cx.reduce(e0.elapsed_time(e1), "max")
    if cx.world_size > 1:
        cx.dist.all_reduce(rays_t, op=cx.dist.ReduceOp.SUM)
    renderer.step_device(seed=seed0 + 100, time_trace=True)
    cx.barrier()
    rs = cx.device.render_stats()
    per_step = {k: rs[k] for k in ("trace_ms", "shade_ms", "finalize_ms", "regen_ms", "trace_launches")}
    per_step["launches"], per_step["waves"] = launches / steps, waves / steps
    # the collective alone (N > 1): one more assembly of the frame that is already there
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    e2.record()
    renderer._assemble()
    e3.record()
    cx.barrier()
    per_step["assemble_ms"] = cx.reduce(e2.elapsed_time(e3), "max")
    return total_ms, int(rays_t.item()), per_step


def captured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of ``kernel`` from the committed `ncu --set full` capture of
    this bench command (profiles/r2n_ncu_full_<kernel>.txt; a mid-frame launch over the same 1.2 M-slot wave).  A figure from a
    profile, not from this run: the file is named next to it; None when the summary is not there."""
    try:
        path = os.path.join(ROOT, "profiles", "r2n_ncu_full_%s.txt" % kernel)
        total = 0.0
        for line in open(path):
            words = line.split()
            if len(words) >= 3 and words[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[words[2]]
                total += float(words[1]) * unit
        return (total, os.path.relpath(path, ROOT)) if total > 0 else (None, None)
    except Exception:   # noqa: BLE001
        return None, None


def frame_roofline(cx, renderer, total_ms_step, per_step, bins, spp, seed, kernel):
    counters = renderer.count_pass(seed=seed)
    alg_step = algorithmic_bytes(counters, bins, spp)
    alg = trace_algorithmic_bytes(renderer.local_counters)     # rank 0's own share against rank 0's own kernel time
    trace_ms = per_step["trace_ms"]
    achieved = alg / (trace_ms * 1e-3) / 1e9 if trace_ms else None
    launches = max(1.0, per_step["trace_launches"])
    traffic, traffic_source = captured_traffic(kernel) if cx.world_size == 1 else (None, None)
    return {"bound": "hbm", "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak if achieved else None,
            "traffic": traffic, "traffic_source": traffic_source, "kernel": kernel, "kernel_ms": trace_ms / launches, "launches_per_step": per_step["trace_launches"],
            "kernel_ms_per_step": trace_ms, "kernel_share_of_step": trace_ms / total_ms_step,
            "algorithmic_bytes_per_launch": alg / launches, "algorithmic_bytes_per_step": alg,
            # whole step, per GPU: every rank's share of the frame's algorithmic bytes against the step time
            "whole_step": {"algorithmic_bytes": alg_step, "achieved": alg_step / cx.world_size / (total_ms_step * 1e-3) / 1e9,
                           "frac": alg_step / cx.world_size / (total_ms_step * 1e-3) / 1e9 / cx.peak},
            "phases": phase_shares(per_step, total_ms_step), "peak_source": cx.peak_source, "counters": counters,
            "collective": None if cx.world_size == 1 else {
                "kind": "one NCCL gather of every rank's own tile rows to rank 0 (distributed.TileGather)",
                "ms_per_step": per_step["assemble_ms"], "share_of_step": per_step["assemble_ms"] / total_ms_step}}


# ----------------------------------------------------------------------------------------------------
def plugin_e2e(cx, w, args, steps):
    """N = 1: the real drop-in seam.  raysect objects, camera.render_engine = CudaRenderEngine(passes), wall clock around
    camera.observe(): flatten + upload + render + Pipeline.update into pipeline.frame's numpy arrays.  Measured twice: with
    camera.frame_sampler = WholeFrameSampler2D() (the headline: both of the reference's plugin seams filled in) and with
    the stock FullFrameSampler2D, whose nx*ny-tuple task list the reference builds and shuffles in Python before it calls
    the engine."""
    from oracle import harness
    if not harness.available():
        return None
    import scenes
    api_ref = harness.ref_api()
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D

    def measure(sampler, rgb_only=False):
        world = scenes.cornell_box(api_ref)
        cam, pipe = scenes.cornell_camera(api_ref, world, pixels=(w["pixels"], w["pixels"]), samples=w["spp"], bins=w["bins"],
                                          path_weight=RAY_CFG["important_path_weight"])
        pipe.accumulate = False
        if rgb_only:
            from raysect.optical.observer import RGBPipeline2D
            cam.pipelines = [RGBPipeline2D(display_progress=False)]
        if sampler is not None:
            cam.frame_sampler = sampler
        eng = CudaRenderEngine(seed=1, rng=args.rng, device=cx.device, passes=args.passes)
        cam.render_engine = eng
        cam.observe()                       # warm-up: allocations, kernel loading
        cx.torch.cuda.synchronize()
        eng.ray_count = 0
        parts = {"flatten_upload_s": 0.0, "render_s": 0.0, "update_s": 0.0}
        t0 = time.perf_counter()
        for _ in range(steps):
            cam.observe()
            for k in parts:
                parts[k] += eng.timing.get(k, 0.0) / steps
        dt = time.perf_counter() - t0
        parts["reference_host_s"] = dt / steps - sum(parts.values())    # observe() outside the engine: task list generation, pipeline set-up
        return {"value": eng.ray_count / dt / 1e6, "s_per_step": dt / steps, "breakdown": parts}

    stock = measure(None)
    whole = measure(WholeFrameSampler2D())
    try:
        rgb = dict(measure(WholeFrameSampler2D(), rgb_only=True),
                   note="same frame into an RGBPipeline2D alone (the demos' pipeline): per-sample CIE XYZ projection and statistics on the "
                        "device (rsb_render_slices_xyz, no spectral frame kept), d2h = the (nx, ny, 3) xyz_frame")
    except Exception as exc:   # noqa: BLE001
        rgb = {"failed": repr(exc)}
    frame_bytes = w["pixels"] * w["pixels"] * w["bins"] * 20
    return {"value": whole["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 8 * w["bins"] * 16,
            "d2h_bytes_per_step": frame_bytes + 8, "steps": steps, "s_per_step": whole["s_per_step"], "breakdown": whole["breakdown"],
            "path": "raysect PinholeCamera.observe(), camera.render_engine = CudaRenderEngine, camera.frame_sampler = WholeFrameSampler2D "
                    "-> rsb_render_slices + rsb_slice_update_frame -> SpectralPowerPipeline2D.frame numpy arrays; wall clock",
            "rgb_pipeline": rgb,
            "stock_full_frame_sampler": dict(stock, note="same call with the reference's own FullFrameSampler2D: its Python task list "
                                                          "(nx*ny tuples, shuffled) is built before the engine is called")}


def run_c4(cx, args):
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.distributed import FrameRenderer
    torch = cx.torch
    t0 = time.time()
    note = None
    have_obj = os.path.exists(scenes.BUNNY_OBJ) or os.path.exists(os.path.join(scenes.MESH_CACHE, "bunny_%d.rsm" % C4["triangles"]))
    if have_obj:
        if cx.rank == 0:
            scenes.refined_bunny_rsm(C4["triangles"])          # cached .rsm (mesh + tree); the other ranks read it
        cx.barrier()
        world = scenes.cornell_mesh_scene(api, scenes.refined_bunny_rsm(C4["triangles"]))
        name = C4["name"]
    else:
        note = "demos/resources/stanford_bunny.obj did not travel with this snapshot: bumpy icosphere (1,310,720 triangles) instead"
        verts, tris, normals = scenes.icosphere(8, radius=0.45, bumps=0.15)

        def extra(a, wd):
            a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=wd, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
                   material=a.Lambert(a.ConstantSF(0.7)))
        world = scenes.cornell_box(api, glass=False, extra=extra)
        name = "cornell_box + 1,310,720-triangle icosphere, 1024x1024 x 64 spp x 64 bins"
    cam, pipe = scenes.cornell_camera(api, world, pixels=(C4["pixels"], C4["pixels"]), samples=C4["spp"], bins=C4["bins"],
                                      path_weight=RAY_CFG["important_path_weight"])
    cam.rng_mode = cabi.RNG_MT19937_64 if args.rng == "mt" else cabi.RNG_PHILOX
    pipe.accumulate = False
    world._device = cx.device
    accel = world.build_accelerator()
    setup_s = time.time() - t0
    passes = auto_passes(C4["spp"], cx.world_size)
    renderer = FrameRenderer(cam, accel, cx.rank, cx.world_size, tile=16, passes=passes)
    steps = max(1, min(args.steps, 3))
    total_ms, rays, per_step = timed_frames(cx, renderer, steps, 1, 1000)
    ms_step = total_ms / steps
    roof = frame_roofline(cx, renderer, ms_step, per_step, C4["bins"], C4["spp"], 1100,
                          "trace phase = k_rq_walk (world walk) + k_rq_mesh (Mesh.hit) + k_rq_walk (resume)")
    out = {"workload": name, "Mrays_per_s": rays / total_ms / 1e3, "frames_per_s": 1e3 / ms_step, "ms_per_step": ms_step,
           "rays_per_step": rays / steps, "steps": steps, "n_gpus": cx.world_size, "scaling": "strong", "setup_s": setup_s,
           "passes": passes,
           "waves_per_step": per_step["waves"], "roofline": roof}
    if note:
        out["note"] = note
    accel.close()
    del renderer
    torch.cuda.empty_cache()
    return out


def run_c5(cx, args):
    import ctypes as C
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    torch, dev = cx.torch, cx.device
    it = iter(dev.rng_uniform(7, 4 * C5["spheres"]))           # the reference generator's stream after seed(7), on the device
    world = scenes.sweep_spheres(api, lambda: float(next(it)), C5["spheres"])
    acc = dev.build(world)
    hits = torch.zeros(1, dtype=torch.int64, device=cx.dev)
    sum_t = torch.zeros(1, dtype=torch.float64, device=cx.dev)
    xr = torch.zeros(1, dtype=torch.int64, device=cx.dev)
    o3, t3 = (C.c_double * 3)(*scenes.SWEEP_ORIGIN), (C.c_double * 3)(*scenes.SWEEP_TARGET)
    st = torch.cuda.current_stream().cuda_stream

    def sweep(first, n, order, count=0):
        cabi.check(dev.lib.rsb_hit_sweep_dev(dev.ctx, acc.scene, C.c_void_p(st), int(n), int(first), C5["seed"], o3, t3, scenes.SWEEP_HALF,
                                             int(order), C.c_void_p(hits.data_ptr()), C.c_void_p(sum_t.data_ptr()), C.c_void_p(xr.data_ptr()),
                                             count))
    # warm-up: a full 64 Mi-query pass allocates the pipeline buffers; three of them (~0.3 s of device work) also bring the
    # clocks back up after the host-side scene build, during which the device sat idle
    for _ in range(3):
        sweep(0, 2 * 10**8, 0)
    torch.cuda.synchronize()
    out = {"workload": C5["name"], "n_gpus": cx.world_size, "scaling": "strong",
           "partition": "contiguous ray ranges per rank, no data-path collective" if cx.world_size > 1 else "single GPU", "sweeps": []}
    # random: every ray independent (incoherent), the library as shipped: each 4 Mi-ray pass is sorted on a coherence key
    # on the device first (rsb_set_query_reorder, on by default; the sort is inside the timed region);
    # random_unsorted: the same rays with the reordering switched off;
    # morton: the rays generated along the Morton curve of the window (coherent, like the pixels of an observer; never sorted)
    for order_name in ("random", "random_unsorted", "morton"):
        dev.set_query_reorder(order_name != "random_unsorted")
        for n in C5["rays"]:
            order = 0 if order_name != "morton" else max(1, int(math.log(n, 4)))
            lo = n * cx.rank // cx.world_size
            hi = n * (cx.rank + 1) // cx.world_size
            sweep(lo, min(hi - lo, 10**7), order)               # untimed: the first pass in a new mode pays one-off set-up
            # a 1e7-ray sweep is 5 ms of device time: it is timed seven times and the BEST reported, all seven kept in `ms_all`
            # (a driver round trip inside the call once put 10-90 ms stalls into some of them: profiles/README.md, round 2, item
            # 11); the 1e9-ray sweep is timed once
            times = []
            for _ in range(7 if n <= 10**8 else 1):
                hits.zero_()
                cx.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                gc.disable()                                    # (no collector pass between the call's return and e1.record())
                e0.record()
                sweep(lo, hi - lo, order)
                e1.record()
                gc.enable()
                cx.barrier()
                times.append(cx.reduce(e0.elapsed_time(e1), "max"))
            ms = min(times)
            h = cx.reduce(int(hits.item()), "sum", torch.int64)
            sweep(lo, min(hi - lo, 10**7), order, 1)            # counting pass on (a prefix of) this rank's rays
            torch.cuda.synchronize()
            c = dev.counters()
            per_ray = trace_algorithmic_bytes(c) / c["rays"]
            achieved = per_ray * (hi - lo) / (ms * 1e-3) / 1e9   # this rank's rays against the slowest rank's time: per-GPU figure
            out["sweeps"].append({"order": order_name, "rays": n, "ms": ms, "ms_all": times, "timing": "best of %d" % len(times), "Mrays_per_s": n / ms / 1e3, "hit_fraction": h / n,
                                  "roofline": {"bound": "hbm", "kernel": "k_rq_world", "achieved": achieved, "peak": cx.peak, "unit": "GB/s",
                                               "frac": achieved / cx.peak, "algorithmic_bytes_per_ray": per_ray,
                                               "per_ray": {k: c[k] / c["rays"] for k in ("branches", "leaves", "items", "prim_tests")}}})
    dev.set_query_reorder(True)
    acc.close()
    return out


def run_c3(cx, args):
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    torch = cx.torch
    world = scenes.prism_scene(api)
    world._device = cx.device
    cam, pipe = scenes.cornell_camera(api, world, pixels=(C3["pixels"], C3["pixels"]), samples=C3["spp"], bins=C3["bins"],
                                      spectral_rays=C3["rays"], path_weight=0.75)
    cam.transform = api.translate(0.3, 0.2, -2.2) * api.rotate(5, -3, 0)
    cam.rng_mode = cabi.RNG_MT19937_64 if args.rng == "mt" else cabi.RNG_PHILOX
    pipe.accumulate = False
    world.build_accelerator()
    best = None
    for it in range(2):                                          # first frame warms up; the second is reported
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cam.observe()
        torch.cuda.synchronize()
        best = (time.perf_counter() - t0, cam.ray_count)
    return {"workload": C3["name"] + ", %d samples per pixel and slice" % C3["spp"], "Mrays_per_s": best[1] / best[0] / 1e6,
            "frames_per_s": 1.0 / best[0], "ms_per_step": 1e3 * best[0], "rays_per_step": best[1], "n_gpus": 1,
            "path": "mirror PinholeCamera.observe(): %d spectral slices, host frame assembled per slice; wall clock" % C3["rays"]}


def run_ours(args):
    cx = Ctx(args)
    torch, dist = cx.torch, cx.dist
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.distributed import FrameRenderer

    w = dict(WORKLOAD)
    if args.pixels:
        w["pixels"] = args.pixels
    if args.spp:
        w["spp"] = args.spp
    if not args.passes:
        args.passes = auto_passes(w["spp"], cx.world_size)
    if w["spp"] % args.passes:
        raise SystemExit("bench.py: --passes must divide the samples per pixel")
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(w["pixels"], w["pixels"]), samples=w["spp"], bins=w["bins"],
                                      path_weight=RAY_CFG["important_path_weight"])
    cam.rng_mode = cabi.RNG_MT19937_64 if args.rng == "mt" else cabi.RNG_PHILOX
    pipe.accumulate = False
    world._device = cx.device
    accel = world.build_accelerator()
    renderer = FrameRenderer(cam, accel, cx.rank, cx.world_size, tile=16, passes=args.passes)

    # ---- device-resident timing -------------------------------------------------------------------
    clocks = ClockSampler(cx.local_rank)
    if cx.rank == 0:
        clocks.start()
    total_ms, total_rays, per_step = timed_frames(cx, renderer, args.steps, args.warmup, 1)
    if cx.rank == 0:
        clocks.stop_flag.set()
    value = total_rays / total_ms / 1e3   # Mrays/s
    ms_step = total_ms / args.steps

    # ---- end to end with host buffers through the C ABI (every N) ---------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    renderer.step_host(seed=7)   # warm-up (allocates pinned buffers)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_rays = 0
    e0.record()
    for i in range(e2e_steps):
        t_rays += renderer.step_host(seed=200 + i)
    e1.record()
    cx.barrier()
    ms2 = cx.reduce(e0.elapsed_time(e1), "max")
    r2 = cx.reduce(t_rays, "sum", torch.int64)
    e2e_cabi = {"value": r2 / ms2 / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": renderer.h2d_bytes, "d2h_bytes_per_step": renderer.d2h_bytes,
                "steps": e2e_steps, "path": "FrameRenderer.step_host: rsb_render_passes_dev + frame assembly + pinned device->host copy; CUDA events"}

    # ---- roofline: counting pass (untimed) -----------------------------------------------------------
    roof = frame_roofline(cx, renderer, ms_step, per_step, w["bins"], w["spp"], 100, "k_wf_trace")
    del renderer
    torch.cuda.empty_cache()

    # ---- the drop-in plugin seam (N = 1) ---------------------------------------------------------------
    e2e = None
    if cx.world_size == 1 and not args.no_plugin:
        try:
            e2e = plugin_e2e(cx, w, args, e2e_steps)
        except Exception as exc:   # noqa: BLE001
            e2e_cabi["plugin_unavailable"] = repr(exc)
    if e2e is None:
        e2e = e2e_cabi

    # ---- the other configurations of the metric --------------------------------------------------------
    configs = {}
    which = [] if args.configs == "none" else args.configs.split(",")
    # (C5 first: the shortest timed regions of the line, measured before the multi-GB workloads churn the allocator)
    for key, fn in (("C5", run_c5), ("C4", run_c4), ("C3", run_c3)):
        if key not in which or (key == "C3" and cx.world_size > 1):
            continue
        # (garbage of the previous workload -- scenes holding GB of device memory -- is collected HERE, not by a collector
        # run that lands inside the next workload's timed region: its cudaFree calls stall the device)
        gc.collect()
        torch.cuda.empty_cache()
        try:
            configs[key] = fn(cx, args)
        except Exception as exc:   # noqa: BLE001
            configs[key] = {"failed": repr(exc)}
        gc.collect()
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1 only): the compiled reference in a clean subprocess --------------------
    cpu = None
    if cx.rank == 0 and cx.world_size == 1 and not args.no_cpu:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                  "--configs", args.configs],
                                 capture_output=True, text=True, timeout=1200).stdout.strip().splitlines()
            ref = json.loads(out[-1])
            cpu = ref.get("cpu_baseline", {"unavailable": ref.get("unavailable")})
            for key, val in (cpu.pop("configs", None) or {}).items():
                if key in configs:
                    configs[key]["cpu_reference"] = val
        except Exception as exc:   # noqa: BLE001
            cpu = {"unavailable": repr(exc)}

    if cx.rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": cx.world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"] if not (args.pixels or args.spp) else "cornell_box %dx%d x %d spp x %d bins" % (w["pixels"], w["pixels"], w["spp"], w["bins"]),
                       "passes": "%d spp accumulated as %d observe() passes of %d spp (streams keyed on (pass, pixel)), merged with "
                                 "StatsArray3D.combine_samples" % (w["spp"], args.passes, w["spp"] // args.passes)
                                 if args.passes > 1 else "one observe() pass of %d spp" % w["spp"],
                       "rng": "mt19937_64 per (pass, pixel)" if args.rng == "mt" else "philox4x32-10 per (pass, pixel, sample)",
                       "partition": "16x16 px tiles interleaved over ranks; owned tiles gathered on rank 0 over NCCL" if cx.world_size > 1 else "single GPU",
                       "l2": "frame buffers 1.07 GB per step exceed the 126 MB L2; scene (4.5 KB) is shared-memory resident by design"},
            "frames_per_s": 1e3 / ms_step, "rays_per_step": total_rays / args.steps,
            "gpu_launches": int(per_step["launches"] * args.steps), "waves_per_step": per_step["waves"],
            "e2e": e2e, "e2e_cabi": e2e_cabi, "roofline": roof, "configs": configs, "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if cx.world_size > 1:
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
    ap.add_argument("--passes", type=int, default=0,
                    help="render the frame's samples as this many accumulated observe() passes, concurrently "
                         "(default: 8 per GPU, at least 4 samples per pass)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--configs", default="C4,C5,C3", help="comma list of the extra configurations to measure, or 'none'")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-plugin", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
