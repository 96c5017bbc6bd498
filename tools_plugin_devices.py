"""Development tool: the drop-in seam with several GPUs in ONE process -- raysect PinholeCamera.observe() with
camera.render_engine = CudaRenderEngine(devices=[0..n-1]) on the bench workload, wall clock, one JSON line per n.

    python tools_plugin_devices.py --devices 1 2 [--rgb]
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--rgb", action="store_true", help="feed an RGBPipeline2D next to the spectral pipeline")
    args = ap.parse_args()
    import torch
    import bench
    import scenes
    from oracle import harness
    api = harness.ref_api()
    from source_b200.engine import Device
    from source_b200.plugin import CudaRenderEngine, WholeFrameSampler2D
    w = bench.WORKLOAD
    devs = [Device(k) for k in range(min(max(args.devices), torch.cuda.device_count()))]
    for n in args.devices:
        if n > len(devs):
            continue
        world = scenes.cornell_box(api)
        cam, pipe = scenes.cornell_camera(api, world, pixels=(w["pixels"], w["pixels"]), samples=w["spp"], bins=w["bins"],
                                          path_weight=bench.RAY_CFG["important_path_weight"])
        pipe.accumulate = False
        if args.rgb:
            from raysect.optical.observer import RGBPipeline2D
            cam.pipelines = [pipe, RGBPipeline2D(display_progress=False)]
        cam.frame_sampler = WholeFrameSampler2D()
        passes = bench.auto_passes(w["spp"], n)
        eng = CudaRenderEngine(seed=1, rng="mt", passes=passes, **(dict(devices=devs[:n]) if n > 1 else dict(device=devs[0])))
        cam.render_engine = eng
        cam.observe()
        eng.ray_count = 0
        parts = {"render_s": 0.0, "update_s": 0.0}
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cam.observe()
            for k in parts:
                parts[k] += eng.timing.get(k, 0.0) / args.steps
        dt = (time.perf_counter() - t0) / args.steps
        print(json.dumps({"devices_in_one_process": n, "passes": passes, "pipelines": [type(p).__name__ for p in cam.pipelines],
                          "Mrays_per_s": eng.ray_count / args.steps / dt / 1e6, "s_per_observe": dt, "breakdown": parts,
                          "workload": w["name"]}), flush=True)


if __name__ == "__main__":
    main()
