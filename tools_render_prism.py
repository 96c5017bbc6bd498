#!/usr/bin/env python
"""BASELINE config 3 on one GPU: the dispersive CSG prism of demos/prism.py (tests/scenes.prism_scene), PinholeCamera
512 x 512, 512 spectral bins traced as 512 spectral rays (one bin per slice), through the public mirror API
(PinholeCamera.observe -> rsb_render per slice).  Prints one JSON line with Mrays/s and seconds per frame.

    python tools_render_prism.py [--pixels 512] [--bins 512] [--rays 512] [--spp 2] [--passes 1]

Every slice is its own render call today (slices as concurrent work items is the next step);
this tool is the measurement that step will be judged against.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import torch
    import scenes
    import source_b200 as api
    from source_b200.engine import Device
    ap = argparse.ArgumentParser()
    ap.add_argument("--pixels", type=int, default=512)
    ap.add_argument("--bins", type=int, default=512)
    ap.add_argument("--rays", type=int, default=512, help="spectral rays = slices per observe()")
    ap.add_argument("--spp", type=int, default=2)
    ap.add_argument("--passes", type=int, default=1)
    args = ap.parse_args()
    device = Device(0)
    world = scenes.prism_scene(api)
    world._device = device
    cam, pipe = scenes.cornell_camera(api, world, pixels=(args.pixels, args.pixels), samples=args.spp, bins=args.bins,
                                      spectral_rays=args.rays, path_weight=0.75)
    cam.transform = api.translate(0.3, 0.2, -2.2) * api.rotate(5, -3, 0)
    pipe.accumulate = False
    world.build_accelerator()
    best = None
    for it in range(2):
        cam.seed = 1 + it
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cam.observe(passes=args.passes) if args.passes > 1 else cam.observe()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, cam.ray_count)
    print(json.dumps({"scene": "prism (CSG, dispersive)", "pixels": args.pixels, "bins": args.bins, "spectral_rays": args.rays,
                      "spp_per_slice": args.spp, "passes": args.passes, "s_per_frame": best[0], "rays": best[1],
                      "Mrays_per_s": best[1] / best[0] / 1e6, "mean_sum": float(pipe.frame.mean.sum())}))


if __name__ == "__main__":
    main()
