#!/usr/bin/env python
"""BASELINE config 4 geometry on one GPU: Cornell box + a 20*4^k-triangle closed mesh (k = 8: 1,310,720 triangles, SAH
tree built by this package), PinholeCamera, 64 spectral bins.  Prints one JSON line with Mrays/s and frames/s.

    python tools_render_mesh.py [--subdiv 8] [--pixels 1024] [--spp 64] [--rng mt|philox]
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
    from source_b200 import _cabi as cabi
    from source_b200.engine import Device, camera_desc, ray_config
    from source_b200.flatten import flatten_world
    ap = argparse.ArgumentParser()
    ap.add_argument("--subdiv", type=int, default=8)
    ap.add_argument("--pixels", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--bins", type=int, default=64)
    ap.add_argument("--rng", default="mt", choices=["mt", "philox"])
    ap.add_argument("--glass", action="store_true", help="keep the glass box and sphere of the Cornell scene")
    args = ap.parse_args()
    dev = Device(0)
    verts, tris, normals = scenes.icosphere(args.subdiv, radius=0.45, bumps=0.15)
    t0 = time.time()

    def extra(a, w):
        a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
               material=a.Lambert(a.ConstantSF(0.7)))
    world = scenes.cornell_box(api, glass=args.glass, extra=extra)
    build_s = time.time() - t0
    flat = flatten_world(world)
    acc = dev.build(world)
    N = args.pixels
    cam = camera_desc(N, N, args.spp, 45, 1.0, api.translate(0, 0, -3.3))
    cfg = ray_config(args.bins, 375.0, 740.0, 0.01, 3, 500, True, 0.25)
    sp = flat.spectral(375.0, 740.0, args.bins)
    mode = cabi.RNG_MT19937_64 if args.rng == "mt" else cabi.RNG_PHILOX
    mean = torch.zeros((N, N, args.bins), dtype=torch.float64, device="cuda")
    var = torch.zeros_like(mean)
    best = None
    for it in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        m, v, rays = acc.render_device(cam, cfg, sp, mode, 1 + it, None, mean, var)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if best is None or ms < best[0]:
            best = (ms, int(rays))
    rs = dev.render_stats()
    print(json.dumps({"scene": "cornell + %d-triangle mesh" % len(tris), "pixels": N, "spp": args.spp, "bins": args.bins, "rng": args.rng,
                      "mesh_kdtree_build_s": build_s, "ms": best[0], "rays": best[1], "Mrays_per_s": best[1] / best[0] / 1e3,
                      "frames_per_s": 1e3 / best[0], "waves": rs["waves"], "mean_sum": float(m.sum())}))


if __name__ == "__main__":
    main()
