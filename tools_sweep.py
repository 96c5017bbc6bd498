#!/usr/bin/env python
"""BASELINE config 5: ray-batch sweep, N primary rays over a 10,000-sphere random scene, rays generated on the
device from (seed, index) (rsb_hit_sweep_dev).  Prints one JSON line per N with Mrays/s and the algorithmic-bytes
roofline of the traversal kernel; optionally the same for the Cornell + mesh scene (config 4 geometry).

    python tools_sweep.py [--spheres 10000] [--n 1e5 1e6 1e7 1e8] [--mesh-subdiv 7]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import numpy as np
    import torch
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.engine import Device
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=10000)
    ap.add_argument("--n", type=float, nargs="+", default=[1e5, 1e6, 1e7, 1e8])
    ap.add_argument("--mesh-subdiv", type=int, default=0, help="also sweep a Cornell box holding a 20*4^k-triangle mesh")
    ap.add_argument("--no-spheres", action="store_true", help="skip the sphere field (profiling the mesh scene)")
    ap.add_argument("--order", nargs="+", default=["random", "sorted", "morton"], choices=["random", "sorted", "morton"],
                    help="random: every ray aims anywhere in the window (incoherent), query reordering switched off; sorted: the same "
                         "random rays with the library's default query reordering (rsb_set_query_reorder); morton: rays walk a grid of cells along the "
                         "Morton curve (coherent, like the pixels of an observer)")
    args = ap.parse_args()
    dev = Device(0)
    peak = 6554.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])

    def sweep(name, acc, origin, target, half, ns):
        for order_name in args.order:
            sweep_order(name, acc, origin, target, half, ns, order_name)

    def sweep_order(name, acc, origin, target, half, ns, order_name):
        dev.set_query_reorder(order_name != "random")
        hits = torch.zeros(1, dtype=torch.int64, device="cuda")
        sum_t = torch.zeros(1, dtype=torch.float64, device="cuda")
        xr = torch.zeros(1, dtype=torch.int64, device="cuda")
        o3, t3 = (C.c_double * 3)(*origin), (C.c_double * 3)(*target)
        st = torch.cuda.current_stream().cuda_stream

        def run(n, count):
            # "morton": a 2^g x 2^g grid of cells along the Morton curve with about one ray per cell (primary rays of an image)
            order = 0 if order_name in ("random", "sorted") else max(1, int(math.log(max(n, 4), 4)))
            hits.zero_(); sum_t.zero_(); xr.zero_()
            cabi.check(dev.lib.rsb_hit_sweep_dev(dev.ctx, acc.scene, C.c_void_p(st), int(n), 0, 2024, o3, t3, half, order,
                                                 C.c_void_p(hits.data_ptr()), C.c_void_p(sum_t.data_ptr()), C.c_void_p(xr.data_ptr()), count))
        for n in ns:
            n = int(n)
            run(min(n, 10**7), 0)           # warm-up (at size: the chunk buffers are allocated on first use)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 if n <= 10**8 else 1
            e0.record()
            for _ in range(reps):
                run(n, 0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            h = int(hits.item())
            run(min(n, 10**7), 1)           # counting pass on (a prefix of) the same rays
            torch.cuda.synchronize()
            c = dev.counters()
            per_ray = (72 * c["rays"] + 16 * c["branches"] + 8 * c["leaves"] + 4 * c["items"] + 128 * c["prim_tests"] + 48 * c["tri_tests"]) / c["rays"]
            achieved = per_ray * n / (ms * 1e-3) / 1e9
            print(json.dumps({"scene": name, "order": order_name, "rays": n, "ms": ms, "Mrays_per_s": n / ms / 1e3, "hit_fraction": h / n,
                              "algorithmic_bytes_per_ray": per_ray, "achieved_GBps": achieved, "roofline_frac": achieved / peak,
                              "per_ray": {k: c[k] / c["rays"] for k in ("branches", "leaves", "items", "prim_tests", "tri_tests")}}), flush=True)

    if not args.no_spheres:
        world = scenes.random_spheres(api, args.spheres, seed=7)
        t0 = time.time()
        acc = dev.build(world)
        print(json.dumps({"scene": "%d spheres" % args.spheres, "flatten_build_upload_s": time.time() - t0}), flush=True)
        sweep("%d spheres" % args.spheres, acc, (0, 0, -4.0), (0, 0, 0), 0.9, args.n)
        acc.close()

    if args.mesh_subdiv:
        verts, tris, normals = scenes.icosphere(args.mesh_subdiv, radius=0.45, bumps=0.15)
        t0 = time.time()

        def extra(a, w):
            a.Mesh(verts, tris, normals, smoothing=True, closed=True, parent=w, transform=a.translate(0.1, -0.5, 0.1) * a.rotate(20, 10, 0),
                   material=a.Lambert(a.ConstantSF(0.7)))
        world = scenes.cornell_box(api, glass=False, extra=extra)
        tb = time.time() - t0
        t0 = time.time()
        acc = dev.build(world)
        print(json.dumps({"scene": "cornell + %d-triangle mesh" % len(tris), "mesh_kdtree_build_s": tb, "flatten_upload_s": time.time() - t0}), flush=True)
        sweep("cornell + %d-triangle mesh" % len(tris), acc, (0, 0, -3.3), (0.1, -0.5, 0.1), 0.5, args.n)


if __name__ == "__main__":
    main()
