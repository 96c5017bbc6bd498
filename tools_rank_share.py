"""Development tool (one GPU): renders rank 0's share of an N-GPU bench frame ALONE -- same tile partition, same passes,
no collective -- so that what an 8-GPU frame costs each GPU can be studied at the price of one.  Prints one JSON line
per (world, passes) pair: ms per step, waves, the per-phase kernel times of one instrumented frame.

    python tools_rank_share.py --world 8 --passes 32 64 128
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, nargs="+", default=[8])
    ap.add_argument("--passes", type=int, nargs="+", default=[0])
    ap.add_argument("--rank", type=int, nargs="+", default=[0])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import bench
    import scenes
    import source_b200 as api
    from source_b200 import _cabi as cabi
    from source_b200.distributed import FrameRenderer
    from source_b200.engine import Device

    w = bench.WORKLOAD
    dev = Device(0)
    torch.cuda.set_device(0)
    world = scenes.cornell_box(api)
    cam, pipe = scenes.cornell_camera(api, world, pixels=(w["pixels"], w["pixels"]), samples=w["spp"], bins=w["bins"],
                                      path_weight=bench.RAY_CFG["important_path_weight"])
    cam.rng_mode = cabi.RNG_MT19937_64
    pipe.accumulate = False
    world._device = dev
    accel = world.build_accelerator()
    for n in args.world:
        for p in args.passes:
          for rank in args.rank:
            passes = p or bench.auto_passes(w["spp"], n)
            if w["spp"] % passes or rank >= n:
                continue
            r = FrameRenderer(cam, accel, rank, n, tile=16, passes=passes, backend_reduce=lambda stats: stats)
            for i in range(args.warmup):
                r.step_device(seed=1 + i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            rays = torch.zeros(1, dtype=torch.int64, device="cuda:0")
            waves = 0
            e0.record()
            for i in range(args.steps):
                rays += r.step_device(seed=101 + i)
                waves += dev.render_stats()["waves"]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            r.step_device(seed=101, time_trace=True)
            torch.cuda.synchronize()
            rs = dev.render_stats()
            ph = {k: round(rs[k + "_ms"], 3) for k in ("trace", "shade", "finalize", "regen")}
            print(json.dumps({"world": n, "rank": rank, "deal": os.environ.get("RSB_TILE_DEAL", "diagonal"), "passes": passes, "spp_per_pass": w["spp"] // passes, "ms_per_step": round(ms, 3),
                              "Mrays_per_s_this_gpu": round(rays.item() / args.steps / ms / 1e3, 1),
                              "waves_per_step": waves / args.steps, "phases_ms": ph, "outside_wave_kernels_ms": round(ms - sum(ph.values()), 3)}),
                  flush=True)
            del r
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
