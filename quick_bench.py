import sys, time
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import numpy as np, torch
import scenes
import source_b200 as api
from source_b200 import _cabi as cabi
from source_b200.engine import Device, camera_desc, ray_config
from source_b200.flatten import flatten_world
dev = Device(0)
world = scenes.cornell_box(api)
flat = flatten_world(world)
acc = dev.build(world)
N, BINS = int(sys.argv[1]), 64
for mode, name in ((cabi.RNG_PHILOX, "philox"), (cabi.RNG_MT19937_64, "mt")):
    for spp in (int(sys.argv[2]),):
        cam = camera_desc(N, N, spp, 45, 1.0, api.translate(0, 0, -3.3))
        cfg = ray_config(BINS, 375.0, 740.0, 0.01, 3, 500, True, 0.25)
        sp = flat.spectral(375.0, 740.0, BINS)
        mean = torch.zeros((N, N, BINS), dtype=torch.float64, device="cuda"); var = torch.zeros_like(mean)
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.time()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            m, v, rays = acc.render_device(cam, cfg, sp, mode, 1 + it, None, mean, var)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(name, "N", N, "spp", spp, "ms", round(ms, 2), "rays", int(rays), "Mrays/s", round(int(rays) / ms / 1e3, 1), "mean sum", float(m.sum()))
m, v, rays = acc.render_device(cam, cfg, sp, cabi.RNG_PHILOX, 3, None, mean, var, count=True)
print(dev.counters())
