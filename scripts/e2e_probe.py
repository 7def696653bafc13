"""Where the end-to-end call spends its time beyond the kernels (development aid): wall clock of rayz_cuda_upload_scene and of
rayz_cuda_render with pinned host buffers on config 2, next to the library's own RzTiming of the same render.

    python scripts/e2e_probe.py [--build auto|host|device] [--reps 8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rayz_b200
from rayz_b200 import Backend

build = sys.argv[sys.argv.index("--build") + 1] if "--build" in sys.argv else "auto"
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 8
t = rayz_b200.random_bouncing(1200, seed=42)
W, H = t.img.w, t.img.h
be = Backend((0,), bvh_build=build)
scene = t.pool.arrays()
p = Backend.params(W, H, 500, 50, seed=1)
lin = torch.empty((H, W, 4), dtype=torch.float32).pin_memory().numpy()
rgb = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory().numpy()
for _ in range(2):
    be.upload_scene(scene)
    be.render(t.camera.rz, p, out_linear=lin, out_rgb8=rgb)
up, rn, both, km, tm, bu = [], [], [], [], [], []
for _ in range(reps):
    t0 = time.perf_counter()
    be.upload_scene(scene)
    t1 = time.perf_counter()
    be.render(t.camera.rz, p, out_linear=lin, out_rgb8=rgb)
    t2 = time.perf_counter()
    ti = be.timing()
    up.append((t1 - t0) * 1e3); rn.append((t2 - t1) * 1e3); both.append((t2 - t0) * 1e3)
    km.append(ti["kernel_ms"]); tm.append(ti["total_ms"]); bu.append(ti["bvh_build_us"])
med = lambda v: sorted(v)[len(v) // 2]
print(f"[build={build}] upload_scene {med(up):.3f} ms (bvh build {med(bu)} us) | render(host buffers) {med(rn):.3f} ms wall = kernels {med(km):.3f} "
      f"+ rest {med(rn) - med(km):.3f} (library total_ms {med(tm):.3f}) | both {med(both):.3f} ms = {W * H * 500 / med(both) / 1e3:.0f} Mpaths/s", flush=True)
be.close()
