"""Config 4 (99,856 spheres), staged BVH pipeline: camera-stage time and its counters per chunk size (development aid).
    python scripts/c4_diag.py [--so lib.so] [chunk ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import _abi as abi
from rayz_b200 import Backend

args = sys.argv[1:]
if "--so" in args:
    abi.SO_PATH = os.path.abspath(args[args.index("--so") + 1])
    del args[args.index("--so"):args.index("--so") + 2]
tag = os.path.basename(abi.SO_PATH)
t = rayz_b200.random_bouncing(1920, seed=42, grid_lo=-158, grid_hi=158)
for chunk in [int(c) for c in args] or [16, 64, 256]:
    be = Backend((0,))
    be.set_tuning(chunk_primary=chunk)
    be.upload_scene(t.pool.arrays())
    p = Backend.params(t.img.w, t.img.h, 256, 50, seed=1, variant="auto", serial_passes=True)
    be.render_device(t.camera.rz, p)
    best = None
    for _ in range(2):
        be.render_device(t.camera.rz, p)
        ti = be.timing()
        if best is None or ti["kernel_ms"] < best["kernel_ms"]:
            best = ti
    be.render_device(t.camera.rz, Backend.params(t.img.w, t.img.h, 256, 50, seed=1, variant="auto", collect_stats=True))
    st = be.stage_stats(0)
    seg = max(1, st["segments"])
    print(f"[{tag}] chunk_primary={chunk}: serial {best['kernel_ms']:.2f} ms, camera stage {best['primary_ms']:.2f} ms, passes {best['passes']} | camera stage per ray: "
          f"{st['sphere_tests'] / seg:.1f} sphere tests, {st['node_tests'] / seg:.2f} box tests ({seg} rays)", flush=True)
    be.close()
