"""Experiment probe (development aid): the BVH kernel on config 2 (variant=bvh), as the staged K1's tail, and on config 4.
    python scripts/exp_bvh.py [--so scripts/_build/exp/NAME.so] [--only 2|4] [--build auto|host|device] [--set field=value,...]..."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import _abi as abi
from rayz_b200 import Backend

if "--so" in sys.argv:
    abi.SO_PATH = os.path.abspath(sys.argv[sys.argv.index("--so") + 1])
tag = os.path.basename(abi.SO_PATH) if "--so" in sys.argv else "default"
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""

build = sys.argv[sys.argv.index("--build") + 1] if "--build" in sys.argv else "auto"   # host: binned SAH even for the 100k scene
groups = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--set"] or [""]
scenes = [("config2 485 spheres", rayz_b200.random_bouncing(1200, seed=42), 500, "bvh"),
          ("config4 99,856 spheres", rayz_b200.random_bouncing(1920, seed=42, grid_lo=-158, grid_hi=158), 256, "auto")]
scenes = [sc for sc in scenes if not only or sc[0].startswith("config" + only)]
for g in groups:
    kv = {}
    for item in filter(None, g.split(",")):
        k, v = item.split("=")
        kv[k] = float(v) if "." in v else int(v)
    for name, t, spp, variant in scenes:
        be = Backend((0,), bvh_build=build)
        if kv:
            be.set_tuning(**kv)
        be.upload_scene(t.pool.arrays())
        p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant=variant)
        be.render_device(t.camera.rz, p)
        best = min(be.render_device(t.camera.rz, p) and be.timing()["kernel_ms"] for _ in range(3))
        be.render_device(t.camera.rz, Backend.params(t.img.w, t.img.h, max(1, spp // 8), 50, seed=1, variant=variant, collect_stats=True))
        st = be.stats()
        print(f"[{tag}: {g or 'defaults'}, build={build}] {name}: {best:.2f} ms = {t.img.w * t.img.h * spp / best / 1e3:.0f} Mpaths/s | per segment: {st['node_tests'] / st['segments']:.1f} box tests, "
              f"{st['sphere_tests'] / st['segments']:.2f} sphere tests | build {be.timing()['bvh_build_us']} us", flush=True)
        be.close()
