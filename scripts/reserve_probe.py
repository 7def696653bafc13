"""Times rayz_cuda_reserve (queue allocation + sort-graph construction) with and without the device-sized sort."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend
t = rayz_b200.random_bouncing(1200, seed=42)
for g in ("0", "1", "1"):
    os.environ["RZ_SORT_GRAPH"] = g
    t0 = time.perf_counter(); be = Backend((0,)); t1 = time.perf_counter()
    p = Backend.params(t.img.w, t.img.h, 500, 50, seed=1, variant="auto")
    be.reserve(p); t2 = time.perf_counter()
    be.upload_scene(t.pool.arrays()); t3 = time.perf_counter()
    be.render_device(t.camera.rz, p); t4 = time.perf_counter()
    be.render_device(t.camera.rz, p); t5 = time.perf_counter()
    be.close(); t6 = time.perf_counter()
    print(f"RZ_SORT_GRAPH={g}: create {t1-t0:.3f} s, reserve {t2-t1:.3f} s, upload {t3-t2:.3f} s, render#1 {t4-t3:.3f} s, render#2 {t5-t4:.3f} s, close {t6-t5:.3f} s", flush=True)
