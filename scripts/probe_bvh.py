import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 11
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 500
t = rayz_b200.random_bouncing(w, seed=42, grid_lo=-grid, grid_hi=grid)
be = Backend((0,)); be.upload_scene(t.pool.arrays())
p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="bvh")
be.render_device(t.camera.rz, p); be.render_device(t.camera.rz, p)
ti = be.timing()
print(os.environ.get("RZ_BVH_NO_STAGES"), os.environ.get("RZ_BVH_STAGES"), {k: round(v,2) if isinstance(v,float) else v for k,v in ti.items() if k in ('kernel_ms','primary_ms','second_ms','sort_ms','passes','launches')}, round(t.img.w*t.img.h*spp/ti['kernel_ms']/1e3,1))
