import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend
t = rayz_b200.random_bouncing(1200, seed=42)
be = Backend((0,)); be.upload_scene(t.pool.arrays())
for spp in (500,):
    p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega")
    be.render_device(t.camera.rz, p); be.render_device(t.camera.rz, p)
    ti = be.timing()
    print(os.environ.get("RZ_QUEUE_LOG2"), spp, {k: round(v,2) if isinstance(v,float) else v for k,v in ti.items() if k in ('kernel_ms','primary_ms','passes','launches')}, round(t.img.w*t.img.h*spp/ti['kernel_ms']/1e3,1))
