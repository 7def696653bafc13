import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend
glass = "--glass" in sys.argv
t = rayz_b200.random_bouncing(1200, seed=42, glass_heavy=glass)
be = Backend((0,)); be.upload_scene(t.pool.arrays())
for spp, serial in ((500, False), (500, True)):
    p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega", serial_passes=serial)
    be.render_device(t.camera.rz, p); be.render_device(t.camera.rz, p)
    ti = be.timing()
    print("serial" if serial else "overlap", spp, {k: round(v,2) if isinstance(v,float) else v for k,v in ti.items() if k in ('kernel_ms','primary_ms','second_ms','passes','launches')}, round(t.img.w*t.img.h*spp/ti['kernel_ms']/1e3,1))
ps = Backend.params(t.img.w, t.img.h, 8, 50, seed=1, variant="mega", collect_stats=True)
be.render_device(t.camera.rz, ps); st = be.stats()
print({k: round(st[k]/st['paths'],3) for k in ('segments','sphere_tests')})
