"""Experiment probe (development aid): timing of the staged K1 under the env knobs rz_context.cu reads per render,
and a bitwise check of the BVH tail against the brute-force tail.  `--quick` = timing of the default setting only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rayz_b200
from rayz_b200 import Backend

quick = "--quick" in sys.argv
tag = [a for a in sys.argv[1:] if not a.startswith("--")]
tag = tag[0] if tag else "default"

def setenv(**kw):
    for k in ("RZ_TAIL", "RZ_SECOND_STAGES", "RZ_BVH_ACTIVE_MIN", "RZ_BVH_DESCEND_MIN", "RZ_SORT_GRAPH", "RZ_CELL_BITS", "RZ_QUEUE_LOG2"):
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ[k] = str(v)

def timing(be, t, spp=500, **env):
    setenv(**env)
    out = []
    for serial in (False, True):
        p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega", serial_passes=serial)
        be.render_device(t.camera.rz, p); be.render_device(t.camera.rz, p)
        ti = be.timing()
        out.append(ti)
    o, s = out
    tail = s["kernel_ms"] - s["primary_ms"] - s["second_ms"] - s["sort_ms"]
    print(f"[{tag}] {env} overlap {o['kernel_ms']:.2f} ms = {t.img.w*t.img.h*spp/o['kernel_ms']/1e3:.0f} Mpaths/s | serial {s['kernel_ms']:.2f}: "
          f"primary {s['primary_ms']:.2f} second {s['second_ms']:.2f} sort {s['sort_ms']:.2f} tail {tail:.2f}", flush=True)

t = rayz_b200.random_bouncing(1200, seed=42)
be = Backend((0,)); be.upload_scene(t.pool.arrays())
def tests_per_segment(be, t, spp=500, **env):
    setenv(**env)
    p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega", collect_stats=True)
    be.render_device(t.camera.rz, p)
    st = [be.stage_stats(k) for k in range(3)]
    print(f"[{tag}] {env} tests/segment: primary {st[0]['sphere_tests']/max(1,st[0]['segments']):.1f} sorted {st[1]['sphere_tests']/max(1,st[1]['segments']):.1f} "
          f"(segments {st[1]['segments']/1e6:.0f} M) tail {st[2]['sphere_tests']/max(1,st[2]['segments']):.1f} + {st[2]['node_tests']/max(1,st[2]['segments']):.1f} boxes", flush=True)

if "--cells" in sys.argv:
    for cb in (9, 8, 7, 6, 5, 3):
        tests_per_segment(be, t, RZ_CELL_BITS=cb)
        timing(be, t, RZ_CELL_BITS=cb)
    for ns in (1, 2, 3, 4):
        tests_per_segment(be, t, RZ_SECOND_STAGES=ns)
    sys.exit(0)
if "--chunk" in sys.argv:
    for ch in (16, 32, 64, 125, 250):
        b = Backend((0,)); b.upload_scene(t.pool.arrays()); b.set_tuning(0, ch)
        print("chunk", ch, end=" ")
        timing(b, t)
        b.close()
    sys.exit(0)
if "--queue" in sys.argv:
    for q in (27, 28, 26):
        b = Backend((0,)); b.upload_scene(t.pool.arrays())     # buffers are sized at the first render of a context
        timing(b, t, RZ_QUEUE_LOG2=q)
        b.close()
    sys.exit(0)
if "--stages" in sys.argv:
    tg = rayz_b200.random_bouncing(1200, seed=42, glass_heavy=True)
    bg = Backend((0,)); bg.upload_scene(tg.pool.arrays())
    for ns in (3, 4, 5, 6):
        timing(be, t, RZ_SECOND_STAGES=ns)
        timing(bg, tg, RZ_SECOND_STAGES=ns)
    sys.exit(0)
if quick:
    os.environ["RZ_SORT_GRAPH_VERBOSE"] = "1"
    timing(be, t, RZ_SORT_GRAPH=0)
    timing(be, t, RZ_SORT_GRAPH=1)
    for ns in (2, 4):
        timing(be, t, RZ_SECOND_STAGES=ns)
    sys.exit(0)

# ---- bitwise: BVH tail vs brute-force tail (standard and glass-heavy scene, small image so that it is quick)
for glass in (False, True):
    ts = rayz_b200.random_bouncing(400, seed=42, glass_heavy=glass)
    b2 = Backend((0,)); b2.upload_scene(ts.pool.arrays())
    imgs = {}
    for tail in ("brute", "bvh"):
        setenv(RZ_TAIL=tail)
        p = Backend.params(ts.img.w, ts.img.h, 64, 50, seed=3, variant="mega")
        lin, _, n = b2.render(ts.camera.rz, p)
        imgs[tail] = lin.copy()
    d = np.any(imgs["brute"] != imgs["bvh"], axis=-1)
    print(f"[{tag}] glass={glass}: pixels differing bvh tail vs brute tail: {int(d.sum())} of {d.size}; max abs {float(np.abs(imgs['brute']-imgs['bvh']).max()):.3g}", flush=True)
    b2.close()

for ns in (0, 1, 2, 3):
    timing(be, t, RZ_TAIL="bvh", RZ_SECOND_STAGES=ns)
for am, dm in ((4, 24), (16, 24), (8, 16)):
    timing(be, t, RZ_TAIL="bvh", RZ_SECOND_STAGES=0, RZ_BVH_ACTIVE_MIN=am, RZ_BVH_DESCEND_MIN=dm)
be.close()
tg = rayz_b200.random_bouncing(1200, seed=42, glass_heavy=True)
bg = Backend((0,)); bg.upload_scene(tg.pool.arrays())
print("glass-heavy:")
for ns in (0, 1, 2, 3):
    timing(bg, tg, RZ_TAIL="bvh", RZ_SECOND_STAGES=ns)
