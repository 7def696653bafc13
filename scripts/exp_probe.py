"""Experiment probe (development aid): stage timings of the staged K1 on config 2 under RzTuning settings and, optionally,
an alternative build of the library (scripts/exp_build.sh puts variants under scripts/_build/exp/).

    python scripts/exp_probe.py [--so scripts/_build/exp/NAME.so] [--spp 500] [--set field=value[,field=value...]]... [--glass]

Each --set group is one measurement: overlapped render (Mpaths/s) + serial-pass stage breakdown + tests per segment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import _abi as abi
from rayz_b200 import Backend

args = sys.argv[1:]
so = args[args.index("--so") + 1] if "--so" in args else None
spp = int(args[args.index("--spp") + 1]) if "--spp" in args else 500
groups = [args[i + 1] for i, a in enumerate(args) if a == "--set"] or [""]
if so:
    abi.SO_PATH = os.path.abspath(so)
tag = os.path.basename(so) if so else "default"
t = rayz_b200.random_bouncing(1200, seed=42, glass_heavy="--glass" in args)
for g in groups:
    be = Backend((0,))
    kv = {}
    for item in filter(None, g.split(",")):
        k, v = item.split("=")
        kv[k] = float(v) if "." in v else int(v)
    if kv:
        be.set_tuning(**kv)
    be.upload_scene(t.pool.arrays())
    out = []
    for serial in (False, True):
        p = Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega", serial_passes=serial)
        be.render_device(t.camera.rz, p)
        best = None
        for _ in range(3):
            be.render_device(t.camera.rz, p)
            ti = be.timing()
            if best is None or ti["kernel_ms"] < best["kernel_ms"]:
                best = ti
        out.append(best)
    o, s = out
    be.render_device(t.camera.rz, Backend.params(t.img.w, t.img.h, spp, 50, seed=1, variant="mega", collect_stats=True))
    st = [be.stage_stats(k) for k in range(3)]
    tail = s["kernel_ms"] - s["primary_ms"] - s["second_ms"] - s["sort_ms"]
    print(f"[{tag}] {g or 'defaults'}: overlap {o['kernel_ms']:.2f} ms = {t.img.w * t.img.h * spp / o['kernel_ms'] / 1e3:.0f} Mpaths/s | serial {s['kernel_ms']:.2f}: "
          f"primary {s['primary_ms']:.2f} second {s['second_ms']:.2f} sort {s['sort_ms']:.2f} tail {tail:.2f} | passes {s['passes']} stages {s['sorted_stages']} | "
          f"tests/seg primary {st[0]['sphere_tests'] / max(1, st[0]['segments']):.1f} sorted {st[1]['sphere_tests'] / max(1, st[1]['segments']):.1f}", flush=True)
    be.close()
