"""Renders one frame of the default scene with one variant (profiling aid: exactly one render, no stats pass)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="mega")
ap.add_argument("--width", type=int, default=1200)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--glass", action="store_true")
ap.add_argument("--grid", type=int, default=11)
a = ap.parse_args()
t = rayz_b200.random_bouncing(a.width, seed=42, glass_heavy=a.glass, grid_lo=-a.grid, grid_hi=a.grid)
be = Backend((0,))
be.upload_scene(t.pool.arrays())
be.render_device(t.camera.rz, Backend.params(t.img.w, t.img.h, a.spp, 50, seed=1, variant=a.variant))
print(be.timing())
