"""Throughput sweep over kernel variants and tunings (development aid; prints one line each)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayz_b200
from rayz_b200 import Backend


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1200)
    ap.add_argument("--spp", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--glass", action="store_true")
    ap.add_argument("--grid", type=int, default=11)
    ap.add_argument("--bvh-build", default="auto", choices=["auto", "host", "device"])
    ap.add_argument("--configs", default="mega:1:16,mega:2:16,mega:2:8,mega:2:32,mega:1:32,bvh:1:16,wavefront:2:16")
    args = ap.parse_args()
    t = rayz_b200.random_bouncing(args.width, seed=42, glass_heavy=args.glass, grid_lo=-args.grid, grid_hi=args.grid)
    be = Backend((0,), bvh_build=args.bvh_build)
    be.upload_scene(t.pool.arrays())
    print(json.dumps({"bvh_build": args.bvh_build, "bvh_build_us": be.timing()["bvh_build_us"]}))
    W, H = t.img.w, t.img.h
    peak, sms = be.fp32_peak(300)
    print(json.dumps({"fp32_peak_tflops": peak, "sms": sms}))
    for cfg in args.configs.split(","):
        variant, rpt, chunk = cfg.split(":")
        be.set_tuning(int(rpt), int(chunk))
        p = Backend.params(W, H, args.spp, 50, seed=1, variant=variant)
        ps = Backend.params(W, H, 4, 50, seed=1, variant=variant, collect_stats=True)
        be.render_device(t.camera.rz, ps)
        st = be.stats()
        ti = be.timing()
        best = 1e30
        for _ in range(args.reps):
            be.render_device(t.camera.rz, p)
            best = min(best, be.timing()["kernel_ms"])
        paths = W * H * args.spp
        S = st["segments"] / st["paths"]
        f_isect = ti["n_static"] * 16 + ti["n_moving"] * 22
        n_sph = max(1, ti["n_static"] + ti["n_moving"])
        tests = st["sphere_tests"] / st["paths"] if variant != "bvh" and st["sphere_tests"] else S * n_sph   # bvh: brute-force EQUIVALENT flop
        f_path = f_isect * tests / n_sph + (S - st["ended_sky"] / st["paths"]) * 70 + 45 + 19 * st["ended_sky"] / st["paths"]
        tf = paths * f_path / (best * 1e-3) / 1e12
        print(json.dumps({"cfg": cfg, "kernel_ms": best, "mpaths_s": paths / best / 1e3, "seg_per_path": S,
                          "brute_equiv_tflops": tf, "frac_of_measured_peak": tf / peak,
                          "node_tests_per_seg": st["node_tests"] / max(1, st["segments"]),
                          "sphere_tests_per_seg": st["sphere_tests"] / max(1, st["segments"])}), flush=True)


if __name__ == "__main__":
    main()
