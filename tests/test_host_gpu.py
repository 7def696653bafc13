"""GPU test of the C++ drop-in host (host/rayz_host.cpp, the stand-in for rayz.zig:12-43 — no zig toolchain in the image):
built here with the committed Makefile, run as the reference executable is run (`<img_w> [out.ppm]`), and checked on
  * its scene: the flattened randomBouncing it uploads equals the oracle's scene bytes (rayz.zig:45-168 + Zig-std PRNG restated
    a third time, in C++) — `--dump-scene`;
  * its pixels: the P3 file equals, byte for byte, image.zig:29-41's format applied to the RGB8 the C ABI returns for the same
    scene and seeds through the Python mirror, and the oracle's quantise of the returned linear floats (<= 1 LSB: the device
    quantises the f64 mean, the check re-quantises its float32 rounding);
  * its report line: "Finished render ({d:.2}s): {d:.2} rps and {d:.2} us per ray" (rayz.zig:30-34), rays = w*h*spp."""
import os
import re
import subprocess

import numpy as np
import pytest

import rayz_b200

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = (("sphere_center", "f8", 3, "s"), ("sphere_velocity", "f8", 3, "s"), ("sphere_radius", "f8", 1, "s"), ("sphere_material", "u4", 1, "s"),
          ("mat_kind", "u4", 1, "m"), ("mat_fuzz", "f8", 1, "m"), ("mat_ior", "f8", 1, "m"), ("mat_texture", "u4", 1, "m"), ("mat_method", "u4", 1, "m"),
          ("tex_kind", "u4", 1, "t"), ("tex_color", "f8", 3, "t"), ("tex_scale", "f8", 1, "t"), ("tex_even", "u4", 1, "t"), ("tex_odd", "u4", 1, "t"))


@pytest.fixture(scope="module")
def host_exe():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "host")], check=True)
    exe = os.path.join(ROOT, "host", "_build", "rayz_host")
    assert os.path.exists(exe)
    return exe


def read_scene_dump(path):
    raw = open(path, "rb").read()
    ns, nm, nt, _ = np.frombuffer(raw[:16], dtype="<u4")
    n_of, off, out = {"s": int(ns), "m": int(nm), "t": int(nt)}, 16, {}
    for name, dt, k, which in FIELDS:
        cnt = n_of[which] * k
        a = np.frombuffer(raw, dtype="<" + dt, count=cnt, offset=off)
        off += a.nbytes
        out[name] = a.reshape(-1, k) if k > 1 else a
    assert off == len(raw)
    return out


def test_cpp_host_is_a_drop_in_for_the_reference_executable(host_exe, orc, tmp_path):
    w, spp, seed = 160, 10, 42
    ppm, scn, lin = tmp_path / "out.ppm", tmp_path / "scene.bin", tmp_path / "lin.f32"
    r = subprocess.run([host_exe, str(w), str(ppm), "--spp", str(spp), "--seed", str(seed), "--dump-scene", str(scn), "--dump-linear", str(lin), "--ppm-bench"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    h = int(w / rayz_b200.ASPECT_RATIO)
    # ---- report line (rayz.zig:30-34): seconds, rays per second, microseconds per ray; rays = primary samples (renderer.zig:90)
    m = re.search(r"^Finished render \((\d+\.\d\d)s\): (\d+\.\d\d) rps and (\d+\.\d\d) us per ray$", r.stderr, re.M)
    assert m, r.stderr
    assert re.search(r"writePPM: \d+ bytes of P3 text in", r.stderr)
    # ---- scene bytes == the oracle's randomBouncing(seed)
    mine, ref = read_scene_dump(scn), orc.Scene.random_bouncing(seed).arrays()
    assert set(mine) == set(ref)
    for k in ref:
        assert mine[k].shape == ref[k].shape and np.array_equal(mine[k], ref[k]), k
    # ---- P3 bytes: the same render through the Python mirror (same scene, render seed 1, AUTO variant)
    t = rayz_b200.random_bouncing(w, seed=seed)
    t.samples_per_px = spp
    assert t.render() == w * h * spp
    rgb = t.img.rgb8.reshape(-1, 3)
    want = f"P3\n{w} {h}\n255\n" + "".join(f"{a} {b} {c}\n" for a, b, c in rgb.tolist())
    assert ppm.read_bytes() == want.encode()
    # ---- and the oracle's writePPM transform (image.zig:35-38) of the floats the host received
    linear = np.fromfile(lin, dtype="<f4").reshape(h, w, 4)
    assert np.array_equal(linear[..., :3].astype(np.float64).reshape(-1, 3), t.img.pixels)
    q = orc.quantise(linear[..., :3].astype(np.float64)).reshape(-1, 3)
    d = np.abs(q.astype(np.int32) - rgb.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 2e-3


def test_cpp_host_penultimate_scene_and_stdout(host_exe):
    """No output path => P3 on stdout (rayz.zig:36-42); --scene penultimate = rayz.zig:170-239 restated."""
    r = subprocess.run([host_exe, "64", "--spp", "4", "--seed", "1", "--scene", "penultimate"], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    lines = r.stdout.decode().split("\n")
    assert lines[:3] == ["P3", "64 36", "255"] and len(lines) == 3 + 64 * 36 + 1
    t = rayz_b200.penultimate_scene(64)
    t.samples_per_px = 4
    t.render()
    assert lines[3:-1] == [" ".join(str(int(x)) for x in px) for px in t.img.rgb8.reshape(-1, 3)]
    r = subprocess.run([host_exe], capture_output=True)
    assert r.returncode != 0                                    # the reference panics without argv[1] (rayz.zig:16)
