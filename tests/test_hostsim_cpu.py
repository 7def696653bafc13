"""CPU regression of the product's FP32 shading code (rz_device.cuh compiled as host code by
tests/hostsim) against the f64 oracle.  De-risks the GPU kernels in a container without a GPU;
the GPU parity tests (test_gpu_parity.py) are the real gate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rayz_b200
from rayz_b200.host import scene_struct
from metrics import compare

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostsim():
    subprocess.run(["bash", os.path.join(HERE, "hostsim", "build.sh")], check=True)
    lib = C.CDLL(os.path.join(HERE, "_build", "libhostsim.so"))
    lib.hostsim_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                   C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
    return lib


def run(lib, arrays, cam, w, h, spp, depth=50, seed=1):
    sc, keep = scene_struct(arrays)
    out = np.zeros((h, w, 3))
    cnt = np.zeros(10, dtype=np.uint64)
    lib.hostsim_render(C.addressof(sc), C.addressof(cam), w, h, spp, depth, seed, 0.0, 0, out.ctypes.data, cnt.ctypes.data)
    return out, cnt


def test_fp32_shading_matches_oracle_statistics(hostsim, orc):
    t = rayz_b200.random_bouncing(128, seed=42)
    arrays, w, h, spp = t.pool.arrays(), 128, 72, 48
    out, cnt = run(hostsim, arrays, t.camera.rz, w, h, spp)
    sc = orc.Scene.from_arrays(arrays)
    ocam, _ = orc.default_camera(w)
    a, st = sc.render(ocam, w, h, spp, 50, seed=11, threads=0, stats=True)
    b, _ = sc.render(ocam, w, h, spp, 50, seed=12, threads=0)
    floor, got = compare(b, a), compare(out, a)
    assert got["psnr"] >= floor["psnr"] - 0.4
    assert got["block_mae"] <= floor["block_mae"] * 1.25 + 2e-4
    assert max(abs(x) for x in got["mean_diff"]) < 2.5e-3
    paths = w * h * spp
    assert cnt[0] == paths and cnt[7] + cnt[8] + cnt[9] == paths
    assert abs(cnt[1] / paths - st["segments"] / paths) < 0.06
    assert abs(int(cnt[5]) - st["hits_metallic"]) / st["hits_metallic"] < 0.03


def test_fp32_depth_semantics(hostsim):
    t = rayz_b200.random_bouncing(32, seed=42)
    arrays = t.pool.arrays()
    out, cnt = run(hostsim, arrays, t.camera.rz, 32, 18, 2, depth=0)
    assert (out == 0).all() and cnt[1] == 0 and cnt[9] == 32 * 18 * 2       # renderer.zig:104-105
    out, cnt = run(hostsim, arrays, t.camera.rz, 32, 18, 2, depth=1)
    assert cnt[1] == 32 * 18 * 2                                            # one closest-hit query per path


@pytest.mark.parametrize("lo,hi,bits", [
    ((-11.6, -0.1, -11.6), (11.6, 2.1, 11.6), 9),      # the RTOW scene's box around the non-huge spheres: a thin slab
    ((-3.0, -3.0, -3.0), (3.0, 3.0, 3.0), 9),          # a cube: 3 bits per axis
    ((0.0, 0.0, 0.0), (100.0, 0.5, 1.0), 6),           # one long axis takes every bit
    ((-1.0, 2.0, -1.0), (1.0, 2.0, 1.0), 9),           # a degenerate (flat) box
    ((-5.0, -5.0, -5.0), (5.0, 5.0, 5.0), 0),          # no cell bits at all
])
def test_sort_key_bounds_contain_their_rays(hostsim, lo, hi, bits):
    """Staged K1: the sorted-stage kernel culls the sphere set from bounds decoded from the queue's 16-bit sort keys.
    For 2 M random rays in and around the box (faces, axis-parallel and grazing directions included) the decoded bounds
    must contain the ray that produced the key: origin cell, direction octant, reach.  (rz_device.cuh, compiled as host code.)"""
    hostsim.hostsim_key_check.restype = C.c_uint64
    hostsim.hostsim_key_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    l = np.array(lo, dtype=np.float32)
    h = np.array(hi, dtype=np.float32)
    counts = np.zeros(17, dtype=np.uint64)
    n = 2_000_000
    bad = hostsim.hostsim_key_check(l.ctypes.data, h.ctypes.data, bits, n, 7, counts.ctypes.data)
    assert bad == 0, f"{bad} of {n} rays fall outside the bounds of their own key"
    assert int(counts[:16].sum()) == n
    assert int((counts[:16] > 0).sum()) >= 6            # the reach classes are really exercised
    assert int(counts[16]) > (1 << bits) * 4            # and so are cells x octants
