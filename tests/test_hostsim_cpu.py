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
@pytest.mark.parametrize("key_mode", [-1, 0, 1, 2])    # direction field: by the box's shape (flat -> sectors), octants, 8 sectors, 16 sectors
def test_sort_key_bounds_contain_their_rays(hostsim, lo, hi, bits, key_mode):
    """Staged K1: the sorted-stage kernel culls the sphere set from bounds decoded from the queue's 16-bit sort keys.
    For 1 M random rays in and around the box (faces, axis-parallel and grazing directions included) the decoded bounds
    must contain the ray that produced the key: origin cell, direction octant, reach.  (rz_device.cuh, compiled as host code.)"""
    hostsim.hostsim_key_check.restype = C.c_uint64
    hostsim.hostsim_key_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    l = np.array(lo, dtype=np.float32)
    h = np.array(hi, dtype=np.float32)
    counts = np.zeros(17, dtype=np.uint64)
    n = 1_000_000
    hostsim.hostsim_set_key_mode(key_mode)
    try:
        bad = hostsim.hostsim_key_check(l.ctypes.data, h.ctypes.data, bits, n, 7, counts.ctypes.data)
    finally:
        hostsim.hostsim_set_key_mode(-1)
    assert bad == 0, f"{bad} of {n} rays fall outside the bounds of their own key"
    assert int(counts[:16].sum()) == n
    assert int((counts[:16] > 0).sum()) >= 6            # the reach classes are really exercised
    assert int(counts[16]) > (1 << min(bits, 8 if key_mode == 2 else 9)) * 4            # and so are cells x octants


def _cam(width, defocus_angle, look_from=(13, 2, 3), look_at=(0, 0, 0), vfov=20.0, focus=10.0):
    h = int(width / rayz_b200.host.ASPECT_RATIO)
    return rayz_b200.Camera.init(vfov, focus, defocus_angle, look_from, look_at, (0, 1, 0), h, width).rz, h


@pytest.mark.parametrize("width,defocus,look_from,vfov", [
    (1920, 0.6, (13, 2, 3), 20.0),        # config 4's camera
    (1205, 0.0, (13, 2, 3), 20.0),        # partial blocks on both edges, pinhole
    (203, 2.0, (3, 3, 2), 60.0),          # wide blur, wide field
])
def test_block_cull_of_the_bvh_camera_stage_never_drops_a_box_or_a_sphere(hostsim, width, defocus, look_from, vfov):
    """K3's camera stage culls the tree once per 8 x 4-pixel block against the block's cone (rz_tile_keep_box, the function the
    kernel calls, on the block layout the kernel uses): whatever a camera ray of the block hits must be kept, and so must every
    box around it — the sphere's own box over the shutter and enclosing boxes up to thousands of units — or the walk down the
    tree would lose the sphere before its own test."""
    hostsim.hostsim_block_cull_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    h = int(width / rayz_b200.ASPECT_RATIO)
    cam = rayz_b200.Camera.init(vfov, 10.0, defocus, look_from, (0, 0, 0), (0, 1, 0), h, width).rz
    out = np.zeros(5, dtype=np.uint64)
    assert hostsim.hostsim_block_cull_check(C.addressof(cam), width, h, 3000, 5, out.ctypes.data) == 0
    dropped, hits, boxes_dropped, boxes, culled = (int(x) for x in out)
    assert hits > 100_000 and boxes == 8 * hits
    assert dropped == 0, f"{dropped} of {hits} hit spheres culled"
    assert boxes_dropped == 0, f"{boxes_dropped} of {boxes} boxes around hit spheres culled"
    assert culled > 1_000                 # and the cone does cull


@pytest.mark.parametrize("width,defocus,look_from,vfov", [
    (1200, 0.6, (13, 2, 3), 20.0),        # the benchmark camera (rayz.zig:152-160)
    (400, 0.0, (13, 2, 3), 20.0),         # pinhole
    (160, 8.0, (0, 1.0, 8.0), 50.0),      # wide lens, wide field of view, coarse pixels
    (33, 2.0, (3, 3, 2), 90.0),           # tiles that wrap around image rows
])
def test_primary_tile_cull_never_drops_a_hit_sphere(hostsim, width, defocus, look_from, vfov):
    """Staged K1, primary kernel: the cone cull of a 32-pixel tile (rz_tile_keep, the function the kernel calls) must keep
    every sphere — moving ones included — that a camera ray of the tile hits."""
    hostsim.hostsim_tile_cull_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
    cam, h = _cam(width, defocus, look_from=look_from, vfov=vfov)
    out = np.zeros(4, dtype=np.uint64)
    assert hostsim.hostsim_tile_cull_check(C.addressof(cam), width, h, 4000, 3, out.ctypes.data) == 0
    bad, hits, kept, culled_misses = (int(x) for x in out)
    assert hits > 200_000 and kept == hits and bad == 0, (bad, hits)
    if width >= 400:
        assert culled_misses > 1_000      # and it does cull, even though every test sphere sits next to a ray of the tile


@pytest.mark.parametrize("lo,hi,bits,huge", [
    ((-11.6, -0.1, -11.6), (11.6, 2.1, 11.6), 9, 1.6),
    ((-6.0, -0.5, -6.0), (6.0, 7.0, 9.5), 9, 4.8),
    ((-3.0, -3.0, -3.0), (3.0, 3.0, 3.0), 5, 1.0),
    ((0.0, 0.0, 0.0), (100.0, 0.5, 1.0), 9, 0.2),
])
@pytest.mark.parametrize("key_mode", [0, 1, 2])
def test_sorted_unit_cull_never_drops_a_hit_sphere(hostsim, lo, hi, bits, huge, key_mode):
    """Staged K1, sorted-stage kernel: bounds merged from the keys of a unit's rays (rz_unit_bounds_add_key) and the cull
    built on them (rz_unit_keep) must keep every sphere inside the sphere box that one of the rays hits, and every huge one."""
    hostsim.hostsim_unit_cull_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_uint64, C.c_void_p]
    l = np.array(lo, dtype=np.float32)
    h = np.array(hi, dtype=np.float32)
    out = np.zeros(4, dtype=np.uint64)
    hostsim.hostsim_set_key_mode(key_mode)
    try:
        assert hostsim.hostsim_unit_cull_check(l.ctypes.data, h.ctypes.data, bits, huge, 20000, 5, out.ctypes.data) == 0
    finally:
        hostsim.hostsim_set_key_mode(-1)
    bad, hits, kept, culled_misses = (int(x) for x in out)
    assert hits > 100_000 and kept == hits and bad == 0, (bad, hits)
    if hi[1] - lo[1] > 1.0:               # (in the 0.5-high box hardly any test sphere fits inside)
        assert culled_misses > 5_000      # and it does cull (the spheres placed behind coherent units)


@pytest.mark.parametrize("glass", [False, True])
def test_staged_searches_equal_brute_force_on_the_benchmark_scene(hostsim, glass):
    """The staged K1's claim, stage by stage, on the RTOW scene (moving spheres, the r = 1000 ground, thin lens): a camera
    ray searched over its tile's culled list, and a scattered ray searched over the culled list of its sorted unit, find
    exactly the closest hit (t and sphere) of the search over every sphere."""
    hostsim.hostsim_staged_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
    t = rayz_b200.random_bouncing(400, seed=42, glass_heavy=glass)
    sc, keep = scene_struct(t.pool.arrays())
    out = np.zeros(7, dtype=np.uint64)
    assert hostsim.hostsim_staged_check(C.addressof(sc), C.addressof(t.camera.rz), 400, 225, 4, 3, 17, out.ctypes.data) == 0
    bad1, rays1, list1, bad2, rays2, list2, units = (int(x) for x in out)
    n = len(t.pool.arrays()["sphere_radius"])
    assert rays1 > 100_000 and rays2 > 50_000 and units > 100          # (units: the non-empty sort groups)
    assert bad1 == 0, f"{bad1} of {rays1} camera rays find a different hit over the tile list"
    assert bad2 == 0, f"{bad2} of {rays2} scattered rays find a different hit over the unit list"
    assert list1 / rays1 < 0.15 * n          # the tile cull keeps a small part of the set ...
    assert list2 / rays2 < 0.12 * n          # ... and a sort group's list about 7 % of it (10 % on the GPU, which tests a whole batch up to its largest class)


def test_bvh_camera_stage_block_lists_find_the_brute_force_hit(hostsim):
    """K3's camera stage, on the CPU: the library's own binned-SAH tree culled against the cone of 8 x 4-pixel blocks the way
    rz_bvh_stage_kernel does it, and the blocks' camera rays searched over the surviving spheres and over every sphere: same
    closest hit (t and sphere), on the RTOW scene and on a 2,500-sphere version of config 4's scene."""
    hostsim.hostsim_bvh_block_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                                C.c_void_p]
    for grid, width, stride in ((11, 480, 37), (25, 960, 211)):
        t = rayz_b200.random_bouncing(width, seed=42, grid_lo=-grid, grid_hi=grid)
        arrays = t.pool.arrays()
        sc, keep = scene_struct(arrays)
        out = np.zeros(6, dtype=np.uint64)
        assert hostsim.hostsim_bvh_block_check(C.addressof(sc), C.addressof(t.camera.rz), t.img.w, t.img.h, 6, stride, 256, 9, out.ctypes.data) == 0
        bad, rays, list_sum, blocks, over, frontier = (int(x) for x in out)
        n = len(arrays["sphere_radius"])
        assert rays > 10_000 and blocks > 50, (rays, blocks)
        assert bad == 0, f"{bad} of {rays} camera rays find a different hit over the block list ({n} spheres)"
        assert list_sum / blocks < 0.1 * n        # the cull keeps a small part of the set ...
        assert over <= 0.1 * blocks               # ... and few blocks outgrow the list (the kernel walks the tree there)
        assert frontier <= 256                    # the kernel's ring buffer
