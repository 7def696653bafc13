"""The host-side tree builders of rayz_cuda_upload_scene (rayz_b200/csrc/rz_host_bvh.hpp, compiled on their own by
tests/hostsim/hostbvh.cu): the binned-SAH BVH2 the FP32 traversal kernel walks on scenes below 8192 spheres, the
reference-shaped BVH of the bit-exact primary-id kernel (BVH.build, reference src/hit.zig:130-161) and the outward rounding of
f64 boxes to FP32.  Structural invariants only — whether a tree yields the right hits is the GPU parity tests' business
(LBVH == SAH == brute force, bitwise)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rayz_b200

HERE = os.path.dirname(os.path.abspath(__file__))

NODE = np.dtype([("lox", "<f4", 2), ("hix", "<f4", 2), ("loy", "<f4", 2), ("hiy", "<f4", 2), ("loz", "<f4", 2), ("hiz", "<f4", 2),
                 ("child", "<i4", 2), ("cnt", "<u4", 2)])            # RzBvhNode (rz_device.cuh), 64 bytes
REFNODE = np.dtype([("low", "<f8", 3), ("high", "<f8", 3), ("left", "<i4"), ("right", "<i4"), ("start", "<i4"), ("end", "<i4")])   # RzRefNode


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["bash", os.path.join(HERE, "hostsim", "build_hostbvh.sh")], check=True)
    return C.CDLL(os.path.join(HERE, "_build", "libhostbvh.so"))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def spheres(arrays):
    c = np.ascontiguousarray(arrays["sphere_center"], dtype=np.float64).reshape(-1, 3)
    v = np.ascontiguousarray(arrays["sphere_velocity"], dtype=np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(arrays["sphere_radius"], dtype=np.float64).reshape(-1)
    return c, v, r


def boxes(lib, c, v, r):
    n = len(r)
    lo, hi = np.empty((n, 3)), np.empty((n, 3))
    lib.hostbvh_sphere_boxes(C.c_uint32(n), ptr(c), ptr(v), ptr(r), ptr(lo), ptr(hi))
    return lo, hi


def sah(lib, c, v, r, leaf=4, node_cost=0.5):
    n = len(r)
    nodes = np.zeros(max(n, 1), dtype=NODE)
    order = np.full(max(n, 1), 0xffffffff, dtype=np.uint32)
    assert NODE.itemsize == 64
    k = lib.hostbvh_sah(C.c_uint32(n), ptr(c), ptr(v), ptr(r), C.c_int(leaf), C.c_double(node_cost), ptr(nodes), C.c_uint32(len(nodes)), ptr(order))
    assert k >= 1
    return nodes[:k], order[:n]


def check_sah_tree(nodes, order, lo, hi, leaf):
    """Every sphere in exactly one leaf; every child box (FP32) strictly contains the f64 boxes of the spheres below it;
    leaves hold 1..leaf spheres; the nodes form a tree rooted at 0 with parents stored before their children."""
    n = len(order)
    assert sorted(order.tolist()) == list(range(n))
    seen_nodes, seen_spheres = set(), np.zeros(n, dtype=np.int32)

    def walk(i):
        assert i not in seen_nodes
        seen_nodes.add(i)
        nd = nodes[i]
        below = []
        for c in range(2):
            ch, cnt = int(nd["child"][c]), int(nd["cnt"][c])
            if ch >= 0:
                assert cnt == 0 and ch > i
                mine = walk(ch)
            elif cnt == 0:
                mine = []                                   # an unused slot (a scene that is one leaf): an empty box
                assert nd["lox"][c] > nd["hix"][c]
            else:
                first = ~ch
                assert 1 <= cnt <= leaf and 0 <= first and first + cnt <= n
                mine = order[first:first + cnt].tolist()
                for s in mine:
                    seen_spheres[s] += 1
            if mine:
                blo = np.array([nd["lox"][c], nd["loy"][c], nd["loz"][c]], dtype=np.float64)
                bhi = np.array([nd["hix"][c], nd["hiy"][c], nd["hiz"][c]], dtype=np.float64)
                assert (blo < lo[mine].min(axis=0)).all() and (bhi > hi[mine].max(axis=0)).all()
                # ... and tightly: within two FP32 steps of the f64 bound
                assert (np.abs(blo - lo[mine].min(axis=0)) <= 3 * np.spacing(np.abs(blo).astype(np.float32)).astype(np.float64) + 3e-45).all()
                assert (np.abs(bhi - hi[mine].max(axis=0)) <= 3 * np.spacing(np.abs(bhi).astype(np.float32)).astype(np.float64) + 3e-45).all()
            below += mine
        return below

    import sys
    sys.setrecursionlimit(10000)
    assert sorted(walk(0)) == list(range(n))
    assert len(seen_nodes) == len(nodes) and (seen_spheres == 1).all()


@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_sah_tree_of_the_rtow_scene(lib, leaf):
    c, v, r = spheres(rayz_b200.random_bouncing(64, seed=42).pool.arrays())
    lo, hi = boxes(lib, c, v, r)
    nodes, order = sah(lib, c, v, r, leaf=leaf)
    assert len(nodes) <= len(r) - 1 or len(r) == 1
    check_sah_tree(nodes, order, lo, hi, leaf)
    again, order2 = sah(lib, c, v, r, leaf=leaf)
    assert nodes.tobytes() == again.tobytes() and (order == order2).all()      # deterministic


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 200])
def test_sah_tree_of_small_and_degenerate_sets(lib, n):
    rng = np.random.default_rng(n)
    c = rng.uniform(-20, 20, (n, 3))
    v = np.where(rng.random((n, 1)) < 0.5, rng.uniform(-1, 1, (n, 3)), 0.0)
    r = rng.uniform(0.05, 2.0, n)
    if n >= 5:
        c[1] = c[0]; c[2] = c[0]; v[1] = v[0] = v[2] = 0.0           # coincident centroids: the count split
        r[3] = 500.0                                                 # one huge sphere among small ones
    c, v, r = np.ascontiguousarray(c), np.ascontiguousarray(v), np.ascontiguousarray(r)
    lo, hi = boxes(lib, c, v, r)
    nodes, order = sah(lib, c, v, r)
    check_sah_tree(nodes, order, lo, hi, 4)
    same = np.ascontiguousarray(np.tile(c[:1], (n, 1)))                # every centre in one point
    z = np.zeros_like(same)
    nodes, order = sah(lib, same, z, np.ascontiguousarray(np.full(n, 0.5)))
    lo, hi = boxes(lib, same, z, np.full(n, 0.5))
    check_sah_tree(nodes, order, lo, hi, 4)


def test_sphere_box_is_the_union_over_the_shutter(lib):
    """Sphere.boundingBox (geom.zig:24-31): the boxes at time 0 and time 1, united."""
    rng = np.random.default_rng(7)
    c, v, r = rng.uniform(-5, 5, (50, 3)), rng.uniform(-2, 2, (50, 3)), rng.uniform(0.1, 1.5, 50)
    lo, hi = boxes(lib, c, v, r)
    assert (lo == np.minimum(c - r[:, None], c + v - r[:, None])).all()
    assert (hi == np.maximum(c + r[:, None], c + v + r[:, None])).all()


def test_reference_shaped_tree(lib):
    """BVH.build (hit.zig:130-161): enclose the range, leaf at <= 2 hittables, else stable-sort by the low corner on the
    longest axis and split at n / 2; parents before children, the order array a permutation."""
    c, v, r = spheres(rayz_b200.random_bouncing(64, seed=42).pool.arrays())
    n = len(r)
    lo, hi = boxes(lib, c, v, r)
    nodes = np.zeros(2 * n, dtype=REFNODE)
    order = np.zeros(n, dtype=np.uint32)
    assert REFNODE.itemsize == 64
    k = lib.hostbvh_ref(C.c_uint32(n), ptr(c), ptr(v), ptr(r), ptr(nodes), C.c_uint32(len(nodes)), ptr(order))
    nodes = nodes[:k]
    assert sorted(order.tolist()) == list(range(n))

    def walk(i, si, ei):
        nd = nodes[i]
        ids = order[si:ei]
        assert (nd["low"] == lo[ids].min(axis=0)).all() and (nd["high"] == hi[ids].max(axis=0)).all()      # exact f64 enclosure
        if ei - si <= 2:
            assert nd["left"] == -1 and nd["right"] == -1 and (nd["start"], nd["end"]) == (si, ei)
            return 1
        ext = nd["high"] - nd["low"]
        axis = (0 if ext[0] > ext[2] else 2) if ext[0] > ext[1] else (1 if ext[1] > ext[2] else 2)         # amax tie rule, vec.zig:150-156
        mid = (ei - si) // 2 + si
        assert lo[order[si:mid], axis].max() <= lo[order[mid:ei], axis].min()      # (the halves are re-sorted on their own axes below)
        assert nd["left"] == i + 1
        nl = walk(int(nd["left"]), si, mid)
        assert nd["right"] == i + 1 + nl
        return 1 + nl + walk(int(nd["right"]), mid, ei)

    assert walk(0, 0, n) == k


def test_rounding_helpers_match_nextafter(lib):
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2**32, 200000, dtype=np.uint64).astype(np.uint32)
    special = np.array([0x00000000, 0x80000000, 0x00000001, 0x80000001, 0x007fffff, 0x00800000, 0x7f7fffff, 0xff7fffff, 0x7f800000, 0xff800000,
                        0x3f800000, 0xbf800000], dtype=np.uint32)
    f = np.concatenate([bits, special]).view(np.float32)
    f = np.ascontiguousarray(f[~np.isnan(f)])
    down, up = np.empty_like(f), np.empty_like(f)
    lib.hostbvh_next(ptr(f), C.c_uint32(len(f)), ptr(down), ptr(up))
    with np.errstate(over="ignore"):   # FLT_MAX -> inf is the point
        want_d, want_u = np.nextafter(f, np.float32(-np.inf)), np.nextafter(f, np.float32(np.inf))
    assert (down.view(np.uint32) == want_d.view(np.uint32)).all() and (up.view(np.uint32) == want_u.view(np.uint32)).all()
    # SahBuilder::down / up: strictly outside the f64 value, by at most two FP32 steps
    d = np.ascontiguousarray(np.concatenate([rng.uniform(-1e3, 1e3, 100000), rng.normal(0, 1e-3, 1000), [0.0, 1.0, -1.0, 0.1, 1e30, -1e30]]))
    lo, hi = np.empty(len(d), dtype=np.float32), np.empty(len(d), dtype=np.float32)
    lib.hostbvh_round(ptr(d), C.c_uint32(len(d)), ptr(lo), ptr(hi))
    assert (lo.astype(np.float64) < d).all() and (hi.astype(np.float64) > d).all()
    nearest = d.astype(np.float32)
    assert (lo >= np.nextafter(np.nextafter(nearest, np.float32(-np.inf)), np.float32(-np.inf))).all()
    assert (hi <= np.nextafter(np.nextafter(nearest, np.float32(np.inf)), np.float32(np.inf))).all()
