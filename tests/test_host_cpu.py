"""CPU tests of the host-side logic and of the C-ABI library surface (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rayz_b200
from rayz_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = abi.load()
    hdr = open(os.path.join(ROOT, "include", "rayz_cuda.h")).read()
    declared = set(re.findall(r"\b(rayz_cuda_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(abi.SYMBOLS), (declared ^ set(abi.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rayz_cuda_abi_version() == 3


def test_struct_layouts_match_header():
    assert C.sizeof(abi.RzCamera) == 18 * 8 + 8
    assert C.sizeof(abi.RzRenderParams) == 56
    assert C.sizeof(abi.RzScene) == 16 + 14 * 8
    assert C.sizeof(abi.RzStats) == 80
    assert C.sizeof(abi.RzTiming) == 56
    assert C.sizeof(abi.RzConfig) == 40
    assert C.sizeof(abi.RzTuning) == 96


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(abi.BackendError) as e:
        rayz_b200.Backend((0,))
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_shard_rows_partition_the_image():
    lib = abi.load()
    for h in (1, 3, 4, 7, 225, 675, 2160):
        for count in (1, 2, 3, 4, 8):
            for band in (1, 4, 8):
                rows = [lib.rayz_cuda_shard_rows(h, s, count, band) for s in range(count)]
                assert sum(rows) == h
                # python restatement of the banding rule
                want = [0] * count
                for j in range(h):
                    want[(j // band) % count] += 1
                assert rows == want
    assert lib.rayz_cuda_shard_rows(100, 0, 0, 0) == 100


def test_xoshiro_matches_oracle_restatement(orc):
    r = rayz_b200.Xoshiro256(42)
    want = np.zeros(64)
    orc.lib().orc_rng_f64(42, 64, want.ctypes.data)
    got = np.array([r.float() for _ in range(64)])
    assert np.array_equal(got, want)


def test_camera_init_matches_oracle_bitwise(orc):
    for w in (400, 1200, 3840):
        h = int(w / rayz_b200.ASPECT_RATIO)
        mine = rayz_b200.Camera.init(20.0, 10.0, 0.6, (13, 2, 3), (0, 0, 0), (0, 1, 0), h, w).rz
        ref, h2 = orc.default_camera(w)
        assert h == h2
        for f in ("look_from", "px_du", "px_dv", "px_origin", "defocus_u", "defocus_v"):
            assert list(getattr(mine, f)) == list(getattr(ref, f)), f
        assert mine.defocus == ref.defocus == 1


def test_camera_golden_rays_from_reference_test():
    """renderer.zig:129-149 "get ray" through the host mirror (pixel-centre ray = px_du*x + px_dv*y + px_origin - look_from)."""
    cam = rayz_b200.Camera.init(90.0, 12 ** 0.5, 0.0, (-2, 2, 1), (0, 0, -1), (0, 1, 0), 225, 400).rz
    def ray(px, py):
        return [cam.px_du[i] * px + cam.px_dv[i] * py + cam.px_origin[i] - cam.look_from[i] for i in range(3)]
    for got, want in ((ray(0, 0), (-0.935834, 0.815856, -7.75169)), (ray(112, 199), (-0.998817, -4.18732, -2.8115))):
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-5 * abs(w)


@pytest.mark.parametrize("kw", [dict(), dict(glass_heavy=True), dict(seed=7), dict(grid_lo=-3, grid_hi=4)])
def test_random_bouncing_matches_oracle_scene_bytes(orc, kw):
    t = rayz_b200.random_bouncing(400, **({"seed": 42} | kw))
    mine = t.pool.arrays()
    ref = orc.Scene.random_bouncing(kw.get("seed", 42), kw.get("grid_lo", -11), kw.get("grid_hi", 11),
                                    kw.get("glass_heavy", False)).arrays()
    assert set(mine) == set(ref)
    for k in ref:
        assert mine[k].shape == ref[k].shape, k
        assert np.array_equal(mine[k], ref[k]), k


def test_tracer_defaults_match_reference():
    t = rayz_b200.Tracer(400, 20.0, 10.0, 0.6, (13, 2, 3), (0, 0, 0), (0, 1, 0))
    assert (t.max_bounces, t.samples_per_px) == (50, 10)      # renderer.zig:23-24
    assert (t.img.w, t.img.h) == (400, 225)                   # renderer.zig:39-40
    assert t.img.pixels.shape == (400 * 225, 3)
    t2 = rayz_b200.Tracer(1200, 20.0, 10.0, 0.6, (13, 2, 3), (0, 0, 0), (0, 1, 0))
    assert t2.img.h == 675


def test_upload_scene_validation_without_gpu():
    """Argument errors are reported before any CUDA call is needed (NULL context)."""
    lib = abi.load()
    assert lib.rayz_cuda_upload_scene(None, None) == -1
    assert b"NULL" in lib.rayz_cuda_last_error()
    assert lib.rayz_cuda_render(None, None, None, None, None, None) == -1
    assert lib.rayz_cuda_primary_ids(None, None, 1, 1, 1, None) == -1


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py's contract is ONE JSON line on stdout; libraries that write to fd 1 (NCCL's version banner) must not leak
    into it.  The reference arm runs on the CPU, so the contract can be checked here on a tiny sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, runpy, sys; os.write(1, b''); sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', "
            "'--ref-spp', '1']; import bench; bench._claim_stdout(); os.write(1, b'library noise on fd 1\\n'); bench.main()")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "library noise" in r.stderr
