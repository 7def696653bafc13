"""CPU test of the C++ stand-in host's PPM writer (host/rayz_host.hpp Image::writePPM), the step right after the hot path
(SURVEY 8f #2): its bytes must be the reference writer's — header "P3\\n{w} {h}\\n255\\n", then "{r} {g} {b}\\n" per pixel
(image.zig:29-41) — and it must keep up with the GPU: the 3840x2160 frame of BASELINE config 3 (~95 MB of text) is formatted
in well under 100 ms on one host core."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def ppm_check():
    exe = os.path.join(HERE, "_build", "ppm_check")
    src = os.path.join(HERE, "hostsim", "ppm_check.cpp")
    hdr = os.path.join(ROOT, "host", "rayz_host.hpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    return exe


def lcg_bytes(n: int, seed: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    s = seed
    for i in range(n):
        s = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = s >> 56
    return out


@pytest.mark.parametrize("w,h", [(1, 1), (7, 3), (160, 90)])
def test_writeppm_bytes_are_the_reference_format(ppm_check, tmp_path, w, h):
    out = tmp_path / "o.ppm"
    r = subprocess.run([ppm_check, str(w), str(h), "42", str(out)], capture_output=True, text=True, check=True)
    rgb = lcg_bytes(w * h * 3, 42).reshape(-1, 3)
    if w * h >= 256:
        assert len(set(rgb.ravel().tolist())) == 256          # every token of the table is exercised
    want = f"P3\n{w} {h}\n255\n" + "".join(f"{a} {b} {c}\n" for a, b, c in rgb.tolist())
    got = out.read_bytes()
    assert got == want.encode()
    assert int(r.stdout.split()[0]) == len(got)


def test_writeppm_throughput_at_config3_size(ppm_check, tmp_path):
    out = tmp_path / "big.ppm"
    r = subprocess.run([ppm_check, "3840", "2160", "7", str(out)], capture_output=True, text=True, check=True)
    n, ms = int(r.stdout.split()[0]), float(r.stdout.split()[1])
    print(f"formatPPM 3840x2160: {n} bytes in {ms:.1f} ms ({n / ms / 1e3:.0f} MB/s)")
    assert n == os.path.getsize(out) and 8_294_400 * 6 < n <= 8_294_400 * 12 + 32
    head = open(out, "rb").read(32)
    assert head.startswith(b"P3\n3840 2160\n255\n")
    assert ms < 100.0, f"{ms} ms to format the config-3 frame"
