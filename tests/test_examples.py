"""examples/pipeline.c -- a plain-C consumer of the C ABI (unpack -> lookup -> scale -> bad
pixels with host buffers): compiles and links against the shipped library without a GPU
(every symbol it needs is exported, the header is valid C99); on a GPU box it runs, and its
result is compared with the oracle's for the same steps."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "pipeline.c")


def _build(tmp_path):
    exe = str(tmp_path / "pipeline")
    libdir = os.path.join(ROOT, "rawspeed_b200")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           SRC, "-L", libdir, "-l:librawspeed_b200.so", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_example_compiles_and_links_as_c99(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the gpu-marked test runs it")
    r = subprocess.run([exe], capture_output=True, text=True)
    # no GPU here: the product fails loudly at rsb200_create, it does not fall back
    assert r.returncode != 0 and "rsb200_create" in r.stderr


def _expected():
    from oracle import port
    import rawspeed_b200 as rs
    W, H, BPS = 4000, 300, 12
    s = np.uint32(1)
    n = W * BPS // 8 * H
    packed = np.empty(n, dtype=np.uint8)
    # s = s * 1664525 + 1013904223, byte = s >> 24 (vectorised: closed form of the LCG)
    a, c = 1664525, 1013904223
    x = 1
    out = bytearray(n)
    for i in range(n):
        x = (x * a + c) & 0xFFFFFFFF
        out[i] = x >> 24
    packed[:] = np.frombuffer(bytes(out), dtype=np.uint8)
    img = port.new_image(W, H, fill=0)
    port.unpack(packed, img, W, 1, (0, 0, W, H), W * BPS // 8, BPS, rs.MSB)
    table = np.where(np.arange(65536) < 4096, np.arange(65536) * 15, 61425).astype(np.uint16)
    port.sixteen_bit_lookup(img, W, 1, table, False)
    port.scale_values(img, W, (8, 2, W - 16, H - 4), (960,) * 4, 61425, dither=True)
    port.fix_bad_pixels(img, W, 1, np.array([(10 << 16) | 100, (200 << 16) | 3999], dtype=np.uint32), True)
    return img[:, :W]


@pytest.mark.gpu
def test_example_runs_and_matches_the_oracle(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"pixel sum (\d+), first pixels (\d+) (\d+) (\d+) (\d+), kernels launched (\d+)", r.stdout)
    assert m, r.stdout
    want = _expected()
    assert int(m.group(1)) == int(want.astype(np.uint64).sum())
    assert [int(m.group(k)) for k in range(2, 6)] == want[0, :4].tolist()
    assert int(m.group(6)) == 4
