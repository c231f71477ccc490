"""K9 (black/white scaling) without a GPU: the kernel's per-lane functions (scale_core.h)
and job builder (scale_host.h) are compiled as plain C++ (tests/emu/scale_emu.cpp) and the
kernel's warp loop is replayed on the CPU, then compared with the oracle.  This checks the
arithmetic, the indexing and the jump-ahead of the dither generators; the parity of the real
kernel is the GPU test's job (tests/test_gpu_scale.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port
from rawspeed_b200._abi import ScaleJob, SCALE_AUTO, SCALE_PLAIN, SCALE_SSE2

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "scale_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libscale_emu.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "rawspeed_b200", "csrc", f)
                for f in ("scale_core.h", "scale_host.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.scale_emu_run.argtypes = [C.c_void_p, C.POINTER(ScaleJob), C.c_int, C.c_char_p, C.c_int]
    lib.scale_emu_mwc_direct.argtypes = lib.scale_emu_mwc_state.argtypes = [C.c_uint32] * 3
    lib.scale_emu_mwc_direct.restype = lib.scale_emu_mwc_state.restype = C.c_uint32
    return lib


def job(offset, img, w, h, cpp, crop, black, white, dither=True, path=SCALE_AUTO):
    j = ScaleJob()
    j.offset, j.pitch, j.width, j.height, j.cpp = offset, img.shape[1] * 2, w, h, cpp
    j.crop_x, j.crop_y, j.crop_w, j.crop_h = crop
    for i in range(4):
        j.black_separate[i] = black[i]
    j.white_point, j.dither, j.path = white, int(dither), path
    return j


def run(lib, buf, jobs):
    arr = (ScaleJob * len(jobs))(*jobs)
    err = C.create_string_buffer(256)
    n = lib.scale_emu_run(buf.ctypes.data, arr, len(jobs), err, 256)
    assert n >= 0, err.value
    return n


def image(w, h, cpp, seed, lo=0, hi=65536):
    rng = np.random.default_rng(seed)
    a = port.new_image(w, h, cpp)
    a[:, :] = rng.integers(lo, hi, size=a.shape, dtype=np.uint16)
    return a


CASES = [
    # w, h, cpp, crop, black, white
    (64, 16, 1, (0, 0, 64, 16), (256, 256, 256, 256), 16383),
    (70, 11, 1, (3, 1, 61, 9), (60, 64, 68, 72), 4095),            # odd offsets, rows % 4 != 0
    (37, 9, 1, (2, 3, 30, 5), (1000, 1010, 990, 1024), 15000),     # width % 8 != 0
    (300, 7, 1, (5, 1, 290, 5), (64, 65, 66, 67), 16000),          # more than 32 groups per row
    (1000, 6, 1, (8, 0, 980, 6), (512, 512, 512, 512), 16383),     # several iterations
    (40, 6, 1, (0, 1, 40, 4), (2048, 2000, 2100, 2047), 3000),     # app_scale ~ 68: plain loop
    (33, 7, 1, (5, 2, 20, 4), (100, 200, 300, 400), 1023),         # plain, odd crop_x, skip = 5
    (530, 9, 1, (11, 2, 515, 6), (10, 20, 30, 40), 900),           # plain, 66 groups
    (40, 10, 3, (2, 1, 30, 8), (100, 100, 100, 100), 15000),       # cpp 3, SSE2 loop
    (40, 10, 3, (3, 1, 30, 8), (100, 100, 100, 100), 900),         # cpp 3, plain loop
]


@pytest.mark.parametrize("path", [SCALE_AUTO, SCALE_SSE2, SCALE_PLAIN])
@pytest.mark.parametrize("dither", [True, False])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_replayed_kernel_matches_oracle(emu, case, dither, path):
    w, h, cpp, crop, black, white = CASES[case]
    a = image(w, h, cpp, 40 + case)
    b = a.copy()
    sse2 = None if path == SCALE_AUTO else path == SCALE_SSE2
    im = port._img(b, w, cpp)  # noqa: F841  (shape check)
    want = b
    # oracle (cpp-aware through scale_black_white with black_sep given)
    port.scale_black_white(want, w, crop, black_sep=list(black), white=white, dither=dither,
                           sse2=sse2, cpp=cpp, is_cfa=cpp == 1)
    assert run(emu, a, [job(0, a, w, h, cpp, crop, black, white, dither, path)]) == 1
    assert np.array_equal(a, want)


def test_two_images_two_loops_in_one_plan(emu):
    # image 0 -> SSE2 loop, image 1 -> plain loop, image 2 -> SSE2 loop again; one buffer
    specs = [(64, 13, (1, 1, 60, 11), (256,) * 4, 16383), (48, 9, (3, 0, 40, 9), (64,) * 4, 1023),
             (96, 6, (0, 2, 96, 3), (10, 20, 30, 40), 4095)]
    imgs = [image(w, h, 1, 7 + i) for i, (w, h, *_) in enumerate(specs)]
    sizes = [(im.nbytes + 255) // 256 * 256 for im in imgs]
    buf = np.zeros(sum(sizes), dtype=np.uint8)
    offs, jobs = [], []
    o = 0
    for im, sz, (w, h, crop, black, white) in zip(imgs, sizes, specs):
        buf[o:o + im.nbytes] = im.reshape(-1).view(np.uint8)
        jobs.append(job(o, im, w, h, 1, crop, black, white))
        offs.append(o)
        o += sz
    assert run(emu, buf, jobs) == 2
    for im, o, (w, h, crop, black, white) in zip(imgs, offs, specs):
        want = im.copy()
        port.scale_values(want, w, crop, black, white)
        got = buf[o:o + im.nbytes].view(np.uint16).reshape(im.shape)
        assert np.array_equal(got, want)


def test_jump_ahead_equals_stepping(emu):
    # v' = 18000 v mod (18000 * 2^16 - 1): states reached by jumping == states reached by stepping,
    # including seeds above the modulus (crop rows beyond ~31900) and the seed whose low half is 65535
    for crop_w, y in [(6000, 0), (8256, 5503), (65535, 65534), (100, 40000), (65535 - 36969 % 65536, 1),
                      (0xFFFF, 0), (0xFFFF, 32768)]:
        for x in [0, 1, 2, 7, 8, 247, 248, 1000, 8255, 60000]:
            assert emu.scale_emu_mwc_state(crop_w, y, x) == emu.scale_emu_mwc_direct(crop_w, y, x)


def test_descriptor_checks(emu):
    a = image(64, 8, 1, 1)
    err = C.create_string_buffer(256)
    good = job(0, a, 64, 8, 1, (0, 0, 64, 8), (256,) * 4, 16383)
    bad = []
    for field, value in [("offset", 8), ("pitch", 120), ("crop_w", 65), ("crop_h", 0), ("cpp", 5),
                         ("white_point", 256), ("path", 3), ("width", 0)]:
        j = ScaleJob.from_buffer_copy(good)
        setattr(j, field, value)
        bad.append(j)
    for j in bad:
        assert emu.scale_emu_run(a.ctypes.data, (ScaleJob * 1)(j), 1, err, 256) == -1
        assert err.value
