"""K12 (whole-image table lookup) without a GPU: the kernel's per-lane functions
(lookup_core.h) and job builder (lookup_host.h) compiled as plain C++ and the kernel's warp
loop replayed on the CPU (tests/emu/lookup_emu.cpp), against the oracle (pinned against the
compiled reference in tests/test_oracle_lookup.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port
from rawspeed_b200._abi import LookupJob
from test_oracle_lookup import CASES, image, curve

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "lookup_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "liblookup_emu.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "rawspeed_b200", "csrc", f)
                for f in ("lookup_core.h", "lookup_host.h", "scale_core.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.lookup_emu_run.argtypes = [C.c_void_p, C.POINTER(LookupJob), C.c_int, C.c_void_p, C.c_int, C.c_int,
                                   C.c_char_p, C.c_int]
    lib.lookup_emu_run_smem.argtypes = [C.c_void_p, C.POINTER(LookupJob), C.c_int, C.c_void_p, C.c_int,
                                        C.c_char_p, C.c_int]
    lib.lookup_emu_mwc_direct.argtypes = lib.lookup_emu_mwc_state.argtypes = [C.c_uint32] * 3
    lib.lookup_emu_mwc_direct.restype = lib.lookup_emu_mwc_state.restype = C.c_uint32
    return lib


def job(offset, img, w, cpp, table=0):
    j = LookupJob()
    j.offset, j.pitch, j.width, j.height, j.cpp, j.table = offset, img.shape[1] * 2, w, img.shape[0], cpp, table
    return j


@pytest.mark.parametrize("dither", [False, True])
@pytest.mark.parametrize("k", range(len(CASES)))
def test_replayed_kernel_matches_oracle(emu, k, dither):
    w, h, cpp, crop, ncurve = CASES[k]
    a = image(w, h, cpp, k)
    want = a.copy()
    t = port.build_table(curve(ncurve, 10 + k), dither)
    port.sixteen_bit_lookup(want, w, cpp, t, dither)
    err = C.create_string_buffer(256)
    assert emu.lookup_emu_run(a.ctypes.data, (LookupJob * 1)(job(0, a, w, cpp)), 1, t.ctypes.data, 1,
                              int(dither), err, 256) == 0, err.value
    assert np.array_equal(a, want)


@pytest.mark.parametrize("dither", [False, True])
def test_two_images_two_tables_wide_rows(emu, dither):
    specs = [(2100, 7, 1), (300, 9, 3)]        # 2100 samples = 263 groups: nine iterations per lane
    imgs = [image(w, h, cpp, 30 + i) for i, (w, h, cpp) in enumerate(specs)]
    tabs = np.stack([port.build_table(curve(4096, 1), dither), port.build_table(curve(700, 2), dither)])
    sizes = [(im.nbytes + 255) // 256 * 256 for im in imgs]
    buf = np.zeros(sum(sizes), dtype=np.uint8)
    jobs, o = [], 0
    for i, (im, sz, (w, h, cpp)) in enumerate(zip(imgs, sizes, specs)):
        buf[o:o + im.nbytes] = im.reshape(-1).view(np.uint8)
        jobs.append(job(o, im, w, cpp, i))
        o += sz
    err = C.create_string_buffer(256)
    assert emu.lookup_emu_run(buf.ctypes.data, (LookupJob * 2)(*jobs), 2, tabs.ctypes.data, 2, int(dither),
                              err, 256) == 0, err.value
    o = 0
    for i, (im, sz, (w, h, cpp)) in enumerate(zip(imgs, sizes, specs)):
        want = im.copy()
        port.sixteen_bit_lookup(want, w, cpp, tabs[i], dither)
        assert np.array_equal(buf[o:o + im.nbytes].view(np.uint16).reshape(im.shape), want)
        o += sz


def test_jump_ahead_equals_stepping(emu):
    # every seed lies above the modulus; rows whose seed has a low half of 0xFFFF stay above for
    # a second step ((w + 13 y) & 0xFFFF == 0xBA7B)
    for w, y in [(8256, 0), (8256, 5503), (40, 3669), (47739, 0), (0xBA7B - 13 * 7, 7), (65535, 65535)]:
        for x in [0, 1, 2, 7, 8, 248, 249, 1000, 24767]:
            assert emu.lookup_emu_mwc_state(w, y, x) == emu.lookup_emu_mwc_direct(w, y, x)


@pytest.mark.parametrize("grid", [1, 3, 148])
def test_persistent_grid_variant_covers_every_quad_once(emu, grid):
    """The RSB200_LUT_SMEM=1 candidate (lookup_smem_kernel): grid-stride walk of the quads."""
    specs = [(2100, 7, 1), (300, 9, 3), (64, 130, 1)]
    imgs = [image(w, h, cpp, 60 + i) for i, (w, h, cpp) in enumerate(specs)]
    t = port.build_table(curve(4096, 3), False)
    sizes = [(im.nbytes + 255) // 256 * 256 for im in imgs]
    buf = np.zeros(sum(sizes), dtype=np.uint8)
    jobs, o = [], 0
    for im, sz, (w, h, cpp) in zip(imgs, sizes, specs):
        buf[o:o + im.nbytes] = im.reshape(-1).view(np.uint8)
        jobs.append(job(o, im, w, cpp, 0))
        o += sz
    err = C.create_string_buffer(256)
    assert emu.lookup_emu_run_smem(buf.ctypes.data, (LookupJob * 3)(*jobs), 3, t.ctypes.data, grid, err, 256) == 0
    o = 0
    for im, sz, (w, h, cpp) in zip(imgs, sizes, specs):
        want = im.copy()
        port.sixteen_bit_lookup(want, w, cpp, t, False)
        assert np.array_equal(buf[o:o + im.nbytes].view(np.uint16).reshape(im.shape), want)
        o += sz
