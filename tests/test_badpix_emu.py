"""K11 (bad-pixel interpolation) without a GPU: the bitmap / list builder (badpix_host.h) and
the per-pixel interpolation (badpix_core.h) compiled as plain C++ and the kernel's grid
replayed on the CPU (tests/emu/badpix_emu.cpp), against the oracle (pinned against the compiled
reference in tests/test_oracle_badpixels.py); plus the host mirror's transferBadPixelsToMap."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port
from rawspeed_b200 import host
from rawspeed_b200._abi import BadPixJob
from test_oracle_badpixels import scenarios, image, pos

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "badpix_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libbadpix_emu.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "rawspeed_b200", "csrc", f)
                for f in ("badpix_core.h", "badpix_host.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.badpix_emu_run.argtypes = [C.c_void_p, C.POINTER(BadPixJob), C.c_int, C.c_void_p, C.c_uint32,
                                   C.c_int, C.c_char_p, C.c_int]
    return lib


def job(offset, img, w, cfa, first, n, prior=None):
    j = BadPixJob()
    j.offset, j.pitch, j.width, j.height = offset, img.shape[1] * 2, w, img.shape[0]
    j.is_cfa, j.first_position, j.num_positions = int(cfa), first, n
    j.prior_map = prior.ctypes.data if prior is not None else None
    return j


CPP1 = [k for k, s in enumerate(scenarios()) if s[3] == 1]


@pytest.mark.parametrize("two_phase", [0, 1])
@pytest.mark.parametrize("k", CPP1)
def test_replayed_kernel_matches_oracle(emu, k, two_phase):
    name, w, h, cpp, cfa, points = scenarios()[k]
    a = image(w, h, 1, k)
    want = a.copy()
    port.fix_bad_pixels(want, w, 1, pos(points), cfa)
    p = pos(points)
    err = C.create_string_buffer(256)
    n = emu.badpix_emu_run(a.ctypes.data, (BadPixJob * 1)(job(0, a, w, cfa, 0, p.size)), 1,
                           p.ctypes.data, p.size, two_phase, err, 256)
    assert n >= 0, err.value
    assert np.array_equal(a, want)


def test_two_images_and_a_prior_map(emu):
    (_, w1, h1, _, cfa1, pts1), (_, w2, h2, _, cfa2, pts2) = scenarios()[0], scenarios()[3]
    a1, a2 = image(w1, h1, 1, 20), image(w2, h2, 1, 21)
    # image 2 already has a map holding half of its bad pixels
    prior = host.fix_bad_pixels(a2.copy(), w2, 1, pos(pts2[::2]), cfa2, map_only=True)
    want1, want2 = a1.copy(), a2.copy()
    port.fix_bad_pixels(want1, w1, 1, pos(pts1), cfa1)
    port.fix_bad_pixels(want2, w2, 1, pos(pts2), cfa2)
    sz1 = (a1.nbytes + 255) // 256 * 256
    buf = np.zeros(sz1 + a2.nbytes, dtype=np.uint8)
    buf[:a1.nbytes] = a1.reshape(-1).view(np.uint8)
    buf[sz1:] = a2.reshape(-1).view(np.uint8)
    p = np.concatenate([pos(pts1), pos(pts2[1::2])])
    jobs = (BadPixJob * 2)(job(0, a1, w1, cfa1, 0, len(pts1)),
                           job(sz1, a2, w2, cfa2, len(pts1), len(pts2[1::2]), prior))
    err = C.create_string_buffer(256)
    assert emu.badpix_emu_run(buf.ctypes.data, jobs, 2, p.ctypes.data, p.size, 1, err, 256) >= 0, err.value
    assert np.array_equal(buf[:a1.nbytes].view(np.uint16).reshape(a1.shape), want1)
    assert np.array_equal(buf[sz1:].view(np.uint16).reshape(a2.shape), want2)


def test_mirror_bitmap_is_the_reference_layout():
    w, h = 70, 5
    m = host.fix_bad_pixels(image(w, h, 1, 1), w, 1, pos([(0, 0), (0, 9), (4, 69), (2, 33), (2, 33)]), map_only=True)
    assert m.shape == (h, 16)
    want = np.zeros_like(m)
    for y, x in [(0, 0), (0, 9), (4, 69), (2, 33)]:
        want[y, x >> 3] |= 1 << (x & 7)
    assert np.array_equal(m, want)


def test_position_outside_the_image_is_refused(emu):
    a = image(32, 4, 1, 1)
    p = pos([(4, 0)])
    err = C.create_string_buffer(256)
    assert emu.badpix_emu_run(a.ctypes.data, (BadPixJob * 1)(job(0, a, 32, True, 0, 1)), 1, p.ctypes.data,
                              1, 1, err, 256) == -1
    with pytest.raises(Exception):
        host.fix_bad_pixels(a, 32, 1, p, map_only=True)
