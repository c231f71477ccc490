"""PanasonicV4 on the device, without a GPU: pana_kernel<4,12>'s packet arithmetic and
section-swap addressing (rawspeed_b200/csrc/pana4_core.h) compiled as plain C++ and the
kernel's per-thread program replayed on the CPU (tests/emu/pana4_emu.cpp), compared with the
oracle (pinned against the compiled reference in tests/test_oracle_panasonic.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "pana4_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libpana4_emu.so")
DEPS = [SRC, os.path.join(HERE, "..", "rawspeed_b200", "csrc", "pana4_core.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.pana4_emu_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                                  C.c_uint32, C.POINTER(C.c_uint32)]
    return lib


def v4_payload(w, h, split, seed, zero_every=7):
    nbytes = w * h // 14 * 16
    if split:
        nbytes = (nbytes + 0x3FFF) // 0x4000 * 0x4000
    data = synth.lcg_bytes(nbytes, seed=seed)
    if zero_every:
        data[::zero_every] = 0   # zero steps: the zero-reference and bad-pixel branches
    return data


@pytest.mark.parametrize("w,h,split,zero_ok", [
    (14, 1, 0, True), (28, 3, 0, False), (1400, 25, 0x1FF8, True), (2800, 13, 0x1FF8, False),
    (1414, 9, 0, False), (4200, 6, 0x2008, True), (4200, 6, 0x2008, False),
    (1428, 11, 0x4000, False),          # split == BlockSize: no swap
    (1428, 30, 0x1235, False),          # split not a multiple of 8: byte-wise gather
    (1428, 30, 3, True), (5600, 12, 0x3FFF, False),
])
def test_replayed_v4_matches_oracle(emu, w, h, split, zero_ok):
    data = v4_payload(w, h, split, w + h)
    want = port.new_image(w, h)
    zwant = port.panasonic_v4(want, w, data, zero_ok, split, cap=1 << 20)
    got = port.new_image(w, h)
    zl = np.zeros(1 << 20, dtype=np.uint32)
    nz = C.c_uint32(0)
    pad = np.concatenate([np.zeros(5, np.uint8), data, np.zeros(16, np.uint8)])   # odd input offset
    emu.pana4_emu_run(pad.ctypes.data, 5, got.ctypes.data, 0, got.shape[1] * 2, w, h, split,
                      int(zero_ok), zl.ctypes.data, zl.size, C.byref(nz))
    assert np.array_equal(got, want)
    assert sorted(zl[:nz.value].tolist()) == zwant
    assert zero_ok is False or nz.value == 0


def test_all_zero_and_all_ones_packets(emu):
    # an all-zero packet: every byte 0 -> the 4-bit fields are read at pixels 12 / 13; all pixels 0
    for fill in (0x00, 0xFF, 0x80, 0x01):
        w, h = 140, 4
        data = np.full(w * h // 14 * 16, fill, dtype=np.uint8)
        want = port.new_image(w, h)
        zwant = port.panasonic_v4(want, w, data, False, 0)
        got = port.new_image(w, h)
        zl = np.zeros(4096, dtype=np.uint32)
        nz = C.c_uint32(0)
        buf = np.concatenate([data, np.zeros(16, np.uint8)])
        emu.pana4_emu_run(buf.ctypes.data, 0, got.ctypes.data, 0, got.shape[1] * 2, w, h, 0, 0,
                          zl.ctypes.data, zl.size, C.byref(nz))
        assert np.array_equal(got, want)
        assert sorted(zl[:nz.value].tolist()) == zwant
        if fill == 0:
            assert nz.value == w * h
