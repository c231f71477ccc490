"""K2H (rawspeed_b200/csrc/hasselblad.cuh) without a GPU: the kernel bodies compiled by g++ against
tests/emu/cuda_emu.h and run in the plan's order (parse / link rounds, serial walk, scan, decode,
row sums), compared with the oracle's HasselbladDecompressor -- pixels, stream position, error class.
Parity of the real kernels is the GPU tests' job (tests/test_gpu_hasselblad.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

import rawspeed_b200 as rs
from rawspeed_b200 import _abi
from oracle import port, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "hasselblad_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libhasselblad_emu.so")
CSRC = os.path.join(HERE, "..", "rawspeed_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_emu.h")] + [
    os.path.join(CSRC, f) for f in ("hasselblad.cuh", "ljpeg_host.h", "ljpeg_types.h")]
NCPL, VALS = synth.DEFAULT_NCPL, synth.DEFAULT_VALUES


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.hass_emu_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int,
                                 C.POINTER(C.c_int)]
    return lib


def run_emu(lib, data, w, h, init_pred, max_rounds=99, ncpl=NCPL, vals=VALS):
    data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
    tab = rs.huff_table(bytes(ncpl), bytes(vals), False)
    out = port.new_image(w, h)
    status, consumed, rounds = C.c_uint32(9), C.c_uint32(0), C.c_int(0)
    rc = lib.hass_emu_run(data.ctypes.data, data.size, C.byref(tab), w, h, out.shape[1] * 2, init_pred,
                          out.ctypes.data, C.byref(status), C.byref(consumed), max_rounds, C.byref(rounds))
    assert rc == 0
    return status.value, consumed.value, out, rounds.value


def oracle_outcome(data, w, h, init_pred, ncpl=NCPL, vals=VALS):
    ht = port.Huff(ncpl, vals, full=False)
    img = port.new_image(w, h)
    try:
        c = port.hasselblad_decompress(img, w, ht, init_pred, bytes(data))
        return 0, c, img
    except port.IOException:
        return 2, None, img
    except port.RawDecoderException:
        return 1, None, img


def check(lib, data, w, h, init_pred, **kw):
    want = oracle_outcome(data, w, h, init_pred)
    status, consumed, out, rounds = run_emu(lib, data, w, h, init_pred, **kw)
    assert status == want[0], (status, want[0])
    if want[0] == 0:
        assert consumed == want[1]
        assert np.array_equal(out, want[2])
    return rounds


@pytest.mark.parametrize("w,h,wild", [(2, 1, False), (66, 9, False), (130, 21, True), (512, 40, False),
                                      (1024, 64, True)])
def test_round_trip(emu, w, h, wild):
    img = synth.image_model(w, h, seed=w, wild=wild, bits=16 if wild else 14)
    if wild:
        img[0, 0:4] = [0x8000, 0x8000, 0, 0xFFFF]    # differences of -32768 and wrap-around
    ht = port.Huff(NCPL, VALS, full=False)
    data = synth.make_hasselblad(img, ht, 0x8000)
    rounds = check(emu, data, w, h, 0x8000)
    assert rounds <= 4
    status, consumed, out, _ = run_emu(emu, data, w, h, 0x8000)
    assert np.array_equal(out[:, :w], img)


def test_the_serial_walk_finishes_what_the_rounds_did_not(emu):
    """With no parallel round at all every start is still a guess: the serial walk alone must
    produce the sequential parse."""
    img = synth.image_model(256, 48, seed=3)
    ht = port.Huff(NCPL, VALS, full=False)
    data = synth.make_hasselblad(img, ht, 0x2000)
    check(emu, data, 256, 48, 0x2000, max_rounds=0)
    check(emu, data, 256, 48, 0x2000, max_rounds=1)


def test_random_payloads(emu):
    """Random bytes: bad codes, or an image of noise -- same outcome as the oracle."""
    for seed in range(6):
        data = synth.lcg_bytes(4096, 9 + seed)
        check(emu, data, 64, 12, 0x2000)
        check(emu, data, 256, 40, 0x2000)     # needs more bits than there are


@pytest.mark.parametrize("cut", [0, 1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15, 16, 17, 20, 24, 28, 33, 64, 200])
def test_truncated_streams(emu, cut):
    """The buffer ends early: zero bits behind the data, IOException exactly when the reference's
    replenisher gets more than 8 bytes behind the buffer."""
    img = synth.image_model(192, 16, seed=5)
    ht = port.Huff(NCPL, VALS, full=False)
    data = bytes(synth.make_hasselblad(img, ht, 0x8000))
    # (the encoder appends slack; find the last byte that matters and cut from there)
    want0 = oracle_outcome(data, 192, 16, 0x8000)
    assert want0[0] == 0
    end = want0[1]
    for base in (end + 16, end + 4, end, end - 1):
        n = base - cut
        if n <= 0:
            continue
        check(emu, data[:n], 192, 16, 0x8000)


@pytest.mark.parametrize("seed", range(10))
def test_random_tables(emu, seed):
    """Random COMPLETE canonical codes over the difference lengths 0 .. 16 (codes of up to 16 bits, values in
    random order; 17 values: "Hasselblad uses 17"), noise from a few bits to the full range: the speculative
    segment parse has to settle on the sequential one whatever the code looks like."""
    from test_ljpeg_stream_emu import _random_table
    rng = np.random.default_rng(4000 + seed)
    w, h = int(rng.choice([64, 130, 512])), int(rng.integers(3, 24))
    bits = int(rng.choice([4, 9, 16]))
    img = ((0x8000 + rng.integers(0, 1 << bits, (h, w)) - (1 << bits) // 2) & 0xFFFF).astype(np.uint16)
    t = _random_table(rng, 17)
    ncpl, vals = bytes(t.ncpl), bytes(t.values)
    ht = port.Huff(ncpl, vals, full=False)
    data = synth.make_hasselblad(img, ht, 0x8000)
    want = oracle_outcome(data, w, h, 0x8000, ncpl=ncpl, vals=vals)
    assert want[0] == 0 and np.array_equal(want[2][:, :w], img)
    status, consumed, out, rounds = run_emu(emu, data, w, h, 0x8000, ncpl=ncpl, vals=vals)
    assert status == 0 and consumed == want[1] and np.array_equal(out, want[2])
