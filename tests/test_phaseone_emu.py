"""Third version of K8 (rawspeed_b200/csrc/phaseone.cuh: group headers walked per row, pixels decoded
in parallel with a segmented warp scan) without a GPU: the kernel bodies compiled by g++ against
tests/emu/cuda_emu.h, compared with the oracle's PhaseOneDecompressor (pinned against the compiled
reference in tests/test_oracle_phaseone.py) -- pixels of the whole padded buffer and the failure flag.
Parity of the real kernels is the GPU tests' job (tests/test_gpu_phaseone.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "phaseone_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libphaseone_emu.so")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_emu.h"),
        os.path.join(HERE, "..", "rawspeed_b200", "csrc", "phaseone.cuh")]


FORM = {"default": 0, "first": 1, "fast_no_touch": 2, "prefetch": 3, "fast_touch": 5, "blocks": 6, "cadence": 8}


@pytest.fixture(scope="module", params=["default", "first", "prefetch", "fast_touch", "blocks", "cadence"])
def emu(request):
    """Forms of the header walk: the default (128-byte lines through a per-row ring in shared memory, table-driven
    length codes), the first one (generic chunk loads, branches), aligned-word windows with a prefetch or with
    look-ahead loads, and 16-byte blocks cached in registers."""
    lib = _load()
    lib.form = FORM[request.param]
    return lib


def _load():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.p1_emu_run.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint32), C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.POINTER(C.c_uint32), C.c_int, C.c_int]
    return lib


def run_emu(lib, blob, strips, w, h, reverse=False):
    blob = np.ascontiguousarray(blob)
    n = len(strips)
    offs = (C.c_uint64 * n)(*[s[0] for s in strips])
    sizes = (C.c_uint32 * n)(*[s[1] for s in strips])
    rows = (C.c_uint32 * n)(*[s[2] for s in strips])
    out = port.new_image(w, h)
    bad = C.c_uint32(9)
    rc = lib.p1_emu_run(blob.ctypes.data, blob.size, offs, sizes, rows, n, w, out.shape[1] * 2,
                        out.ctypes.data, C.byref(bad), int(reverse), getattr(lib, "form", 0))
    assert rc == 0
    return out, bad.value


def oracle_outcome(blob, strips, w, h):
    want = port.new_image(w, h)
    try:
        port.phaseone(want, w, blob, strips)
        return want, 0
    except port.RawDecoderException:
        return want, 1


@pytest.mark.parametrize("w,h,wild", [(2, 1, False), (6, 3, False), (8, 1, False), (10, 2, True), (70, 9, False),
                                      (258, 33, True), (264, 5, False), (1000, 12, False), (2050, 6, True)])
def test_encoded_images(emu, w, h, wild):
    """Widths below one group, one group exactly, tails of 2 / 4 / 6 raw pixels, more than 32 groups (the
    carry from one warp step to the next), wild 16-bit noise (many raw groups: restarts of the scan)."""
    img = synth.image_model(w, h, seed=w, wild=wild, bits=16 if wild else 14)
    blob, strips = synth.make_phaseone(img, shuffle_seed=h, gap=3)
    want, bad = oracle_outcome(blob, strips, w, h)
    assert bad == 0
    for rev in (False, True):
        got, gbad = run_emu(emu, blob, strips, w, h, reverse=rev)
        assert gbad == 0
        assert np.array_equal(got, want)
        assert np.array_equal(got[:, :w], img)


@pytest.mark.parametrize("seed", range(12))
def test_random_payloads_and_over_read(emu, seed):
    """Random bits (valid once both length prefixes at column 0 are five zeros: any mix of lengths, kept
    lengths, raw groups), strips of every length around what the row needs: the tail of a row is decoded
    from the zero padding of the pump, or the row fails, exactly where the oracle says."""
    rng = np.random.default_rng(300 + seed)
    w, h = int(rng.choice([8, 24, 64, 90, 520])), 12
    stride = w * 2 + 64
    blob = rng.integers(0, 256, h * stride + 16, dtype=np.uint8)
    strips = []
    for r in range(h):
        need = int(rng.integers(4, stride))
        strips.append((r * stride + int(rng.integers(0, 4)), need, r))
    for off, _, _ in strips:   # MSB32: the first bits of the stream are the top bits of byte 3
        blob[off + 3] = 0
        blob[off + 2] &= 0x0F
    ok_rows = 0
    for r in range(h):         # row by row: the oracle stops at the first failing row
        want, bad = oracle_outcome(blob, [(strips[r][0], strips[r][1], 0)], w, 1)
        got, gbad = run_emu(emu, blob, [(strips[r][0], strips[r][1], 0)], w, 1)
        assert (gbad != 0) == (bad != 0), (r, strips[r], bad, gbad)
        if not bad:
            ok_rows += 1
            assert np.array_equal(got, want), (r, strips[r])
    want, bad = oracle_outcome(blob, strips, w, h)
    got, gbad = run_emu(emu, blob, strips, w, h)
    assert (gbad != 0) == (bad != 0)
    if not bad:
        assert np.array_equal(got, want)


def test_errors(emu):
    w, h = 16, 4
    img = synth.image_model(w, h, seed=2)
    blob, strips = synth.make_phaseone(img)
    bad = blob.copy()
    bad[strips[2][0] + 3] |= 0x80          # a 1 bit in the first length prefix at column 0
    assert run_emu(emu, bad, strips, w, h)[1] != 0
    bad = blob.copy()
    bad[strips[1][0] + 3] |= 0x02          # ... in the second one (bit 6 of the stream)
    assert oracle_outcome(bad, strips, w, h)[1] != 0
    assert run_emu(emu, bad, strips, w, h)[1] != 0
    short = [(o, 4, r) if r == 1 else (o, n, r) for o, n, r in strips]
    assert oracle_outcome(blob, short, w, h)[1] != 0
    assert run_emu(emu, blob, short, w, h)[1] != 0
    tiny = [(o, 3, r) if r == 0 else (o, n, r) for o, n, r in strips]   # below one chunk
    assert oracle_outcome(blob, tiny, w, h)[1] != 0
    assert run_emu(emu, blob, tiny, w, h)[1] != 0
