"""K10 (fused DNG opcode pass) without a GPU: the host mirror parses and lowers the list
(DngOpcodes::lower -- the product's own code), the kernel's per-thread program (dngop_core.h)
and descriptor builder (dngop_host.h) are compiled as plain C++ and replayed on the CPU
(tests/emu/dngop_emu.cpp), and the outcome -- pixels, crop, mBadPixelPositions in order -- is
compared with the oracle (pinned against the compiled reference in
tests/test_oracle_dngopcodes.py).  The host-side list / crop bookkeeping of applyOpCodes is
repeated here in Python (the real one runs in the GPU test)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import compile_shared

from oracle import port
from rawspeed_b200 import host
from rawspeed_b200._abi import DngOp, DngOpJob
from test_oracle_dngopcodes import scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "dngop_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libdngop_emu%s.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "rawspeed_b200", "csrc", f)
                for f in ("dngop_core.h", "dngop_host.h")]


@pytest.fixture(scope="module", params=["", "_v1", "_v2"])
def emu(request):
    """The versions of the opcode walk: the one the library ships (the third -- closed-form hit masks --
    where a pixel is one sample, the second elsewhere), the second alone (-DRSB200_DNGOP_V2) and the
    first one (-DRSB200_DNGOP_V1)."""
    out = OUT % request.param
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        flags = ["-DRSB200_DNGOP" + request.param.upper()] if request.param else []
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared"] + flags + ["-o", out, SRC])
    lib = C.CDLL(out)
    lib.dngop_emu_run.argtypes = [C.c_void_p, C.POINTER(DngOpJob), C.c_int, C.POINTER(DngOp), C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                                  C.c_char_p, C.c_int]
    return lib


def replay(emu, img, w, cpp, crop, blob):
    """What DngOpcodes(ri, blob).applyOpCodes(ri) does, with the device pass replayed on the CPU.
    Returns (crop, bad list, error or None); img is modified in place."""
    low = host.dngop_lower(img, w, cpp, crop, blob)
    job = DngOpJob()
    job.offset, job.pitch, job.width, job.height = 0, img.shape[1] * img.itemsize, w, img.shape[0]
    job.cpp, job.is_f32, job.first_op, job.num_ops = cpp, int(img.dtype == np.uint32), 0, len(low["ops"])
    const = {}
    if low["ops"]:
        ops = (DngOp * len(low["ops"]))(*low["ops"])
        bad = np.zeros(1 << 20, dtype=np.uint32)
        err = C.create_string_buffer(256)
        tables, deltas = low["tables"], low["deltas"]
        n = emu.dngop_emu_run(img.ctypes.data, (DngOpJob * 1)(job), 1, ops, len(low["ops"]),
                              tables.ctypes.data, tables.shape[0], deltas.ctypes.data, deltas.size,
                              bad.ctypes.data, bad.size, err, 256)
        assert n >= 0, err.value
        i = 0
        while i < n:
            k, cnt = int(bad[i]), int(bad[i + 1])
            const[k] = sorted(bad[i + 2:i + 2 + cnt].tolist())
            i += 2 + cnt
    crop = list(crop)
    lst = []
    for kind, index, extra in low["actions"]:
        if kind == 0:
            lst = list(extra) + lst
        elif kind == 1:
            lst = lst + const[index]
        else:   # subFrame(roi)
            x, y, rw, rh = extra
            if rw <= crop[2] - x and rh <= crop[3] - y:
                crop = [crop[0] + x, crop[1] + y, rw, rh]
    return crop, lst, low["error"]


@pytest.mark.parametrize("k", range(9))
def test_replayed_pass_matches_oracle(emu, k):
    name, img, w, cpp, crop, blob = scenarios()[k]
    want = img.copy()
    got = img.copy()
    try:
        wcrop, wbad = port.dng_opcodes(want, w, cpp, crop, blob)
        werr = None
    except Exception as ex:   # noqa: BLE001
        werr = ex
        wcrop, wbad = port.dng_opcodes.partial[:2]
    gcrop, gbad, gerr = replay(emu, got, w, cpp, crop, blob)
    assert np.array_equal(got, want)
    assert gcrop == wcrop and gbad == wbad
    assert (gerr is None) == (werr is None)
    if werr is not None:
        assert type(gerr).__name__ == type(werr).__name__


def test_constructor_errors_same_class():
    from test_oracle_dngopcodes import u16_image, FULL
    from oracle import synth as S
    w, h = 32, 8
    img = u16_image(w, h, 1, 9)
    area = S.dng_pixel_area(FULL(w, h))
    ok = S.dng_delta(10, area, np.zeros(8, np.float32))
    bad_count = S.dng_opcode_list([ok])
    bad_count[3] = 2
    for blob in (S.dng_opcode_list([ok])[:-3], bad_count, S.dng_opcode_list([(99, b"")]),
                 S.dng_opcode_list([(9, b"\x00" * 8, 0)]), S.dng_opcode_list([(9, b"\x00" * 8, 1)]),
                 S.dng_opcode_list([S.dng_delta(10, S.dng_pixel_area((0, 0, h + 1, w)), np.zeros(9, np.float32))]),
                 S.dng_opcode_list([S.dng_trim_bounds(5, 5, 4, 10)]),
                 S.dng_opcode_list([S.dng_delta(10, S.dng_pixel_area(FULL(w, h), 1, 1), np.zeros(8, np.float32))]),
                 S.dng_opcode_list([S.dng_delta(10, S.dng_pixel_area(FULL(w, h), 0, 1, 0, 1), np.zeros(8, np.float32))]),
                 S.dng_opcode_list([S.dng_delta(11, area, np.zeros(8, np.float32))]),
                 S.dng_opcode_list([S.dng_delta(10, area, np.array([np.nan] * 8, np.float32))]),
                 S.dng_opcode_list([S.dng_map_table(area, np.zeros(0, np.uint16))]),
                 S.dng_opcode_list([S.dng_map_polynomial(area, [0.0] * 10)]),
                 S.dng_opcode_list([(6, S.dng_roi(0, 0, h, w) + b"\x00")]),
                 S.dng_opcode_list([S.dng_fix_bad_list(points=[(h, 0)])]),
                 S.dng_opcode_list([S.dng_fix_bad_list(rects=[(0, 0, h + 1, 2)])]),
                 S.dng_opcode_list([(5, b"\x00\x00\x00\x00\xff\xff\xff\xff\x00\x00\x00\x00")])):
        with pytest.raises(Exception) as a:
            port.dng_opcodes(img.copy(), w, 1, [0, 0, w, h], blob)
        with pytest.raises(Exception) as b:
            host.dngop_lower(img.copy(), w, 1, [0, 0, w, h], blob)
        assert type(a.value).__name__ == type(b.value).__name__, (a.value, b.value)


def test_setup_errors_on_wrong_image_type():
    from test_oracle_dngopcodes import u16_image, f32_image, FULL
    from oracle import synth as S
    w, h = 16, 4
    area = S.dng_pixel_area(FULL(w, h))
    for img, cpp, blob in ((f32_image(w, h, 1, 1), 1, S.dng_opcode_list([S.dng_map_table(area, np.arange(16, dtype=np.uint16))])),
                           (f32_image(w, h, 1, 1), 1, S.dng_opcode_list([S.dng_fix_bad_constant(0)])),
                           (u16_image(w, h, 3, 2), 3, S.dng_opcode_list([S.dng_fix_bad_constant(0)]))):
        low = host.dngop_lower(img, w, cpp, [0, 0, w, h], blob)
        assert low["error"] is not None and not low["ops"]
        with pytest.raises(Exception) as a:
            port.dng_opcodes(img.copy(), w, cpp, [0, 0, w, h], blob)
        assert type(a.value).__name__ == type(low["error"]).__name__
