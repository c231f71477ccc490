"""k2_tile_kernel (rawspeed_b200/csrc/ljpeg_tile.cuh) without a GPU: the kernel body itself is
compiled by g++ against tests/emu/cuda_emu.h (one fiber per CUDA thread; barriers, shuffles and
mbarrier waits are yield points of a deterministic scheduler) and run on the descriptors the
plan builder produces, then compared with the oracle -- pixels of the whole padded buffer,
`consumed`, status.  Every case runs in forward and in reverse thread order (a missing barrier
shows up as a difference) and for both geometries (R = 1, 2).  Parity of the real kernel is the
GPU tests' job (tests/test_gpu_ljpeg.py); this catches arithmetic, indexing and protocol slips
where there is no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import _abi
from oracle import port, synth
from helpers import dng_ljpeg_scans, compile_shared

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "ljpeg_tile_emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libljpeg_tile_emu.so")
CSRC = os.path.join(HERE, "..", "rawspeed_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_emu.h")] + [
    os.path.join(CSRC, f) for f in ("ljpeg_tile.cuh", "ljpeg_host.h", "ljpeg_types.h")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        compile_shared(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas",
                               "-Wno-unused-function", "-fPIC", "-shared", "-o", OUT, SRC])
    lib = C.CDLL(OUT)
    lib.tile_emu_run.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    return lib


def run_emu(lib, blob, tabs, scans, out, R=1, reverse=False, preroll=-1, npieces=0):
    tarr = (_abi.HuffTable * len(tabs))(*tabs)
    sarr = (_abi.LJpegScan * len(scans))(*scans)
    res = (_abi.ScanResult * len(scans))()
    blob = np.ascontiguousarray(blob)
    rc = lib.tile_emu_run(blob.ctypes.data, blob.size, tarr, len(tabs), sarr, len(scans),
                          out.ctypes.data, res, R, int(reverse), preroll, npieces)
    assert rc == 0, "emu rc %d (-1 = a scan is not eligible for the tile kernel)" % rc
    return [(r.status, r.consumed) for r in res]


def check_tiles(lib, img, tile_w, tile_h, Rs=(1, 2), preroll=-1, npieces=0, **kw):
    h, w = img.shape
    cpp = kw.pop("cpp", 1)
    w //= cpp
    t = synth.make_dng_ljpeg(img, tile_w, tile_h, cpp=cpp, **kw)
    want = port.new_image(w, h, cpp)
    port.dng_decompress(t.blob, t.offsets, t.lengths, want, w, cpp, tile_w, tile_h, 7,
                        fix_ljpeg=kw.get("fix16", False), nthreads=4)
    tabs, scans = dng_ljpeg_scans(t, want.shape[1] * 2, fix16=kw.get("fix16", False))
    for R in Rs:
        for rev in (False, True):
            got = port.new_image(w, h, cpp)
            res = run_emu(lib, t.blob, tabs.tabs, scans, got, R=R, reverse=rev, preroll=preroll,
                          npieces=npieces)
            assert all(s == 0 for s, _ in res), (R, rev, res[:4])
            bad = np.argwhere(got != want)
            assert bad.size == 0, (R, rev, bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
    return t, tabs, scans


def test_single_tile(emu):
    img = synth.image_model(256, 64, 1)
    check_tiles(emu, img, 256, 64)


def test_tiles_ragged_edges(emu):
    img = synth.image_model(600, 200, 7)
    check_tiles(emu, img, 256, 64)


def test_wild_noise_long_codes(emu):
    img = synth.image_model(512, 96, 9, wild=True)
    check_tiles(emu, img, 256, 32)


def test_sixteen_bit_ssss16(emu):
    img = synth.image_model(256, 64, 11, wild=True, bits=16)
    img[0, 0:8] = [0, 0x8000, 0, 0x8000, 0xFFFF, 0x7FFF, 0, 0x8000]
    check_tiles(emu, img, 256, 64, prec=16)
    check_tiles(emu, img, 256, 64, prec=16, fix16=True)


def test_components_1_and_4(emu):
    img = synth.image_model(512, 48, 13)
    check_tiles(emu, img, 256, 24, ncomp=1)
    check_tiles(emu, img, 512, 48, ncomp=4)


def test_restart_intervals(emu):
    img = synth.image_model(320, 96, 17)
    check_tiles(emu, img, 160, 48, restart_rows=1)
    check_tiles(emu, img, 160, 48, restart_rows=5)


def test_odd_width_crop_inside_a_unit(emu):
    img = synth.image_model(301, 33, 19)
    check_tiles(emu, img, 160, 16)


def test_many_chunks_and_small_chunks(emu):
    """A tile of several chunks; and the same with tiny chunks (16 pieces = 1 KiB) so that the
    carry between chunks, the deferred tail and the leftover differences are exercised a lot."""
    img = synth.image_model(512, 256, 23)
    check_tiles(emu, img, 512, 256)                       # ~130 KiB per tile
    check_tiles(emu, img, 256, 128, Rs=(1,), npieces=16)
    check_tiles(emu, img, 256, 128, Rs=(2,), npieces=40, preroll=96)


def test_flat_image_many_symbols_per_byte(emu):
    """Two bits per sample: a chunk holds several batches of the sample buffer."""
    img = np.full((128, 1024), 2000, dtype=np.uint16)
    check_tiles(emu, img, 512, 128)
    img[::7, ::5] += 3
    check_tiles(emu, img, 512, 128)


def test_stuffing_everywhere(emu):
    """Images whose streams are full of FF bytes (all-ones mantissas): most pieces are irregular."""
    rng = np.random.default_rng(5)
    img = np.zeros((64, 512), dtype=np.uint16)
    # alternating big positive differences: long runs of one-bits
    img[:, 0::2] = 0x3FFF
    img[:, 1::2] = 0
    img[::3, 2::4] = 0x3FFF
    check_tiles(emu, img, 256, 32)
    img = rng.integers(0, 1 << 14, size=(64, 512)).astype(np.uint16)
    check_tiles(emu, img, 256, 32)


def test_consumed_matches_the_oracle(emu):
    img = synth.image_model(512, 64, 21)
    t, tabs, scans = check_tiles(emu, img, 256, 32)
    hts = synth.default_tables(1)
    got = port.new_image(512, 64)
    for R in (1, 2):
        res = run_emu(emu, t.blob, tabs.tabs, scans, got, R=R)
        for (status, consumed), s, off, ln in zip(res, scans, t.offsets, t.lengths):
            data = t.blob[s.in_offset:off + ln]
            o = port.new_image(512, 64)
            want = port.ljpeg_decompress(o, 512, 1, (s.out_x, s.out_y, s.store_w, s.rows),
                                         (2, 1), (s.frame_w, s.rows), [hts[0], hts[0]],
                                         [1 << 13] * 2, s.rows, data)
            assert (status, consumed) == (0, want)


def _one_scan_outcome(s, data, hts, w, h):
    """The oracle on one segment: (status, consumed, image)."""
    o = port.new_image(w, h)
    try:
        c = port.ljpeg_decompress(o, w, 1, (s.out_x, s.out_y, s.store_w, s.rows), (2, 1),
                                  (s.frame_w, s.rows), [hts[0], hts[0]], [1 << 13] * 2, s.rows, data)
        return 0, c, o
    except port.IOException:
        return 2, None, o
    except port.RawDecoderException:
        return 1, None, o


@pytest.mark.parametrize("cut", list(range(1, 34)) + [40, 64, 100])
def test_streams_that_end_early(emu, cut):
    """A segment truncated by `cut` bytes (the buffer ends; no marker): the reference reads zero
    bits behind the data and throws only when its pump gets more than 16 bytes past the buffer
    (BitStreamer.h:120-127) -- same pixels, same `consumed`, same status here."""
    img = synth.image_model(256, 32, 53)
    t = synth.make_dng_ljpeg(img, 256, 32)
    tabs, scans = dng_ljpeg_scans(t, port.image_pitch(256))
    hts = synth.default_tables(1)
    s = rs.LJpegScan.from_buffer_copy(scans[0])
    s.in_size = scans[0].in_size - cut
    blob = t.blob[:s.in_offset + s.in_size].copy()
    want_status, want_cons, want_img = _one_scan_outcome(s, blob[s.in_offset:], hts, 256, 32)
    for R in (1, 2):
        got = port.new_image(256, 32)
        (status, consumed), = run_emu(emu, blob, tabs.tabs, [s], got, R=R)
        assert status == want_status, (R, cut)
        if want_status == 0:
            assert consumed == want_cons
            assert np.array_equal(got, want_img)


@pytest.mark.parametrize("cut", [2, 3, 5, 8, 11, 16, 19, 24, 27, 32, 40])
def test_streams_with_an_early_marker(emu, cut):
    """The same with an end marker `cut` bytes before the true end of the data."""
    img = synth.image_model(256, 32, 57)
    t = synth.make_dng_ljpeg(img, 256, 32)
    tabs, scans = dng_ljpeg_scans(t, port.image_pitch(256))
    hts = synth.default_tables(1)
    s = rs.LJpegScan.from_buffer_copy(scans[0])
    blob = t.blob.copy()
    end = s.in_offset + s.in_size          # behind EOI
    pos = end - 2 - cut
    if blob[pos - 1] == 0xFF:              # do not turn a stuffing pair into something else
        pos -= 2
    blob[pos] = 0xFF
    blob[pos + 1] = 0xD9
    want_status, want_cons, want_img = _one_scan_outcome(s, blob[s.in_offset:end], hts, 256, 32)
    for R in (1, 2):
        got = port.new_image(256, 32)
        (status, consumed), = run_emu(emu, blob[:end], tabs.tabs, [s], got, R=R)
        assert status == want_status, (R, cut)
        if want_status == 0:
            assert consumed == want_cons
            assert np.array_equal(got, want_img)


def test_garbage_behind_the_last_symbol(emu):
    """Rows below the crop are not decoded: the stream goes on behind the last needed symbol."""
    img = synth.image_model(256, 64, 59)
    t = synth.make_dng_ljpeg(img, 256, 64)
    tabs, scans = dng_ljpeg_scans(t, port.image_pitch(256))
    hts = synth.default_tables(1)
    s = rs.LJpegScan.from_buffer_copy(scans[0])
    s.rows = 40
    want_status, want_cons, want_img = _one_scan_outcome(s, t.blob[s.in_offset:], hts, 256, 64)
    for R in (1, 2):
        got = port.new_image(256, 64)
        (status, consumed), = run_emu(emu, t.blob, tabs.tabs, [s], got, R=R)
        assert (status, consumed) == (want_status, want_cons)
        assert np.array_equal(got, want_img)


def test_bad_huffman_code(emu):
    img = synth.image_model(256, 32, 23, wild=True)
    t = synth.make_dng_ljpeg(img, 256, 32)
    tabs, scans = dng_ljpeg_scans(t, port.image_pitch(256))
    blob = t.blob.copy()
    s = scans[0]
    blob[s.in_offset + 40:s.in_offset + 49] = [0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFE]
    for R in (1, 2):
        (status, _), = run_emu(emu, blob, tabs.tabs, scans[:1], port.new_image(256, 32), R=R)
        assert status == 1


@pytest.mark.parametrize("seed", range(16))
def test_random_tables(emu, seed):
    """Differential fuzz with random COMPLETE canonical codes (lengths up to 16 bits, values in random
    order): the self-synchronising parse must lock onto the true symbol boundaries whatever the code
    looks like, 1 / 2 / 4 components, noise from a few bits to the full 14-bit range."""
    from test_ljpeg_stream_emu import _random_table
    rng = np.random.default_rng(9000 + seed)
    ncomp = int(rng.choice([1, 2, 4]))
    bits = int(rng.choice([3, 6, 10, 14]))
    h, tw = int(rng.choice([16, 32])), int(rng.choice([128, 256]))
    w = tw * int(rng.integers(1, 3))
    img = (8192 + rng.integers(0, 1 << bits, (h, w)) - (1 << bits) // 2).astype(np.uint16) & 0x3FFF
    check_tiles(emu, img, tw, h, Rs=(1,), ncomp=ncomp, tabs=[_random_table(rng, 16)])
