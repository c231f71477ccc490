"""K1 parity: CUDA packed unpack vs the oracle, through the C ABI.
Bit-exact (integer path)."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from oracle import port, synth
from helpers import gpu_run

pytestmark = pytest.mark.gpu


def _job(in_size, img, w_samples, rows, row0, in_pitch, bps, order, in_offset=0, col0=0):
    j = rs.UnpackJob()
    j.in_offset, j.in_size, j.out_offset = in_offset, in_size, 0
    j.out_pitch = img.shape[1] * 2
    j.row0, j.rows, j.samples, j.out_col0 = row0, rows, w_samples, col0
    j.in_pitch, j.bps, j.order = in_pitch, bps, order
    return j


@pytest.mark.parametrize("order", [rs.LSB, rs.MSB, rs.MSB16, rs.MSB32])
@pytest.mark.parametrize("bps", [10, 12, 14, 8, 16, 7, 13, 1])
def test_unpack_small(ctx, order, bps):
    w, h = 264, 9          # 264*bps % 8 == 0 for every bps
    for skip in (0, 5):
        data, pitch = synth.packed_frame(w, h, bps, seed=100 + bps, pitch=w * bps // 8 + skip)
        want = port.new_image(w, h + 3)
        got0 = want.copy()
        port.unpack(data, want, w, 1, (0, 2, w, h), pitch, bps, order)
        plan = rs.unpack_plan(ctx, [_job(data.size, got0, w, h, 2, pitch, bps, order)])
        got, _ = gpu_run(plan, data, got0)
        assert np.array_equal(got, want), (order, bps, skip)


def test_unpack_ragged_width(ctx):
    # widths that are not a multiple of 8 samples / odd pitches / cpp=3
    for (w, cpp, bps) in [(2, 1, 12), (6, 1, 12), (10, 3, 12), (4, 1, 14), (12, 1, 10), (20, 2, 14)]:
        for order in (rs.MSB, rs.LSB, rs.MSB32, rs.MSB16):
            h = 7
            pitch = w * cpp * bps // 8 + 3
            data = synth.lcg_bytes(pitch * h, 5)
            want = port.new_image(w, h, cpp)
            got0 = want.copy()
            port.unpack(data, want, w, cpp, (0, 0, w, h), pitch, bps, order)
            plan = rs.unpack_plan(ctx, [_job(data.size, got0, w * cpp, h, 0, pitch, bps, order)])
            got, _ = gpu_run(plan, data, got0)
            assert np.array_equal(got, want), (w, cpp, bps, order)


def test_unpack_c1_12bit_4000x3000(ctx):
    """BASELINE configs[0]: 12-bit packed 4000x3000 (MSB and LSB)."""
    w, h, bps = 4000, 3000, 12
    data, pitch = synth.packed_frame(w, h, bps, seed=1)
    for order in (rs.MSB, rs.LSB):
        want = port.new_image(w, h)
        got0 = want.copy()
        port.unpack(data, want, w, 1, (0, 0, w, h), pitch, bps, order)
        plan = rs.unpack_plan(ctx, [_job(data.size, got0, w, h, 0, pitch, bps, order)])
        got, _ = gpu_run(plan, data, got0)
        assert np.array_equal(got, want)


def test_unpack_c2_14bit_45mp(ctx):
    """BASELINE configs[1]: 14-bit packed 8256x5504, all four bit orders;
    full-size check through the rstest hash recipe + full compare."""
    w, h, bps = 8256, 5504, 14
    data, pitch = synth.packed_frame(w, h, bps, seed=2)
    for order in (rs.MSB, rs.LSB, rs.MSB16, rs.MSB32):
        want = port.new_image(w, h)
        got0 = want.copy()
        port.unpack(data, want, w, 1, (0, 0, w, h), pitch, bps, order)
        plan = rs.unpack_plan(ctx, [_job(data.size, got0, w, h, 0, pitch, bps, order)])
        got, _ = gpu_run(plan, data, got0)
        assert np.array_equal(got, want)
    assert synth.md5_of_row_md5s(got) == synth.md5_of_row_md5s(want)


def test_unpack_unaligned_offsets_and_tiles(ctx):
    """DNG compression-1 tiles: several jobs in one plan, arbitrary byte offsets
    (AbstractDngDecompressor.cpp:54-110).  One tile column: the reference writes
    packed integers at column 0 of the row whatever the tile's x offset
    (UncompressedDecompressor.cpp:196), so several tile columns would race in
    the reference itself."""
    W, H, tw, th, bps = 96, 60, 96, 16, 12
    rng = np.random.default_rng(3)
    tiles_y = (H + th - 1) // th
    pitch = tw * bps // 8
    blob = np.zeros(0, dtype=np.uint8)
    offs = []
    for n in range(tiles_y):
        pad = rng.integers(0, 7)
        offs.append(blob.size + pad)
        blob = np.concatenate([blob, np.zeros(pad, np.uint8),
                               rng.integers(0, 256, pitch * th, dtype=np.uint8)])
    want = port.new_image(W, H)
    got0 = want.copy()
    port.dng_decompress(blob, offs, [pitch * th] * len(offs), want, W, 1, tw, th, 1, bps=bps)
    jobs = []
    for n, off in enumerate(offs):
        h = min(th, H - n * th)
        jobs.append(_job(pitch * th, got0, W, h, n * th, pitch, bps, rs.MSB, in_offset=off))
    plan = rs.unpack_plan(ctx, jobs)
    got, _ = gpu_run(plan, blob, got0)
    assert np.array_equal(got, want)


def test_unpack_16bit_lsb_tiles_honour_x_offset(ctx):
    """bps == 16 LSB is a row copy that does honour offset.x
    (UncompressedDecompressor.cpp:255-264): several tile columns are well defined."""
    W, H, tw, th = 100, 40, 32, 16
    tiles_x, tiles_y = (W + tw - 1) // tw, (H + th - 1) // th
    pitch = tw * 2
    blob = synth.lcg_bytes(pitch * th * tiles_x * tiles_y + 13, 4)
    offs = [13 + n * pitch * th for n in range(tiles_x * tiles_y)]
    want = port.new_image(W, H)
    got0 = want.copy()
    port.dng_decompress(blob, offs, [pitch * th] * len(offs), want, W, 1, tw, th, 1, bps=16)
    jobs = []
    for n, off in enumerate(offs):
        ty, tx = divmod(n, tiles_x)
        w = min(tw, W - tx * tw)
        h = min(th, H - ty * th)
        jobs.append(_job(pitch * th, got0, w, h, ty * th, pitch, 16, rs.LSB, in_offset=off,
                         col0=tx * tw))
    plan = rs.unpack_plan(ctx, jobs)
    got, _ = gpu_run(plan, blob, got0)
    assert np.array_equal(got, want)
