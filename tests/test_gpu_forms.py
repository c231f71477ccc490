"""K1b parity: the fixed-layout UncompressedDecompressor forms on the GPU vs the
oracle, through the C ABI.  Bit-exact (integer work; the float forms are compared
as uint32 bit patterns, NaN payloads included)."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import formats as F
from oracle import port, synth

pytestmark = pytest.mark.gpu


def run(plan, data, out_np):
    import torch
    d_in = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:data.size] = torch.from_numpy(np.ascontiguousarray(data))
    d_out = torch.from_numpy(out_np.view(np.uint8).reshape(-1).copy()).cuda()
    plan.run((d_in.data_ptr(), data.size), d_out)
    torch.cuda.synchronize()
    plan.results()
    return d_out.cpu().numpy().view(out_np.dtype).reshape(out_np.shape)


def job(fmt, data_size, img, samples, rows, in_pitch, row0=0, col0=0, in_offset=0, table=0):
    j = rs.RawJob()
    j.in_offset, j.in_size, j.out_offset = in_offset, data_size, 0
    j.out_pitch = img.shape[1] * img.itemsize
    j.row0, j.rows, j.samples, j.out_col0 = row0, rows, samples, col0
    j.in_pitch, j.format, j.table = in_pitch, fmt, table
    return j


@pytest.mark.parametrize("w,h", [(70, 9), (8, 1), (3, 4), (4000, 30)])
def test_8bit_plain_and_table(ctx, w, h):
    data = synth.lcg_bytes(w * h, seed=3)
    curve = (np.arange(256, dtype=np.uint32) ** 2 // 2 % 65536).astype(np.uint16)
    curve[77] = 3
    for kind in ("uncorrected", "plain", "dither"):
        want = port.new_image(w, h)
        got0 = want.copy()
        if kind == "uncorrected":
            port.unpack_form(data, want, w, 1, (0, 0, w, h), w, 8, port.LSB, port.FORM_8BIT_UNCORRECTED)
            plan = rs.raw_plan(ctx, [job(F.RAW_8BIT, data.size, got0, w, h, w)])
        else:
            dither = kind == "dither"
            t = port.build_table(curve, dither)
            port.unpack_form(data, want, w, 1, (0, 0, w, h), w, 8, port.LSB, port.FORM_8BIT, t, dither)
            dev_table = t[0::2] if dither else t   # dither counter stays 0: base only
            plan = rs.raw_plan(ctx, [job(F.RAW_8BIT_TABLE, data.size, got0, w, h, w)], dev_table)
        got = run(plan, data, got0)
        assert np.array_equal(got, want), kind


@pytest.mark.parametrize("w", [10, 20, 38, 46, 64, 100, 4000])
def test_12bit_with_control(ctx, w):
    h = 7
    perline = 12 * w // 8 + (w + 2) // 10
    data = synth.lcg_bytes(perline * h, seed=w)
    for fmt, form in ((F.RAW_12BIT_CONTROL_BE, port.FORM_12BIT_CONTROL_BE),
                      (F.RAW_12BIT_CONTROL_LE, port.FORM_12BIT_CONTROL_LE)):
        want = port.new_image(w, h)
        got0 = want.copy()
        port.unpack_form(data, want, w, 1, (0, 0, w, h), perline, 12, port.MSB, form)
        plan = rs.raw_plan(ctx, [job(fmt, data.size, got0, w, h, perline)])
        assert np.array_equal(run(plan, data, got0), want), (w, fmt)


@pytest.mark.parametrize("w", [37, 8, 1, 1000])
def test_12bit_left_aligned(ctx, w):
    h = 5
    data = synth.lcg_bytes(2 * w * h, seed=9)
    for fmt, form in ((F.RAW_12BIT_LEFT_BE, port.FORM_12BIT_LEFT_BE),
                      (F.RAW_12BIT_LEFT_LE, port.FORM_12BIT_LEFT_LE)):
        want = port.new_image(w, h)
        got0 = want.copy()
        port.unpack_form(data, want, w, 1, (0, 0, w, h), 2 * w, 16, port.LSB, form)
        plan = rs.raw_plan(ctx, [job(fmt, data.size, got0, w, h, 2 * w)])
        assert np.array_equal(run(plan, data, got0), want), (w, fmt)


@pytest.mark.parametrize("order", [port.LSB, port.MSB])
@pytest.mark.parametrize("bps", [16, 24, 32])
@pytest.mark.parametrize("cpp", [1, 3])
def test_float_forms(ctx, order, bps, cpp):
    w, h, ox, oy = 24, 6, 4, 1
    W, H = w + 8, h + 2
    pitch = w * cpp * bps // 8 + 4
    rng = np.random.default_rng(bps * 10 + cpp)
    data = rng.integers(0, 256, pitch * h, dtype=np.uint8)
    # force every exponent class to appear: zero / subnormal / inf / NaN
    nb = bps // 8
    for i in range(0, min(200, data.size - nb), 5 * nb):
        hi = i if order == port.MSB else i + nb - 1
        data[hi] = [0x00, 0x80, 0x7C if bps == 16 else 0x7F, 0xFC if bps == 16 else 0xFF][(i // (5 * nb)) % 4]
    want = port.new_image_f32(W, H, cpp)
    got0 = want.copy()
    port.unpack_form(data, want, W, cpp, (ox, oy, w, h), pitch, bps, order, port.FORM_READ)
    fmt = {(16, port.MSB): F.RAW_FP16_MSB, (16, port.LSB): F.RAW_FP16_LSB,
           (24, port.MSB): F.RAW_FP24_MSB, (24, port.LSB): F.RAW_FP24_LSB}.get((bps, order), F.RAW_F32_COPY)
    # decodePackedFP writes at column offset.x; the 32-bit copy at offset.x*cpp
    col0 = ox * cpp if bps == 32 else ox
    plan = rs.raw_plan(ctx, [job(fmt, data.size, got0, w * cpp, h, pitch, row0=oy, col0=col0)])
    got = run(plan, data, got0)
    assert np.array_equal(got, want)


def test_fp16_all_values(ctx):
    """Exhaustive: all 65536 binary16 patterns through the LSB form."""
    w, h = 256, 256
    data = np.arange(65536, dtype=np.uint16).view(np.uint8)
    want = port.new_image_f32(w, h)
    got0 = want.copy()
    port.unpack_form(data, want, w, 1, (0, 0, w, h), 2 * w, 16, port.LSB, port.FORM_READ)
    plan = rs.raw_plan(ctx, [job(F.RAW_FP16_LSB, data.size, got0, w, h, 2 * w)])
    got = run(plan, data, got0)
    assert np.array_equal(got, want)
    fin = np.isfinite(np.arange(65536, dtype=np.uint16).view(np.float16))
    ref = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32).view(np.uint32)
    assert np.array_equal(got[:, :w].reshape(-1)[fin], ref[fin])


def test_several_jobs_mixed_formats_and_offsets(ctx):
    """Tiles of different formats in one plan, unaligned input offsets."""
    w, h = 50, 6
    per12 = 12 * w // 8 + (w + 2) // 10
    blobs = [synth.lcg_bytes(w * h, 1), synth.lcg_bytes(per12 * h, 2), synth.lcg_bytes(2 * w * h, 3)]
    offs, buf = [], bytearray(b"\x00" * 3)
    for b in blobs:
        offs.append(len(buf))
        buf += b.tobytes() + b"\x00" * 5
    data = np.frombuffer(bytes(buf), dtype=np.uint8)
    want = port.new_image(w, 3 * h)
    got0 = want.copy()
    forms = [(F.RAW_8BIT, port.FORM_8BIT_UNCORRECTED, w, 8, port.LSB),
             (F.RAW_12BIT_CONTROL_LE, port.FORM_12BIT_CONTROL_LE, per12, 12, port.MSB),
             (F.RAW_12BIT_LEFT_BE, port.FORM_12BIT_LEFT_BE, 2 * w, 16, port.LSB)]
    jobs = []
    for k, (fmt, form, pitch, bps, order) in enumerate(forms):
        sub = want[k * h:(k + 1) * h]
        port.unpack_form(blobs[k], sub, w, 1, (0, 0, w, h), pitch, bps, order, form)
        jobs.append(job(fmt, blobs[k].size, got0, w, h, pitch, row0=k * h, in_offset=offs[k]))
    plan = rs.raw_plan(ctx, jobs)
    assert plan.launches == 3
    assert np.array_equal(run(plan, data, got0), want)


def test_malformed_jobs_rejected(ctx):
    img = port.new_image(16, 4)
    bad = job(F.RAW_12BIT_CONTROL_BE, 100, img, 7, 2, 20)      # odd width
    with pytest.raises(rs.Rsb200Error):
        rs.raw_plan(ctx, [bad])
    bad = job(F.RAW_8BIT, 10, img, 16, 4, 16)                   # not enough input
    with pytest.raises(rs.Rsb200Error):
        rs.raw_plan(ctx, [bad])
    bad = job(F.RAW_8BIT_TABLE, 64, img, 16, 4, 16)             # table missing
    with pytest.raises(rs.Rsb200Error):
        rs.raw_plan(ctx, [bad])
    bad = job(99, 64, img, 16, 4, 16)
    with pytest.raises(rs.Rsb200Error):
        rs.raw_plan(ctx, [bad])
