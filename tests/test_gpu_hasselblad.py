"""K2H parity: the Hasselblad decode on the GPU (rawspeed_b200/csrc/hasselblad.cuh) vs the oracle's
HasselbladDecompressor (pinned against the compiled reference in tests/test_oracle_hasselblad.py),
through the C ABI: pixels of the whole padded buffer, stream position, error class."""
import numpy as np
import pytest
import torch

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port, synth

pytestmark = pytest.mark.gpu
NCPL, VALS = synth.DEFAULT_NCPL, synth.DEFAULT_VALUES


def _job(w, h, in_offset, in_size, out_offset, init_pred, table=0):
    j = rs.HasselbladJob()
    j.in_offset, j.in_size, j.width, j.height = in_offset, in_size, w, h
    j.out_pitch, j.out_offset, j.init_pred, j.table = port.image_pitch(w), out_offset, init_pred, table
    return j


def _run(ctx, jobs, tables, blob, out_bytes):
    plan = rs.hasselblad_plan(ctx, tables, jobs)
    d_in = torch.zeros(len(blob) + 64, dtype=torch.uint8, device="cuda")
    d_in[:len(blob)] = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy())
    d_out = torch.full((out_bytes,), 0xA5, dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), len(blob)), d_out)
    torch.cuda.synchronize()
    return plan.results(check=False), d_out.cpu().numpy()


def _oracle(data, w, h, init_pred, ncpl=NCPL, vals=VALS):
    ht = port.Huff(ncpl, vals, full=False)
    img = port.new_image(w, h)
    try:
        c = port.hasselblad_decompress(img, w, ht, init_pred, bytes(data))
        return 0, c, img
    except port.IOException:
        return 2, None, img
    except port.RawDecoderException:
        return 1, None, img


def _check_one(ctx, data, w, h, init_pred):
    want = _oracle(data, w, h, init_pred)
    tab = rs.huff_table(bytes(NCPL), bytes(VALS), False)
    pitch = port.image_pitch(w)
    res, out = _run(ctx, [_job(w, h, 0, len(data), 0, init_pred)], [tab], data, h * pitch)
    status, consumed = res[0]
    assert status == want[0]
    if want[0] == 0:
        assert consumed == want[1]
        got = out.view(np.uint16).reshape(h, pitch // 2)
        assert np.array_equal(got[:, :w], want[2][:, :w])
    return out


@pytest.mark.parametrize("w,h,wild", [(2, 1, False), (66, 9, False), (130, 21, True), (1024, 300, False),
                                      (2048, 512, True), (8272, 1200, False)])
def test_hasselblad_matches_oracle(ctx, w, h, wild):
    img = synth.image_model(w, h, seed=w, wild=wild, bits=16 if wild else 14)
    if wild:
        img[0, 0:4] = [0x8000, 0x8000, 0, 0xFFFF]    # differences of -32768 and wrap-around
    ht = port.Huff(NCPL, VALS, full=False)
    data = synth.make_hasselblad_fast(img, ht, 0x8000) if hasattr(synth, "make_hasselblad_fast") else \
        synth.make_hasselblad(img, ht, 0x8000)
    out = _check_one(ctx, data, w, h, 0x8000)
    pitch = port.image_pitch(w)
    assert np.array_equal(out.view(np.uint16).reshape(h, pitch // 2)[:, :w], img)


def test_hasselblad_random_payloads(ctx):
    for seed in range(6):
        data = synth.lcg_bytes(4096, 9 + seed)
        _check_one(ctx, data, 64, 12, 0x2000)
        _check_one(ctx, data, 256, 40, 0x2000)     # needs more bits than there are


@pytest.mark.parametrize("cut", [0, 1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 15, 16, 17, 20, 24, 28, 33, 64, 200])
def test_hasselblad_truncated_streams(ctx, cut):
    img = synth.image_model(192, 16, seed=5)
    ht = port.Huff(NCPL, VALS, full=False)
    data = bytes(synth.make_hasselblad(img, ht, 0x8000))
    want0 = _oracle(data, 192, 16, 0x8000)
    assert want0[0] == 0
    end = want0[1]
    for base in (end + 16, end + 4, end, end - 1):
        n = base - cut
        if n > 0:
            _check_one(ctx, data[:n], 192, 16, 0x8000)


def test_hasselblad_two_frames_in_one_plan(ctx):
    """Two jobs, different sizes and predictors, streams at 4-byte aligned offsets of one buffer,
    images at different offsets of one output buffer."""
    ht = port.Huff(NCPL, VALS, full=False)
    a = synth.image_model(320, 40, seed=1)
    b = synth.image_model(130, 77, seed=2, wild=True, bits=16)
    da, db = bytes(synth.make_hasselblad(a, ht, 0x8000)), bytes(synth.make_hasselblad(b, ht, 0x1234))
    off_b = (len(da) + 3) // 4 * 4 + 8
    blob = da + bytes(off_b - len(da)) + db
    pa, pb = port.image_pitch(320), port.image_pitch(130)
    out_b = 40 * pa + 64
    tab = rs.huff_table(bytes(NCPL), bytes(VALS), False)
    jobs = [_job(320, 40, 0, len(da), 0, 0x8000), _job(130, 77, off_b, len(db), out_b, 0x1234)]
    res, out = _run(ctx, jobs, [tab], blob, out_b + 77 * pb)
    wa, wb = _oracle(da, 320, 40, 0x8000), _oracle(db, 130, 77, 0x1234)
    assert res[0] == (0, wa[1]) and res[1] == (0, wb[1])
    assert np.array_equal(out[:40 * pa].view(np.uint16).reshape(40, -1)[:, :320], a)
    assert np.array_equal(out[out_b:out_b + 77 * pb].view(np.uint16).reshape(77, -1)[:, :130], b)
    assert np.all(out[40 * pa:out_b] == 0xA5)      # nothing between the images was touched


def test_hasselblad_rejects_malformed_jobs(ctx):
    tab = rs.huff_table(bytes(NCPL), bytes(VALS), False)
    for bad in (dict(w=7), dict(w=12002), dict(h=0), dict(in_offset=2), dict(table=1)):
        kw = dict(w=8, h=2, in_offset=0, in_size=64, out_offset=0, init_pred=0, table=0)
        kw.update(bad)
        with pytest.raises(Exception):
            rs.hasselblad_plan(ctx, [tab], [_job(**kw)])


def test_hasselblad_ljpeg_decoder_of_the_host_mirror(ctx):
    """rawspeed_b200::HasselbladLJpegDecoder(bs, img).decode(): container walk on the host, pair
    stream on the device; the container is pinned against the reference's own decoder in
    tests/test_oracle_hasselblad.py."""
    w, h = 1024, 96
    img = synth.image_model(w, h, seed=3)
    ht = port.Huff(NCPL, VALS, full=False)
    data = synth.make_hasselblad_fast(img, ht, 0x8000)
    o = port.new_image(w, h)
    host.hasselblad_ljpeg_decode(synth.hasselblad_ljpeg_container(w, h, data, NCPL, VALS), o, w)
    assert np.array_equal(o[:, :w], img)
    with pytest.raises(host.RawDecoderException):      # frame does not match the image
        host.hasselblad_ljpeg_decode(synth.hasselblad_ljpeg_container(w, h, data, NCPL, VALS, frame_w=w + 2),
                                     port.new_image(w, h), w)
    with pytest.raises(host.RawDecoderException):      # restart interval
        host.hasselblad_ljpeg_decode(synth.hasselblad_ljpeg_container(w, h, data, NCPL, VALS, dri=4),
                                     port.new_image(w, h), w)
    with pytest.raises(host.IOException):              # the stream ends early
        host.hasselblad_ljpeg_decode(synth.hasselblad_ljpeg_container(w, h, data[:len(data) // 2], NCPL, VALS),
                                     port.new_image(w, h), w)
