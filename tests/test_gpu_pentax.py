"""Pentax PEF parity: PentaxDecompressor on the GPU (plain-MSB Huffman stream through
the multi-CTA parse + the K3P reconstruction kernels) vs the oracle, through the C ABI
and through the C++ host mirror.  Integer work: bit-exact."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host, _abi
from oracle import port, synth
from helpers import gpu_run

pytestmark = pytest.mark.gpu


def plan_for(ctx, table, data_size, img, w, h, in_offset=0, out_offset=0):
    j = rs.PentaxJob()
    j.in_offset, j.in_size, j.table = in_offset, data_size, 0
    j.width, j.height = w, h
    j.out_offset, j.out_pitch = out_offset, img.shape[1] * 2
    return rs.pentax_plan(ctx, [rs.huff_table(table[0], table[1])], [j])


@pytest.mark.parametrize("meta_kind", ["legacy", "modern_be", "modern_le"])
@pytest.mark.parametrize("w,h", [(2, 1), (6, 2), (64, 9), (500, 40), (1000, 333)])
def test_pentax_abi_and_host(ctx, meta_kind, w, h):
    meta = None if meta_kind == "legacy" else synth.pentax_modern_meta(meta_kind == "modern_be")
    be = meta_kind != "modern_le"
    table = port.pentax_table(meta, be)
    img = (synth.image_model(w, h, seed=w + h, bits=12) & 0x0FFF).astype(np.uint16)
    data = synth.make_pentax(img, table)
    want = port.new_image(w, h)
    got0 = want.copy()
    port.pentax_decompress(want, w, data, meta, be)
    assert np.array_equal(want[:, :w], img)
    got, res = gpu_run(plan_for(ctx, table, data.size, got0, w, h), data, got0)
    assert res[0][0] == 0
    assert np.array_equal(got, want)
    b = got0.copy()
    host.pentax_decompress(b, w, data, meta, be)
    assert np.array_equal(b, want)


def test_pentax_full_frame_k3(ctx):
    """A K-3 sized frame (6080x4032... here 6016x4000), several hundred ranges."""
    w, h = 6016, 4000
    table = port.pentax_table(None)
    img = (synth.image_model(w, h, seed=11, bits=12) & 0x0FFF).astype(np.uint16)
    data = synth.make_pentax(img, table)
    got0 = port.new_image(w, h)
    got, res = gpu_run(plan_for(ctx, table, data.size, got0, w, h), data, got0)
    assert res[0][0] == 0
    assert np.array_equal(got[:, :w], img)


def test_pentax_wild_noise_long_codes(ctx):
    w, h = 640, 200
    meta = synth.pentax_modern_meta(True)
    table = port.pentax_table(meta, True)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 16384, (h, w), dtype=np.uint16)   # 14-bit differences, 12-bit codes
    data = synth.make_pentax(img, table)
    want = port.new_image(w, h)
    got0 = want.copy()
    port.pentax_decompress(want, w, data, meta, True)
    got, res = gpu_run(plan_for(ctx, table, data.size, got0, w, h), data, got0)
    assert res[0][0] == 0 and np.array_equal(got, want)


def _oob_stream(table, w, h, where, value):
    d = np.zeros((h, w), dtype=np.int32)
    d[where] = value
    return port.encode_diffs_plain(d.reshape(-1), port.Huff(*table))


@pytest.mark.parametrize("where,value,msg", [((0, 8), 16383, "8:0"), ((1, 3), -5, "3:1"),
                                              ((5, 0), -1, "0:5"), ((7, 31), -2, "31:7")])
def test_pentax_out_of_bounds(ctx, where, value, msg):
    """isIntN(value, 16): the first pixel (stream order) outside 0..65535 is reported."""
    w, h = 32, 10
    meta = synth.pentax_modern_meta(True)
    table = port.pentax_table(meta, True)
    d = np.zeros((h, w), dtype=np.int32)
    if value > 0:
        d[0, 0:8:2] = 16383            # 65532 at (0, 6); + 16383 at (0, 8) overflows
    d[where] = value
    d[9, 20] = -7                      # a later violation must not win
    data = port.encode_diffs_plain(d.reshape(-1), port.Huff(*table))
    with pytest.raises(port.RawDecoderException) as ei:
        port.pentax_decompress(port.new_image(w, h), w, data, meta, True)
    assert msg in ei.value.msg
    got0 = port.new_image(w, h)
    plan = plan_for(ctx, table, data.size, got0, w, h)
    import torch
    d_in = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:data.size] = torch.from_numpy(data)
    d_out = torch.from_numpy(got0.view(np.int16).copy()).cuda()
    plan.run((d_in.data_ptr(), data.size), d_out)
    res = plan.results(check=False)
    row, col = (int(x) for x in reversed(msg.split(":")))
    assert res[0][0] == _abi.ERR_RDE
    assert res[0][1] == _abi.PENTAX_OOB | (row << 14) | col
    with pytest.raises(rs.RawDecoderException) as e2:
        host.pentax_decompress(port.new_image(w, h), w, data, meta, True)
    assert msg in str(e2.value)


def test_pentax_truncated_and_corrupt(ctx):
    w, h = 64, 16
    table = port.pentax_table(None)
    img = (synth.image_model(w, h, seed=3, bits=12) & 0x0FFF).astype(np.uint16)
    data = synth.make_pentax(img, table)
    with pytest.raises(rs.IOException):
        host.pentax_decompress(port.new_image(w, h), w, data[:len(data) // 3])
    with pytest.raises(rs.RawDecoderException):
        host.pentax_decompress(port.new_image(7, 2), 7, data)          # odd width
    meta = bytearray(synth.pentax_modern_meta(True))
    meta[1] = 9
    with pytest.raises(rs.RawDecoderException):
        host.pentax_decompress(port.new_image(w, h), w, data, bytes(meta))
    # a stream of all ones: the legacy table has no code 1111111111 (10 ones)
    bad = np.full(4096, 0xFF, dtype=np.uint8)
    with pytest.raises(port.RawDecoderException):
        port.pentax_decompress(port.new_image(w, h), w, bad)
    with pytest.raises(rs.RawDecoderException):
        host.pentax_decompress(port.new_image(w, h), w, bad)


def test_pentax_two_images_one_plan(ctx):
    w, h = 128, 50
    table = port.pentax_table(None)
    imgs = [(synth.image_model(w, h, seed=s, bits=12) & 0x0FFF).astype(np.uint16) for s in (1, 2)]
    datas = [synth.make_pentax(i, table) for i in imgs]
    off1 = (datas[0].size + 63) // 64 * 64 + 5       # unaligned start of the second stream
    blob = np.zeros(off1 + datas[1].size, dtype=np.uint8)
    blob[:datas[0].size] = datas[0]
    blob[off1:] = datas[1]
    got0 = np.concatenate([port.new_image(w, h)] * 2)
    jobs = []
    for k, (o, d) in enumerate(((0, datas[0]), (off1, datas[1]))):
        j = rs.PentaxJob()
        j.in_offset, j.in_size, j.table, j.width, j.height = o, d.size, 0, w, h
        j.out_offset, j.out_pitch = k * h * got0.shape[1] * 2, got0.shape[1] * 2
        jobs.append(j)
    plan = rs.pentax_plan(ctx, [rs.huff_table(*table)], jobs)
    got, res = gpu_run(plan, blob, got0)
    assert [r[0] for r in res] == [0, 0]
    assert np.array_equal(got[:h, :w], imgs[0]) and np.array_equal(got[h:, :w], imgs[1])
