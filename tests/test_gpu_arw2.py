"""K6 parity: CUDA Sony ARW2 decode vs the oracle (pinned against the compiled reference in
tests/test_oracle_arw2.py), through the C ABI and through the C++ host mirror; bit-exact,
including the dithered curve (the reference's only serial dependency, jumped over with
modular arithmetic on the device)."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port, synth
from helpers import gpu_run

pytestmark = pytest.mark.gpu


def _job(w, h, in_offset=0, out_offset=0, table=-1):
    j = rs.Arw2Job()
    j.in_offset, j.out_offset = in_offset, out_offset
    j.out_pitch, j.width, j.height, j.table = port.image_pitch(w), w, h, table
    return j


@pytest.mark.parametrize("w,h", [(32, 1), (64, 5), (320, 33), (9600, 3), (4000 // 32 * 32, 300)])
@pytest.mark.parametrize("table", ["none", "plain", "dither"])
def test_arw2_matches_oracle(ctx, w, h, table):
    data = synth.arw2_frame(w, h, seed=w + 3 * h)
    want = port.new_image(w, h)
    t = None if table == "none" else port.build_table(synth.sony_curve(), table == "dither")
    port.sony_arw2(want, w, data, t, table == "dither")
    plan = rs.arw2_plan(ctx, [_job(w, h, table=-1 if t is None else 0)], t, table == "dither")
    got, res = gpu_run(plan, data, port.new_image(w, h))
    assert res[0][0] == 0
    assert np.array_equal(got, want)


def test_arw2_unaligned_input_and_batch(ctx):
    """Three images in one plan at odd byte offsets (file offsets are arbitrary)."""
    shapes = [(64, 7), (320, 9), (96, 4)]
    blobs, jobs, wants = [], [], []
    pos, opos = 5, 0
    t = port.build_table(synth.sony_curve(), True)
    for k, (w, h) in enumerate(shapes):
        d = synth.arw2_frame(w, h, seed=70 + k)
        jobs.append(_job(w, h, in_offset=pos, out_offset=opos, table=0))
        blobs.append((pos, d))
        want = port.new_image(w, h)
        port.sony_arw2(want, w, d, t, True)
        wants.append((opos, want))
        pos += d.size + 3
        opos += (want.size * 2 + 255) // 256 * 256
    buf = np.zeros(pos + 16, dtype=np.uint8)
    for o, d in blobs:
        buf[o:o + d.size] = d
    plan = rs.arw2_plan(ctx, jobs, t, True)
    out = np.zeros(opos // 2, dtype=np.uint16)
    got, res = gpu_run(plan, buf, out)
    assert all(s == 0 for s, _ in res)
    for o, want in wants:
        assert np.array_equal(got.reshape(-1)[o // 2:o // 2 + want.size].reshape(want.shape), want)


def test_arw2_invalid_block_is_rde(ctx):
    w, h = 64, 4
    data = synth.arw2_frame(w, h, seed=3).copy()
    blk = data[2 * w + 16:2 * w + 32]
    v = int(blk[2]) | (int(blk[3]) << 8)
    v = (v & ~(15 << 10)) | (((v >> 6) & 15) << 10)   # imin := imax
    blk[2], blk[3] = v & 255, v >> 8
    plan = rs.arw2_plan(ctx, [_job(w, h)])
    import torch
    d_in = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).cuda()
    d_out = torch.zeros(h * port.image_pitch(w) // 2, dtype=torch.int16, device="cuda")
    plan.run((d_in.data_ptr(), data.size), d_out)
    with pytest.raises(rs.RawDecoderException):
        plan.results()
    with pytest.raises(port.RawDecoderException):
        port.sony_arw2(port.new_image(w, h), w, data)


@pytest.mark.parametrize("table", ["none", "dither"])
def test_host_sony_arw2_decompressor(table):
    """The C++ mirror of the reference class, driven like ArwDecoder drives it."""
    w, h = 640, 37
    data = synth.arw2_frame(w, h, seed=11)
    curve = synth.sony_curve()
    want = port.new_image(w, h)
    port.sony_arw2(want, w, data, None if table == "none" else port.build_table(curve, True), True)
    got = port.new_image(w, h)
    host.sony_arw2(got, w, data, None if table == "none" else curve, True)
    assert np.array_equal(got, want)
    with pytest.raises(host.IOException):
        host.sony_arw2(port.new_image(w, h), w, data[:-1])
    with pytest.raises(host.RawDecoderException):
        host.sony_arw2(port.new_image(48, 2), 48, data)


def test_c_full_frame_arw2(ctx):
    """A 6000x4000 ARW2 frame (24 MP) with the dithered curve, bit-exact."""
    w, h = 6016, 4000
    data = synth.arw2_frame(w, h, seed=5)
    t = port.build_table(synth.sony_curve(), True)
    want = port.new_image(w, h)
    port.sony_arw2(want, w, data, t, True)
    plan = rs.arw2_plan(ctx, [_job(w, h, table=0)], t, True)
    got, res = gpu_run(plan, data, port.new_image(w, h))
    assert res[0][0] == 0 and np.array_equal(got, want)
