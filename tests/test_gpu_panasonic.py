"""K7 parity: CUDA Panasonic V5/V6/V7 decode vs the oracle (pinned against the compiled
reference in tests/test_oracle_panasonic.py), through the C ABI and the C++ host mirror;
bit-exact on random payloads (every bit pattern is a valid stream for these codecs)."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port
from helpers import gpu_run
from test_oracle_panasonic import payload, CASES

pytestmark = pytest.mark.gpu


def _job(version, bps, w, h, size, in_offset=0, out_offset=0):
    j = rs.PanaJob()
    j.in_offset, j.in_size, j.out_offset = in_offset, size, out_offset
    j.out_pitch, j.width, j.height, j.version, j.bps = port.image_pitch(w), w, h, version, bps
    return j


@pytest.mark.parametrize("version,bps,w,h", CASES + [(5, 14, 5184 // 9 * 9, 400), (6, 12, 4200, 300),
                                                     (7, 14, 6003, 200)])
def test_panasonic_abi_matches_oracle(ctx, version, bps, w, h):
    data = payload(version, w, h, bps, seed=version * 7 + w)
    want = port.new_image(w, h)
    port.panasonic(version, want, w, data, bps)
    plan = rs.pana_plan(ctx, [_job(version, bps, w, h, data.size)])
    got, res = gpu_run(plan, data, port.new_image(w, h))
    assert np.array_equal(got, want)


def test_panasonic_v6_special_values(ctx):
    for bps, ppb in ((12, 14), (14, 11)):
        w, h = ppb * 4, 2
        for fill in (0x00, 0xFF, 0x0F, 0xF0):
            data = np.full(w * h // ppb * 16, fill, dtype=np.uint8)
            want = port.new_image(w, h)
            port.panasonic(6, want, w, data, bps)
            plan = rs.pana_plan(ctx, [_job(6, bps, w, h, data.size)])
            got, _ = gpu_run(plan, data, port.new_image(w, h))
            assert np.array_equal(got, want)


def test_panasonic_batch_mixed_versions_unaligned(ctx):
    """Five images of different versions in one plan, inputs at odd byte offsets."""
    specs = [(5, 12, 40, 30), (6, 14, 220, 9), (7, 14, 180, 7), (5, 14, 90, 11), (6, 12, 280, 5)]
    jobs, blobs, wants = [], [], []
    pos, opos = 3, 0
    for k, (v, bps, w, h) in enumerate(specs):
        d = payload(v, w, h, bps, 50 + k)
        jobs.append(_job(v, bps, w, h, d.size, in_offset=pos, out_offset=opos))
        blobs.append((pos, d))
        want = port.new_image(w, h)
        port.panasonic(v, want, w, d, bps)
        wants.append((opos, want))
        pos += d.size + 5
        opos += (want.size * 2 + 255) // 256 * 256
    buf = np.zeros(pos + 16, dtype=np.uint8)
    for o, d in blobs:
        buf[o:o + d.size] = d
    plan = rs.pana_plan(ctx, jobs)
    got, _ = gpu_run(plan, buf, np.zeros(opos // 2, dtype=np.uint16))
    for (o, want), (v, bps, w, h) in zip(wants, specs):
        g = got.reshape(-1)[o // 2:o // 2 + want.size].reshape(want.shape)
        assert np.array_equal(g[:, :w], want[:, :w])


@pytest.mark.parametrize("version,bps,w,h", [(5, 12, 400, 33), (6, 14, 1100, 13), (7, 14, 1809, 10)])
def test_host_panasonic_decompressors(version, bps, w, h):
    data = payload(version, w, h, bps, seed=9)
    want = port.new_image(w, h)
    port.panasonic(version, want, w, data, bps)
    got = port.new_image(w, h)
    host.panasonic(version, got, w, data, bps)
    assert np.array_equal(got, want)
    with pytest.raises(host.RawDecoderException):   # truncated
        host.panasonic(version, port.new_image(w, h), w, data[:-1], bps)
    with pytest.raises(host.RawDecoderException):   # width not a multiple of the unit
        host.panasonic(version, port.new_image(w + 1, h), w + 1, data, bps)
