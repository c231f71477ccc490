"""K8 parity: CUDA Phase One IIQ row decode vs the oracle (pinned against the compiled
reference in tests/test_oracle_phaseone.py), through the C ABI and the C++ host mirror."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port, synth
from helpers import gpu_run

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["3", "3w1", "3w3", "3w6", "3w8", "2"])
def p1_version(request, monkeypatch):
    """Every case with the kernel the plans take by default (the third version: group headers walked per
    row, pixels in parallel), with other forms of its walk, and with the second version (one thread per
    row); RSB200_P1 / RSB200_P1W are read when a plan is created."""
    monkeypatch.setenv("RSB200_P1", request.param[0])
    monkeypatch.setenv("RSB200_P1W", request.param[2] if len(request.param) == 3 else "0")
    return request.param


def _plan(ctx, w, h, strips, out_offset=0, first=0):
    j = rs.PhaseOneJob()
    j.out_offset, j.out_pitch, j.width, j.height, j.first_strip = out_offset, port.image_pitch(w), w, h, first
    st = []
    for off, size, row in strips:
        s = rs.PhaseOneStrip()
        s.in_offset, s.in_size, s.row = off, size, row
        st.append(s)
    return rs.phaseone_plan(ctx, [j], st)


@pytest.mark.parametrize("w,h,wild", [(8, 1, False), (70, 9, False), (258, 33, True), (1000, 40, False),
                                      (2050, 64, True)])
def test_phaseone_abi_matches_oracle(ctx, w, h, wild):
    img = synth.image_model(w, h, seed=w, wild=wild, bits=16 if wild else 14)
    blob, strips = synth.make_phaseone(img, shuffle_seed=h, gap=3)
    want = port.new_image(w, h)
    port.phaseone(want, w, blob, strips)
    got, res = gpu_run(_plan(ctx, w, h, strips), blob, port.new_image(w, h))
    assert res[0][0] == 0
    assert np.array_equal(got, want)
    assert np.array_equal(got[:, :w], img)


def test_phaseone_random_payloads_and_over_read(ctx):
    """Random bits (valid once both length prefixes at column 0 are five zeros); strips so
    short that the tail of the row is decoded from the zero padding of the pump."""
    w, h = 64, 40
    rng = np.random.default_rng(4)
    blob = rng.integers(0, 256, h * 200 + 16, dtype=np.uint8)
    strips = [(r * 200, 200 if r % 3 else 120, r) for r in range(h)]
    for off, _, _ in strips:
        blob[off + 3] = 0
        blob[off + 2] &= 0x0F
    want = port.new_image(w, h)
    try:
        port.phaseone(want, w, blob, strips)
        ok = True
    except port.RawDecoderException:
        ok = False
    plan = _plan(ctx, w, h, strips)
    if ok:
        got, res = gpu_run(plan, blob, port.new_image(w, h))
        assert res[0][0] == 0 and np.array_equal(got, want)
    else:
        import torch
        d_in = torch.from_numpy(np.concatenate([blob, np.zeros(64, np.uint8)])).cuda()
        d_out = torch.zeros(h * port.image_pitch(w) // 2, dtype=torch.int16, device="cuda")
        plan.run((d_in.data_ptr(), blob.size), d_out)
        with pytest.raises(rs.RawDecoderException):
            plan.results()


def test_phaseone_errors(ctx):
    w, h = 16, 4
    img = synth.image_model(w, h, seed=2)
    blob, strips = synth.make_phaseone(img)
    bad = blob.copy()
    bad[strips[2][0] + 3] |= 0x80
    import torch
    plan = _plan(ctx, w, h, strips)
    d_in = torch.from_numpy(np.concatenate([bad, np.zeros(64, np.uint8)])).cuda()
    d_out = torch.zeros(h * port.image_pitch(w) // 2, dtype=torch.int16, device="cuda")
    plan.run((d_in.data_ptr(), bad.size), d_out)
    with pytest.raises(rs.RawDecoderException):
        plan.results()
    with pytest.raises(rs.Rsb200Error):     # a row twice
        _plan(ctx, w, h, strips[:-1] + [strips[0]])
    # a strip far too short: the pump runs more than 8 bytes past its end
    short = [(o, 4, r) if r == 1 else (o, n, r) for o, n, r in strips]
    plan = _plan(ctx, w, h, short)
    d_in = torch.from_numpy(np.concatenate([blob, np.zeros(64, np.uint8)])).cuda()
    plan.run((d_in.data_ptr(), blob.size), d_out)
    with pytest.raises(rs.RawDecoderException):
        plan.results()
    with pytest.raises(port.RawDecoderException):
        port.phaseone(port.new_image(w, h), w, blob, short)


def test_host_phaseone_decompressor():
    w, h = 258, 33
    img = synth.image_model(w, h, seed=5, wild=True, bits=16)
    blob, strips = synth.make_phaseone(img, shuffle_seed=1, gap=5)
    want = port.new_image(w, h)
    port.phaseone(want, w, blob, strips)
    got = port.new_image(w, h)
    host.phaseone(got, w, blob, strips)
    assert np.array_equal(got, want)
    with pytest.raises(host.RawDecoderException):
        host.phaseone(port.new_image(w, h), w, blob, strips[:-1])
    bad = blob.copy()
    bad[strips[0][0] + 3] |= 0x80
    with pytest.raises(host.RawDecoderException):
        host.phaseone(port.new_image(w, h), w, bad, strips)
