"""K5 parity: Canon sRaw interpolation (Cr2sRawInterpolator) on the GPU vs the oracle,
through the C ABI and through the C++ host mirror.  Integer work: bit-exact."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port
from test_oracle_sraw import sraw_input, CASES

pytestmark = pytest.mark.gpu


def job(inp, in_w, out, sub, coeffs, hue, version, in_offset=0, out_offset=0):
    per = 4 if sub == (2, 1) else 6
    j = rs.SrawJob()
    j.in_offset, j.in_pitch, j.num_mcus, j.in_rows = in_offset, inp.shape[1] * 2, in_w // per, inp.shape[0]
    j.sub_x, j.sub_y, j.version = sub[0], sub[1], version
    j.sraw_coeffs[0], j.sraw_coeffs[1], j.sraw_coeffs[2] = coeffs
    j.hue, j.out_offset, j.out_pitch = hue, out_offset, out.shape[1] * 2
    return j


def run(plan, inp, out0):
    import torch
    d_in = torch.from_numpy(inp.view(np.int16).copy()).cuda()
    d_out = torch.from_numpy(out0.view(np.int16).copy()).cuda()
    plan.run(d_in, d_out)
    torch.cuda.synchronize()
    plan.results()
    return d_out.cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("sub,version,num_mcus,rows", CASES)
@pytest.mark.parametrize("extreme", [False, True])
def test_sraw_abi(ctx, sub, version, num_mcus, rows, extreme):
    per = 4 if sub == (2, 1) else 6
    inp, in_w = sraw_input(num_mcus, rows, per, seed=version * 100 + num_mcus, extreme=extreme)
    out_w, out_h = 2 * num_mcus, rows * sub[1]
    coeffs, hue = (2100, 1024, 1700), 12 if not extreme else -400
    want = port.new_image(out_w, out_h, 3)
    got0 = want.copy()
    port.sraw_interpolate(inp, in_w, want, out_w, sub, coeffs, hue, version)
    plan = rs.sraw_plan(ctx, [job(inp, in_w, got0, sub, coeffs, hue, version)])
    assert np.array_equal(run(plan, inp, got0), want)
    b = got0.copy()
    host.sraw_interpolate(inp, in_w, b, out_w, sub, coeffs, hue, version)
    assert np.array_equal(b, want)


@pytest.mark.parametrize("sub,version", [((2, 1), 1), ((2, 2), 2)])
def test_sraw_full_frame(ctx, sub, version):
    """sRaw1-sized frame (2592x1728 output of a 5D Mk III class body)."""
    per = 4 if sub == (2, 1) else 6
    num_mcus, rows = 1296, 1728 // sub[1]
    inp, in_w = sraw_input(num_mcus, rows, per, seed=5)
    out_w, out_h = 2 * num_mcus, rows * sub[1]
    want = port.new_image(out_w, out_h, 3)
    got0 = want.copy()
    port.sraw_interpolate(inp, in_w, want, out_w, sub, (2000, 1024, 1500), 0, version)
    plan = rs.sraw_plan(ctx, [job(inp, in_w, got0, sub, (2000, 1024, 1500), 0, version)])
    assert np.array_equal(run(plan, inp, got0), want)


def test_sraw_batch_two_frames_one_plan(ctx):
    inp, in_w = sraw_input(40, 10, 6, seed=9)
    out_w, out_h = 80, 20
    want = port.new_image(out_w, out_h, 3)
    port.sraw_interpolate(inp, in_w, want, out_w, (2, 2), (1800, 1024, 1600), 3, 2)
    both_in = np.concatenate([inp, inp])
    got0 = np.concatenate([port.new_image(out_w, out_h, 3)] * 2)
    jobs = [job(inp, in_w, want, (2, 2), (1800, 1024, 1600), 3, 2, in_offset=f * inp.nbytes,
                out_offset=f * want.nbytes) for f in range(2)]
    plan = rs.sraw_plan(ctx, jobs)
    assert plan.launches == 1
    got = run(plan, both_in, got0)
    assert np.array_equal(got[:out_h], want) and np.array_equal(got[out_h:], want)


def test_sraw_rejects_bad_jobs(ctx):
    inp, in_w = sraw_input(4, 2, 4, seed=1)
    out = port.new_image(8, 2, 3)
    for sub, version in (((1, 1), 1), ((2, 2), 0), ((2, 1), 3)):
        with pytest.raises(rs.Rsb200Error):
            rs.sraw_plan(ctx, [job(inp, in_w, out, sub, (1, 1, 1), 0, version)])
