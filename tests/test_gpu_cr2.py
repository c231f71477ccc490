"""CR2 slice-layout parity (Cr2Decompressor) through the C ABI, bit-exact."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from oracle import port, synth
from helpers import parse_ljpeg, gpu_run, TableSet
from test_oracle_vs_ref import CR2_CASES

pytestmark = pytest.mark.gpu


def cr2_job(blob, w, h, fmt, slicing, out_pitch, tabs):
    """What Cr2LJpegDecoder::decodeScan hands to the Cr2Decompressor ctor
    (Cr2LJpegDecoder.cpp:58-154)."""
    info = parse_ljpeg(blob)
    fw, fh = info["frame_w"], info["frame_h"]
    if info["cps"] != 3 and fw * info["cps"] > 2 * fh:
        fh *= 2
    num, sw, lsw = slicing
    if fmt == (3, 2, 1):
        sw, lsw = sw * 3 // 2, lsw * 3 // 2
    j = rs.Cr2Job()
    j.in_offset = info["data_off"]
    j.in_size = len(blob) - info["data_off"]
    j.n_comp, j.x_s_f, j.y_s_f = fmt
    for c in range(fmt[0]):
        j.table[c] = tabs.add(*info["tables"][info["table_of_comp"][c]])
        j.init_pred[c] = 1 << (info["prec"] - 1)
    j.frame_w, j.frame_h = fw, fh
    j.num_slices, j.slice_w, j.last_slice_w = num, sw, lsw
    j.img_w, j.img_h = w, h
    j.out_offset, j.out_pitch = 0, out_pitch
    return j


def _run(ctx, w, h, fmt, frame, slicing, toc, is_cfa=True, seed=31):
    img = port.new_image(w, h)
    img[:, :w] = synth.image_model(w, h, seed)
    hts = synth.default_tables(2)
    blob = port.cr2_encode(img, w, fmt, frame, slicing, 14, hts, toc, is_cfa=is_cfa)
    want = port.new_image(w, h)
    sub = (fmt[1], fmt[2])
    port.cr2_ljpeg_decode(blob, want, w, slicing, is_cfa=is_cfa, sub=sub)
    tabs = TableSet()
    job = cr2_job(blob, w, h, fmt, slicing, want.shape[1] * 2, tabs)
    plan = rs.cr2_plan(ctx, tabs.tabs, [job])
    got, res = gpu_run(plan, blob, port.new_image(w, h))
    assert res[0][0] == 0
    assert np.array_equal(got, want)
    assert np.array_equal(got[:, :w], img[:, :w])
    # consumed == Cr2Decompressor::decompress() return value
    info = parse_ljpeg(blob)
    o = port.new_image(w, h)
    fw, fh = job.frame_w, job.frame_h
    cons = port.cr2_decompress(o, w, fmt, (fw, fh), (job.num_slices, job.slice_w, job.last_slice_w),
                               [hts[t] for t in toc], [1 << 13] * fmt[0],
                               blob[info["data_off"]:], is_cfa=is_cfa)
    assert res[0][1] == cons


@pytest.mark.parametrize("case", CR2_CASES)
def test_cr2_layouts(ctx, case):
    w, h, fmt, frame, slicing = case
    _run(ctx, w, h, fmt, frame, slicing, [0, 1, 0, 1][:fmt[0]])


def test_cr2_sraw(ctx):
    _run(ctx, 96, 20, (3, 2, 1), (48, 20), (2, 48, 48), [0, 1, 1], is_cfa=False, seed=33)
    _run(ctx, 96, 20, (3, 2, 2), (32, 40), (2, 48, 48), [0, 1, 1], is_cfa=False, seed=34)


def test_cr2_mid_size_multi_cta(ctx):
    """~1 MB streams: several 64 KiB ranges per frame (speculative multi-CTA parse)."""
    _run(ctx, 1440, 960, (2, 1, 1), (720, 960), (3, 480, 480), [0, 1], seed=5)
    _run(ctx, 1440, 960, (4, 1, 1), (360, 960), (3, 480, 480), [0, 1, 0, 1], seed=6)
    _run(ctx, 1440, 960, (2, 1, 1), (720, 960), (3, 480, 480), [0, 0], seed=7)


def test_c4_cr2_6720x4480(ctx):
    """BASELINE configs[3]: Canon CR2 3-slice LJPEG 6720x4480, 2 and 4 components."""
    for fmt, frame in [((2, 1, 1), (3360, 4480)), ((4, 1, 1), (1680, 4480))]:
        _run(ctx, 6720, 4480, fmt, frame, (3, 2240, 2240), [0, 1, 0, 1][:fmt[0]], seed=4)
