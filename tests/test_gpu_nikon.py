"""Nikon NEF codec parity (streams without split): CUDA path (shared multi-CTA Huffman decode
with the plain MSB pump + K3N reconstruction with clamp and dithered curve) vs the oracle
(pinned against the compiled reference in tests/test_oracle_nikon.py), through the C ABI and
through the C++ host mirror; bit-exact."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port, synth
from helpers import gpu_run

pytestmark = pytest.mark.gpu


def _case(kind, bits, w, h, be=True, seed=1):
    half = 1 << (bits - 1)
    pup = [half, half + 2, half - 8, half - 2]
    meta = synth.nikon_meta(kind, bits, (pup[0], pup[2], pup[1], pup[3]), be)
    su = port.nikon_setup(meta, be, bits, w, h)
    img = (synth.image_model(w, h, seed=seed, bits=bits) & ((1 << bits) - 1)).astype(np.uint16)
    data = synth.make_nikon(img, su["huff_select"], pup)
    return meta, su, img, data


def _abi_decode(ctx, w, h, su, data, uncorrected):
    ncpl, values = port.nikon_tree(su["huff_select"])
    j = rs.NikonJob()
    j.in_offset, j.in_size, j.table, j.width, j.height = 0, data.size, 0, w, h
    j.out_offset, j.out_pitch = 0, port.image_pitch(w)
    j.lut = -1 if uncorrected else 0
    for k in range(4):
        j.pup[k] = su["pup"][k]
    lut = None if uncorrected else port.build_table(su["curve"], True)
    plan = rs.nikon_plan(ctx, [rs.huff_table(ncpl, values)], [j], lut)
    return gpu_run(plan, data, port.new_image(w, h))


@pytest.mark.parametrize("kind", ["lossless", "table", "segments", "z7", "skip"])
@pytest.mark.parametrize("bits", [12, 14])
@pytest.mark.parametrize("uncorrected", [False, True])
def test_nikon_abi_matches_oracle(ctx, kind, bits, uncorrected):
    w, h = 130, 37
    meta, su, img, data = _case(kind, bits, w, h, seed=bits + len(kind))
    want = port.new_image(w, h)
    port.nikon_decompress(want, w, meta, True, bits, data, uncorrected)
    got, res = _abi_decode(ctx, w, h, su, data, uncorrected)
    assert res[0][0] == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h", [(2, 1), (2, 5), (66, 2), (1026, 300), (8288, 24)])
def test_nikon_shapes_and_dither_sequence(ctx, w, h):
    meta, su, img, data = _case("table", 14, w, h, seed=w)
    want = port.new_image(w, h)
    port.nikon_decompress(want, w, meta, True, 14, data)
    got, res = _abi_decode(ctx, w, h, su, data, False)
    assert res[0][0] == 0
    assert np.array_equal(got, want)


def test_nikon_24mp_frame(ctx):
    """A D750-class 6032x4032 14-bit lossless NEF payload with curve + dither."""
    w, h = 6032, 4032
    meta, su, img, data = _case("table", 14, w, h, seed=7)
    want = port.new_image(w, h)
    port.nikon_decompress(want, w, meta, True, 14, data)
    got, res = _abi_decode(ctx, w, h, su, data, False)
    assert res[0][0] == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind,bits", [("lossless", 14), ("segments", 12)])
@pytest.mark.parametrize("uncorrected", [False, True])
def test_host_nikon_decompressor(kind, bits, uncorrected):
    w, h = 258, 41
    meta, su, img, data = _case(kind, bits, w, h, be=(bits == 14), seed=3)
    want = port.new_image(w, h)
    port.nikon_decompress(want, w, meta, bits == 14, bits, data, uncorrected)
    got = port.new_image(w, h)
    host.nikon_decompress(got, w, meta, bits == 14, bits, data, uncorrected)
    assert np.array_equal(got, want)


def test_host_nikon_error_classes():
    w, h = 64, 8
    meta, su, img, data = _case("table", 12, w, h)
    with pytest.raises(host.RawDecoderException):
        host.nikon_decompress(port.new_image(63, h), 63, meta, True, 12, data)
    with pytest.raises(host.RawDecoderException):
        host.nikon_decompress(port.new_image(w, h), w, meta, True, 13, data)
    with pytest.raises(host.IOException):
        host.nikon_decompress(port.new_image(w, h), w, meta[:9], True, 12, data)
    with pytest.raises(host.IOException):
        host.nikon_decompress(port.new_image(w, h), w, meta, True, 12, data[:40])
    # a stream with a split is refused (the second decoder is not implemented)
    sp = synth.nikon_meta("segments", 12, split=4)
    with pytest.raises(host.RawDecoderException):
        host.nikon_decompress(port.new_image(w, h), w, sp, True, 12, data)
