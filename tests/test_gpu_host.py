"""Drop-in parity: the C++ host mirror of the reference classes
(UncompressedDecompressor, LJpegDecoder, AbstractDngDecompressor,
Cr2LJpegDecoder, ...) driven exactly like the reference's callers, compared
byte-for-byte over the uncropped buffer with the oracle.  Everything below the
class interface runs on the GPU through the C ABI."""
import numpy as np
import pytest

import rawspeed_b200 as rs
from rawspeed_b200 import host
from oracle import port, synth
from helpers import parse_ljpeg
from test_oracle_vs_ref import CR2_CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [port.LSB, port.MSB, port.MSB16, port.MSB32])
@pytest.mark.parametrize("bps", [10, 12, 14, 16, 7])
def test_uncompressed_decompressor(order, bps):
    w, h = 264, 9
    data, pitch = synth.packed_frame(w, h, bps, seed=bps, pitch=w * bps // 8 + 4)
    a = port.new_image(w, h + 3)
    b = a.copy()
    port.unpack(data, a, w, 1, (0, 2, w, h), pitch, bps, order)
    host.unpack(data, b, w, 1, (0, 2, w, h), pitch, bps, order)
    assert np.array_equal(a, b)


def _dng(img, tw, th, cpp=1, **kw):
    h, wc = img.shape
    w = wc // cpp
    fix = kw.get("fix16", False)
    t = synth.make_dng_ljpeg(img, tw, th, cpp=cpp, **kw)
    a = port.new_image(w, h, cpp)
    b = a.copy()
    port.dng_decompress(t.blob, t.offsets, t.lengths, a, w, cpp, tw, th, 7, fix_ljpeg=fix)
    host.dng_decompress(t.blob, t.offsets, t.lengths, b, w, cpp, tw, th, 7, fix_ljpeg=fix)
    assert np.array_equal(a, b)
    assert np.array_equal(b[:, :wc], img)
    return t


@pytest.mark.parametrize("ljpeg_path", ["auto", "thread"])
def test_abstract_dng_decompressor_ljpeg(ljpeg_path, monkeypatch):
    if ljpeg_path != "auto":
        monkeypatch.setenv("RSB200_LJPEG_PATH", ljpeg_path)
    _dng(synth.image_model(300, 200, 7), 128, 64)
    _dng(synth.image_model(256, 96, 9, wild=True), 128, 32)
    img16 = synth.image_model(128, 64, 11, wild=True, bits=16)
    img16[0, 0:8] = [0, 0x8000, 0, 0x8000, 0xFFFF, 0x7FFF, 0, 0x8000]
    _dng(img16, 64, 64, prec=16)
    _dng(img16, 64, 64, prec=16, fix16=True)
    _dng(synth.image_model(96, 48, 13), 48, 24, ncomp=4, mcu=(2, 2))
    _dng(synth.image_model(96 * 3, 40, 14), 32, 20, ncomp=3, cpp=3)
    _dng(synth.image_model(200, 100, 15), 100, 50, tabs=synth.default_tables(2), tab_of_comp=[0, 1])
    _dng(synth.image_model(160, 96, 17), 80, 48, restart_rows=1)
    _dng(synth.image_model(160, 96, 17), 80, 48, restart_rows=5)
    _dng(synth.image_model(101, 33, 19), 64, 16)


def test_abstract_dng_decompressor_uncompressed():
    W, H, tw, th = 96, 60, 96, 16
    for bps, be in [(12, False), (14, False), (16, False), (16, True), (8, True)]:
        pitch = tw * bps // 8
        ntiles = 4
        blob = synth.lcg_bytes(pitch * th * ntiles + 64, bps)
        offs = [7 + n * pitch * th for n in range(ntiles)]
        a = port.new_image(W, H)
        b = a.copy()
        port.dng_decompress(blob, offs, [pitch * th] * ntiles, a, W, 1, tw, th, 1, bps=bps, big_endian=be)
        host.dng_decompress(blob, offs, [pitch * th] * ntiles, b, W, 1, tw, th, 1, bps=bps, big_endian=be)
        assert np.array_equal(a, b), (bps, be)


@pytest.mark.parametrize("ljpeg_path", ["auto", "thread"])
def test_ljpeg_decoder_single_tile_and_consumed(ljpeg_path, monkeypatch):
    if ljpeg_path != "auto":
        monkeypatch.setenv("RSB200_LJPEG_PATH", ljpeg_path)
    img = synth.image_model(64, 40, 3)
    hts = synth.default_tables(1)
    for rr in (0, 1, 7):
        blob = port.ljpeg_encode(img, 32, 40, (2, 1), 14, hts, [0, 0], rr)
        a = port.new_image(64, 40)
        b = a.copy()
        port.ljpeg_decode(blob, a, 64, 1, (0, 0), (64, 40), (64, 40))
        host.ljpeg_decode(blob, b, 64, 1, (0, 0), (64, 40), (64, 40))
        assert np.array_equal(a, b)
        info = parse_ljpeg(blob)
        data = blob[info["data_off"]:]
        rpr = rr if rr else 40
        ca = port.ljpeg_decompress(port.new_image(64, 40), 64, 1, (0, 0, 64, 40), (2, 1), (32, 40),
                                   [hts[0]] * 2, [1 << 13] * 2, rpr, data)
        cb = host.ljpeg_decompress(port.new_image(64, 40), 64, 1, (0, 0, 64, 40), (2, 1), (32, 40),
                                   [(hts[0].ncpl, hts[0].values)], [0, 0], [1 << 13] * 2, rpr, data)
        assert ca == cb
        if rr == 0:
            # garbage between the scan and EOI: position follows the reference's refill cadence
            for pad in (1, 3, 4, 7, 11):
                data2 = np.concatenate([data[:-2], np.zeros(pad, np.uint8), data[-2:]])
                ca = port.ljpeg_decompress(port.new_image(64, 40), 64, 1, (0, 0, 64, 40), (2, 1),
                                           (32, 40), [hts[0]] * 2, [1 << 13] * 2, rpr, data2)
                cb = host.ljpeg_decompress(port.new_image(64, 40), 64, 1, (0, 0, 64, 40), (2, 1),
                                           (32, 40), [(hts[0].ncpl, hts[0].values)], [0, 0],
                                           [1 << 13] * 2, rpr, data2)
                assert ca == cb, pad


@pytest.mark.parametrize("ljpeg_path", ["auto", "thread"])
def test_stream_errors_same_class(ljpeg_path, monkeypatch):
    if ljpeg_path != "auto":
        monkeypatch.setenv("RSB200_LJPEG_PATH", ljpeg_path)
    img = synth.image_model(64, 32, 23, wild=True)
    t = synth.make_dng_ljpeg(img, 64, 32)
    info = parse_ljpeg(t.blob)
    off = info["data_off"]
    cases = []
    b = t.blob.copy(); b[off + 40:off + 49] = [0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFE]; cases.append(b)  # bad code
    cases.append(t.blob[:off + 300].copy())                                                              # truncated
    b = t.blob.copy(); b[off + 200:off + 202] = [0xFF, 0xD9]; cases.append(b)                            # early EOI
    for blob in cases:
        with pytest.raises(port.OracleError) as eo:
            port.dng_decompress(blob, [0], [len(blob)], port.new_image(64, 32), 64, 1, 64, 32, 7)
        with pytest.raises(rs.Rsb200Error) as eh:
            host.dng_decompress(blob, [0], [len(blob)], port.new_image(64, 32), 64, 1, 64, 32, 7)
        assert isinstance(eh.value, rs.RawDecoderException)  # decompress() rethrows as RDE


@pytest.mark.parametrize("case", CR2_CASES)
def test_cr2_ljpeg_decoder(case):
    w, h, fmt, frame, slicing = case
    img = port.new_image(w, h)
    img[:, :w] = synth.image_model(w, h, 31)
    hts = synth.default_tables(2)
    blob = port.cr2_encode(img, w, fmt, frame, slicing, 14, hts, [0, 1, 0, 1][:fmt[0]])
    a = port.new_image(w, h)
    b = a.copy()
    port.cr2_ljpeg_decode(blob, a, w, slicing)
    host.cr2_ljpeg_decode(blob, b, w, slicing)
    assert np.array_equal(a, b)


def test_cr2_sraw():
    hts = synth.default_tables(2)
    for fmt, w, h, frame, slicing in [((3, 2, 1), 96, 20, (48, 20), (2, 48, 48)),
                                      ((3, 2, 2), 96, 20, (32, 40), (2, 48, 48))]:
        img = port.new_image(w, h)
        img[:, :w] = synth.image_model(w, h, 33)
        blob = port.cr2_encode(img, w, fmt, frame, slicing, 14, hts, [0, 1, 1], is_cfa=False)
        a = port.new_image(w, h)
        b = a.copy()
        port.cr2_ljpeg_decode(blob, a, w, slicing, is_cfa=False, sub=(fmt[1], fmt[2]))
        host.cr2_ljpeg_decode(blob, b, w, slicing, is_cfa=False, sub=(fmt[1], fmt[2]))
        assert np.array_equal(a, b)


# ---- the other UncompressedDecompressor members (fixed layouts, F32 images) ----
def _form_both(data, mk, w, cpp, crop, pitch, bps, order, form, curve=None, dither=False):
    a, b = mk(), mk()
    table = port.build_table(curve, dither) if curve is not None else None
    ea = eb = None
    try:
        port.unpack_form(data, a, w, cpp, crop, pitch, bps, order, form, table, dither)
    except port.OracleError as e:
        ea = e
    try:
        host.unpack_form(data, b, w, cpp, crop, pitch, bps, order, form, curve, dither)
    except rs.Rsb200Error as e:
        eb = e
    if ea is None:
        assert eb is None, eb
        assert np.array_equal(a, b)
    else:
        want = rs.IOException if isinstance(ea, port.IOException) else rs.RawDecoderException
        assert type(eb) is want, (ea, eb)
        assert ea.msg[:30] in str(eb)
    return ea


@pytest.mark.parametrize("form", [port.FORM_8BIT, port.FORM_8BIT_UNCORRECTED])
@pytest.mark.parametrize("curve_kind", ["none", "plain", "dither", "short"])
def test_host_decode8bit(form, curve_kind):
    w, h = 70, 9
    data = synth.lcg_bytes(w * h + 5, seed=3)
    curve, dither = None, False
    if curve_kind != "none":
        n = 256 if curve_kind != "short" else 100
        curve = (np.arange(n, dtype=np.uint32) ** 2 // 2 % 65536).astype(np.uint16)
        curve[n // 2] = 3
        dither = curve_kind == "dither"
    assert _form_both(data, lambda: port.new_image(w, h), w, 1, (0, 0, w, h), w, 8, port.LSB,
                      form, curve, dither) is None


@pytest.mark.parametrize("form", [port.FORM_12BIT_CONTROL_BE, port.FORM_12BIT_CONTROL_LE,
                                  port.FORM_12BIT_LEFT_BE, port.FORM_12BIT_LEFT_LE])
def test_host_decode12_forms(form):
    for w in (10, 38, 64, 250):
        h = 5
        control = form in (port.FORM_12BIT_CONTROL_BE, port.FORM_12BIT_CONTROL_LE)
        perline = 12 * w // 8 + (w + 2) // 10 if control else 2 * w
        data = synth.lcg_bytes(perline * h, seed=w)
        # the reference's callers construct with the real line pitch
        assert _form_both(data, lambda: port.new_image(w, h), w, 1, (0, 0, w, h), perline,
                          12 if control else 16, port.MSB if control else port.LSB, form) is None


@pytest.mark.parametrize("order", [port.LSB, port.MSB])
@pytest.mark.parametrize("bps", [16, 24, 32])
def test_host_float_image(order, bps):
    for cpp in (1, 3):
        w, h, ox, oy = 24, 6, 4, 1
        pitch = w * cpp * bps // 8 + 4
        data = np.random.default_rng(bps + cpp).integers(0, 256, pitch * h, dtype=np.uint8)
        assert _form_both(data, lambda: port.new_image_f32(w + 8, h + 2, cpp), w + 8, cpp,
                          (ox, oy, w, h), pitch, bps, order, port.FORM_READ) is None


def test_host_forms_error_classes():
    w, h = 20, 6
    data = synth.lcg_bytes(w * h, seed=5)
    # member needs more bytes than the stream the constructor accepted -> IOException
    e = _form_both(data, lambda: port.new_image(w, h), w, 1, (0, 0, w, h), w, 8, port.LSB,
                   port.FORM_12BIT_LEFT_LE)
    assert isinstance(e, port.IOException)
    e = _form_both(data, lambda: port.new_image(7, 2), 7, 1, (0, 0, 7, 2), 7, 8, port.LSB,
                   port.FORM_12BIT_CONTROL_LE)   # 12*7 % 8 != 0
    assert isinstance(e, port.IOException)
    e = _form_both(data, lambda: port.new_image_f32(8, 2), 8, 1, (0, 0, 8, 2), 16, 16,
                   port.MSB16, port.FORM_READ)   # unsupported float packing
    assert isinstance(e, port.RawDecoderException)


@pytest.mark.parametrize("bps", [16, 24, 32])
def test_host_float_dng_tiles(bps):
    """Floating-point DNG (compression 1) through AbstractDngDecompressor on an F32 image."""
    W, H, tw, th = 100, 60, 32, 16
    pitch = tw * bps // 8
    blob = synth.lcg_bytes(pitch * th * 16 + 64, bps)
    offs = [7 + n * pitch * th for n in range(16)]
    lens = [pitch * th] * 16
    for be in (False, True):
        a, b = port.new_image_f32(W, H), port.new_image_f32(W, H)
        port.dng_decompress(blob, offs, lens, a, W, 1, tw, th, 1, bps=bps, big_endian=be)
        host.dng_decompress(blob, offs, lens, b, W, 1, tw, th, 1, bps=bps, big_endian=be)
        assert np.array_equal(a, b), (bps, be)
