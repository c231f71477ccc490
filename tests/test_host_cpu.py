"""Host-layer logic that needs no GPU: the C++ mirror validates inputs exactly
like the reference (same exception class) before anything reaches the device."""
import numpy as np
import pytest

from oracle import port, synth
from rawspeed_b200 import host
import rawspeed_b200 as rs


def test_host_library_exports():
    L = host.lib()
    for n in host.EXPORTS:
        assert hasattr(L, n), n


def test_huffman_validation_matches_oracle():
    rng = np.random.default_rng(5)
    for trial in range(300):
        ncpl = [0] * 16
        for _ in range(int(rng.integers(1, 6))):
            ncpl[int(rng.integers(0, 16))] += int(rng.integers(1, 4))
        values = [int(v) for v in rng.integers(0, 18, sum(ncpl))]
        ok_a = ok_b = True
        try:
            port.Huff(ncpl, values)
        except port.OracleError:
            ok_a = False
        try:
            host.huff_check(ncpl, values)
        except rs.Rsb200Error:
            ok_b = False
        assert ok_a == ok_b, (ncpl, values)


def _same_class(fn_oracle, fn_host):
    with pytest.raises(port.OracleError) as eo:
        fn_oracle()
    with pytest.raises(rs.Rsb200Error) as eh:
        fn_host()
    want = rs.IOException if isinstance(eo.value, port.IOException) else rs.RawDecoderException
    assert type(eh.value) is want, (eo.value, eh.value)


def test_unpack_ctor_errors_same_class():
    w, h, bps = 16, 4, 12
    data = synth.lcg_bytes(24 * 4, 1)
    cases = [dict(crop=(0, 0, w, h), pitch=23), dict(crop=(0, 0, w, 5), pitch=24),
             dict(crop=(1, 0, w, h), pitch=24), dict(crop=(0, 9, w, h), pitch=24),
             dict(crop=(0, 0, 15, h), pitch=24, bps=13), dict(crop=(0, 0, 0, h), pitch=24),
             dict(crop=(0, 0, w, h), pitch=24, bps=17), dict(crop=(0, 0, w, h), pitch=24, order=4)]
    for c in cases:
        b, o = c.get("bps", bps), c.get("order", port.MSB)
        _same_class(lambda: port.unpack(data, port.new_image(w, h), w, 1, c["crop"], c["pitch"], b, o),
                    lambda: host.unpack(data, port.new_image(w, h), w, 1, c["crop"], c["pitch"], b, o))


def test_ljpeg_header_errors_same_class():
    img = synth.image_model(64, 32, 3)
    good = port.ljpeg_encode(img, 32, 32, (2, 1), 14, synth.default_tables(1), [0, 0])
    bad = []
    b = good.copy(); b[1] = 0xD9; bad.append(b)                      # no SOI
    b = good.copy(); b[3] = 0xC0; bad.append(b)                      # SOF0 instead of SOF3 -> no SOF
    b = good.copy(); b[6] = 1; bad.append(b)                         # precision 1
    b = good.copy(); b[6] = 17; bad.append(b)                        # precision 17
    b = good.copy(); b[11] = 5; bad.append(b)                        # 5 components
    bad.append(good[:20].copy())                                      # truncated header
    b = good.copy(); b[9:11] = [0, 16]; bad.append(b)                # frame width mismatch
    for i, blob in enumerate(bad):
        _same_class(lambda: port.ljpeg_decode(blob, port.new_image(64, 32), 64, 1, (0, 0), (64, 32), (64, 32)),
                    lambda: host.ljpeg_decode(blob, port.new_image(64, 32), 64, 1, (0, 0), (64, 32), (64, 32)))
    # tile geometry errors are raised before the stream is looked at
    for off, size, mx in [((64, 0), (64, 32), (64, 32)), ((0, 0), (65, 32), (64, 32)),
                          ((0, 0), (64, 32), (32, 32)), ((0, 8), (64, 32), (64, 32))]:
        _same_class(lambda: port.ljpeg_decode(good, port.new_image(64, 32), 64, 1, off, size, mx),
                    lambda: host.ljpeg_decode(good, port.new_image(64, 32), 64, 1, off, size, mx))


def test_cr2_ctor_errors_same_class():
    w, h = 64, 40
    img = port.new_image(w, h)
    img[:, :w] = synth.image_model(w, h, 31)
    hts = synth.default_tables(2)
    blob = port.cr2_encode(img, w, (2, 1, 1), (32, 40), (2, 32, 32), 14, hts, [0, 1])
    for slicing in [(2, 31, 33), (2, 32, 16), (2, 32, 48), (1, 0, 62)]:
        _same_class(lambda: port.cr2_ljpeg_decode(blob, port.new_image(w, h), w, slicing),
                    lambda: host.cr2_ljpeg_decode(blob, port.new_image(w, h), w, slicing))


def test_vendor_codec_ctor_errors_same_class():
    """Constructors of the vendor codec mirrors (Nikon, Sony ARW2, Panasonic V5/V6/V7, Phase
    One) reject what the reference rejects, with the same exception class, before any GPU work."""
    # Nikon: odd width, bad bit depth, truncated maker note, bad curve segment count
    w, h = 64, 8
    meta = synth.nikon_meta("table", 12)
    data = synth.lcg_bytes(256, 3)
    for args in [(63, h, meta, 12), (w, h, meta, 13), (w, h, meta[:9], 12)]:
        ww, hh, m, bits = args
        _same_class(lambda: port.nikon_decompress(port.new_image(ww, hh), ww, m, True, bits, data),
                    lambda: host.nikon_decompress(port.new_image(ww, hh), ww, m, True, bits, data))
    bad = bytearray(synth.nikon_meta("segments", 12))
    bad[10:12] = bytes([0, 30])
    _same_class(lambda: port.nikon_decompress(port.new_image(w, h), w, bytes(bad), True, 12, data),
                lambda: host.nikon_decompress(port.new_image(w, h), w, bytes(bad), True, 12, data))
    # Sony ARW2: width not a multiple of 32, not enough data
    a = synth.arw2_frame(64, 4, seed=1)
    _same_class(lambda: port.sony_arw2(port.new_image(48, 2), 48, a),
                lambda: host.sony_arw2(port.new_image(48, 2), 48, a))
    _same_class(lambda: port.sony_arw2(port.new_image(64, 4), 64, a[:-1]),
                lambda: host.sony_arw2(port.new_image(64, 4), 64, a[:-1]))
    # Panasonic: width not a multiple of the unit, unsupported bps, too few blocks
    blob = synth.lcg_bytes(0x8000, 1)
    for ver, bps, ww in [(5, 12, 41), (5, 13, 40), (6, 12, 27), (6, 16, 28), (7, 14, 20)]:
        _same_class(lambda: port.panasonic(ver, port.new_image(ww, 2), ww, blob, bps),
                    lambda: host.panasonic(ver, port.new_image(ww, 2), ww, blob, bps))
    _same_class(lambda: port.panasonic(7, port.new_image(18, 2), 18, blob[:63], 14),
                lambda: host.panasonic(7, port.new_image(18, 2), 18, blob[:63], 14))
    # Phase One: strip count, a row twice, odd width
    img = synth.image_model(16, 4, seed=2)
    pb, strips = synth.make_phaseone(img)
    for ww, st in [(16, strips[:-1]), (16, strips[:-1] + [strips[0]]), (15, strips)]:
        _same_class(lambda: port.phaseone(port.new_image(ww, 4), ww, pb, st),
                    lambda: host.phaseone(port.new_image(ww, 4), ww, pb, st))


# ---- RawImageData::scaleBlackWhite: the host half (estimate + calculateBlackAreas) -----------

def _sensor(w, h, seed, black=512, white=15000):
    rng = np.random.default_rng(seed)
    img = port.new_image(w, h)
    img[:, :] = rng.integers(black, white, size=img.shape, dtype=np.uint16)
    img[:, :16] = (black + rng.integers(-6, 7, size=(h, 16))).astype(np.uint16)
    img[:8, :] = (black + 3 + rng.integers(-6, 7, size=(8, img.shape[1]))).astype(np.uint16)
    return img


@pytest.mark.parametrize("name,w,h,crop,kw", [
    ("vertical_area", 96, 40, (16, 8, 80, 32), dict(white=15000, areas=[(1, 0, 16)])),
    ("horizontal_area", 96, 40, (16, 8, 80, 32), dict(white=15000, areas=[(0, 0, 8)])),
    ("both_odd_sizes", 97, 41, (17, 9, 80, 32), dict(white=15000, areas=[(1, 1, 15), (0, 1, 7)])),
    ("not_cfa_average", 96, 40, (16, 8, 80, 32), dict(white=15000, areas=[(1, 0, 16)], is_cfa=False)),
    ("black_level_only", 64, 24, (0, 0, 64, 24), dict(black_level=500, white=15000)),
    ("nothing_to_do", 64, 24, (0, 0, 64, 24), dict(black_level=0, white=65535)),
    ("estimate_both", 640, 560, (4, 4, 620, 540), dict()),
    ("estimate_white", 640, 560, (4, 4, 620, 540), dict(black_level=600)),
])
def test_scale_black_white_host_half_matches_oracle(name, w, h, crop, kw):
    a = _sensor(w, h, len(name))
    keep = a.copy()
    got = host.scale_black_white(a, w, crop, host_part_only=True, **kw)
    assert np.array_equal(a, keep)          # the host half does not touch the pixels
    b = keep.copy()
    want = port.scale_black_white(b, w, crop, **kw)
    assert got == want


def test_scale_black_white_area_errors_same_class():
    a = _sensor(64, 24, 5)
    for areas in ([(0, 20, 8)], [(1, 60, 8)]):
        kw = dict(white=15000, areas=areas)
        _same_class(lambda: port.scale_black_white(a.copy(), 64, (0, 0, 64, 24), **kw),
                    lambda: host.scale_black_white(a.copy(), 64, (0, 0, 64, 24), host_part_only=True, **kw))


def test_scale_black_white_needs_the_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the device pass runs (tests/test_gpu_scale.py)")
    a = _sensor(64, 24, 6)
    with pytest.raises(Exception):
        host.scale_black_white(a, 64, (0, 0, 64, 24), black_level=500, white=15000)


def test_panasonic_v4_ctor_errors_same_class():
    data = synth.lcg_bytes(0x8000, 2)
    for args in ((port.new_image(15, 2), 15, data),                     # width % 14
                 (port.new_image(14, 2), 14, data, True, 0x4001),       # split beyond the block
                 (port.new_image(1400, 40), 1400, data),                # not enough data
                 (port.new_image(1400, 40), 1400, data, True, 0x2008)):
        _same_class(lambda: port.panasonic_v4(*args),
                    lambda: host.panasonic_v4(*args, construct_only=True))
    # a well-formed descriptor constructs (nothing runs)
    host.panasonic_v4(port.new_image(28, 2), 28, data, True, 0x2008, construct_only=True)


# ---- AbstractDngDecompressor::prepareLJpeg (the host half of the DNG LJPEG path) ---------------

def _dng(w, h, tile, **kw):
    img = synth.image_model(w, h, 3)
    t = synth.make_dng_ljpeg(img, tile, tile, **kw)
    return t, [int(o) for o in t.offsets], [int(n) for n in t.lengths]


@pytest.mark.parametrize("kw", [dict(), dict(restart_rows=1), dict(restart_rows=4),
                                dict(tabs=synth.default_tables(2), tab_of_comp=[0, 1])])
def test_prepare_ljpeg_is_the_same_on_any_number_of_threads(kw):
    w, h, tile = 1024, 768, 64
    t, offs, lens = _dng(w, h, tile, **kw)
    runs = [host.dng_ljpeg_host_half(t.blob, offs, lens, w, h, 1, tile, tile, False, th, 1) for th in (1, 2, 5, 16, 0)]
    assert len({r["digest"] for r in runs}) == 1
    assert runs[0]["errors"] == 0 and runs[0]["tables"] == len(kw.get("tabs", [0]))
    rows = kw.get("restart_rows", 0)
    per_tile = 1 if not rows else -(-tile // rows)
    assert runs[0]["scans"] == len(offs) * per_tile


def test_prepare_ljpeg_reports_tile_errors_in_tile_order_on_any_number_of_threads():
    w, h, tile = 1024, 768, 64
    t, offs, lens = _dng(w, h, tile, restart_rows=2)
    blob = t.blob.copy()
    rng = np.random.default_rng(4)
    bad = sorted(rng.choice(len(offs), 9, replace=False).tolist())
    for k, i in enumerate(bad):
        o = offs[i]
        if k % 3 == 0:
            blob[o + 3] = 0x00                     # SOF marker destroyed
        elif k % 3 == 1:
            blob[o + lens[i] // 2:o + lens[i]] = 0  # restart markers gone: "Jpeg marker not encountered"
        else:
            lens[i] = 20                            # tile cut inside its headers
    runs = [host.dng_ljpeg_host_half(blob, offs, lens, w, h, 1, tile, tile, False, th, 1) for th in (1, 3, 16)]
    assert len({r["digest"] for r in runs}) == 1    # the digest covers the error texts, in order
    assert runs[0]["errors"] == len(bad)
    assert runs[0]["scans"] == (len(offs) - len(bad)) * (tile // 2)


@pytest.mark.parametrize("kw", [dict(), dict(restart_rows=2), dict(tabs=synth.default_tables(2), tab_of_comp=[0, 1]),
                                dict(ncomp=4, restart_rows=8)])
def test_prepare_ljpeg_descriptors_equal_the_test_suite_s_own_scan_builder(kw):
    """The C++ host half against tests/helpers.dng_ljpeg_scans (the Python builder every GPU
    parity test of the C ABI feeds the device with): same scans, field by field."""
    from helpers import dng_ljpeg_scans
    w, h, tile = 520, 300, 64            # ragged right and bottom tiles
    t, offs, lens = _dng(w, h, tile, **kw)
    pitch = port.image_pitch(w)
    tabs, want = dng_ljpeg_scans(t, pitch)
    got = host.dng_ljpeg_host_half(t.blob, offs, lens, w, h, 1, tile, tile, False, 3, 1, want_scans=True)
    assert got["errors"] == 0 and got["scans"] == len(want) and got["tables"] == len(tabs.tabs)
    for a, b in zip(got["scan_list"], want):
        ncomp = b.mcu_w * b.mcu_h
        assert (a.in_offset, a.in_size, a.rows, a.frame_w, a.mcu_w, a.mcu_h) == \
               (b.in_offset, b.in_size, b.rows, b.frame_w, b.mcu_w, b.mcu_h)
        assert list(a.table)[:ncomp] == list(b.table)[:ncomp]
        assert list(a.init_pred)[:ncomp] == list(b.init_pred)[:ncomp]
        assert (a.out_offset, a.out_pitch, a.out_x, a.out_y, a.store_w) == \
               (b.out_offset, b.out_pitch, b.out_x, b.out_y, b.store_w)


@pytest.mark.parametrize("seed", range(200))
def test_host_half_against_the_oracle_on_mutated_tiles(seed):
    """What the mirror decides about a tile BEFORE anything reaches the device, against the
    oracle's verdict on the whole tile (byte-mutated LJPEG tiles as in
    tests/test_oracle_mutation_fuzz.py): a tile the oracle decodes is never refused; a tile the
    host half refuses fails in the oracle too, with the same exception class -- except for the
    documented deviation: restart markers are validated up front, so a tile whose entropy data
    is ALSO corrupt before a bad marker reports the marker (known deviations)."""
    from test_oracle_mutation_fuzz import _mutate
    rng = np.random.default_rng(7000 + seed)
    ncomp = int(rng.choice([1, 2, 3, 4]))
    tw, th = 8 * ncomp * int(rng.integers(1, 5)), int(rng.integers(2, 12))
    img = synth.image_model(tw, th, seed, bits=14)
    two = ncomp >= 2 and bool(rng.integers(0, 2))
    t = synth.make_dng_ljpeg(img, tw, th, ncomp=ncomp, tabs=synth.default_tables(2 if two else 1),
                             tab_of_comp=[c % 2 if two else 0 for c in range(ncomp)],
                             restart_rows=int(rng.choice([0, 0, 1, 2])))
    blob = _mutate(rng, t.blob)
    try:
        port.ljpeg_decode(blob, port.new_image(tw, th), tw, 1, (0, 0), (tw, th), (tw, th))
        want = "ok"
    except port.IOException:
        want = "IOE"
    except port.RawDecoderException:
        want = "RDE"
    r = host.dng_ljpeg_host_half(blob, [0], [blob.size], tw, th, 1, tw, th, False, 1, 1)
    if want == "ok":
        assert r["errors"] == 0, r["first_error"]
    if r["errors"]:
        assert want != "ok"
        marker_first = any(m in r["first_error"] for m in ("Not a restart marker!", "Jpeg marker not encountered",
                                                           "Unexpected restart marker found"))
        if not marker_first:
            assert r["first_error"].startswith(want + ": "), (want, r["first_error"])
