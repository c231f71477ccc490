"""ctypes binding of librs_oracle.so (rs_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "librs_oracle.so")

LSB, MSB, MSB16, MSB32, JPEG = 0, 1, 2, 3, 4
OK, RDE, IOE = 0, 1, 2


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code
        self.msg = msg


class RawDecoderException(OracleError):
    pass


class IOException(OracleError):
    pass


def raise_for(code, msg):
    if code == OK:
        return
    if code == IOE:
        raise IOException(code, msg)
    raise RawDecoderException(code, msg)


class Err(C.Structure):
    _fields_ = [("code", C.c_int), ("msg", C.c_char * 240)]

    def check(self, rc):
        raise_for(rc, self.msg.decode("utf-8", "replace"))


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int), ("h", C.c_int),
                ("cpp", C.c_int), ("pitch", C.c_int), ("is_cfa", C.c_int),
                ("sub_x", C.c_int), ("sub_y", C.c_int), ("is_f32", C.c_int)]


class Frame(C.Structure):
    _fields_ = [("mcu_x", C.c_int), ("mcu_y", C.c_int), ("dim_x", C.c_int),
                ("dim_y", C.c_int)]


class Dht(C.Structure):
    _fields_ = [("ncpl", C.c_uint8 * 16), ("values", C.c_uint8 * 162),
                ("nvalues", C.c_int)]


def build(force=False):
    """(Re)build librs_oracle.so with gcc; returns its path."""
    src = os.path.join(_HERE, "rs_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "rs_oracle.h"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.rso_huff_create.restype = C.c_void_p
        _lib.rso_huff_destroy.argtypes = [C.c_void_p]
        _lib.rso_huff_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.rso_huff_decode.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int,
                                         C.c_int, C.c_void_p, C.POINTER(Err)]
        _lib.rso_encode_diffs.restype = C.c_int64
        _lib.rso_ljpeg_encode.restype = C.c_int64
        _lib.rso_cr2_encode.restype = C.c_int64
    return _lib


def image_pitch(w, cpp=1):
    return lib().rso_image_pitch(w, cpp)


def new_image(w, h, cpp=1, fill=0xA5A5):
    """Uncropped uint16 buffer, shape (h, pitch/2), like RawImageData::createData."""
    pitch = image_pitch(w, cpp)
    return np.full((h, pitch // 2), fill, dtype=np.uint16)


def _img(arr, w, cpp, is_cfa=True, sub=(1, 1)):
    """uint16 array = UINT16 image; uint32 array (new_image_f32) = F32 image."""
    assert arr.dtype in (np.uint16, np.uint32) and arr.flags.c_contiguous
    return Image(arr.ctypes.data, w, arr.shape[0], cpp, arr.shape[1] * arr.itemsize,
                 1 if is_cfa else 0, sub[0], sub[1], int(arr.dtype == np.uint32))


def _u8(data):
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        return data.ctypes.data_as(C.c_char_p), data.size
    b = bytes(data)
    return b, len(b)


def pump_getbits(order, data, lens, want_pos=False):
    p, n = _u8(data)
    lens_a = (C.c_int * len(lens))(*lens)
    out = (C.c_uint32 * len(lens))()
    pos = C.c_int(0)
    e = Err()
    rc = lib().rso_pump_getbits_pos(order, p, n, lens_a, len(lens), out,
                                    C.byref(pos), C.byref(e))
    e.check(rc)
    return (list(out), pos.value) if want_pos else list(out)


def huff_extend(diff, length):
    return lib().rso_huff_extend(C.c_uint32(diff), C.c_uint32(length))


class Huff:
    """HuffmanCode + PrefixCodeDecoder<> (DHT counts/values -> decoder)."""

    def __init__(self, ncpl, values, full=True, fix16=False):
        self.ncpl = bytes(ncpl)
        self.values = bytes(values)
        assert len(self.ncpl) == 16
        e = Err()
        self.h = lib().rso_huff_create(self.ncpl, self.values, len(self.values),
                                       int(full), int(fix16), C.byref(e))
        if not self.h:
            raise_for(e.code or RDE, e.msg.decode("utf-8", "replace"))

    def __del__(self):
        if getattr(self, "h", None):
            lib().rso_huff_destroy(self.h)
            self.h = None

    def symbols(self):
        codes = (C.c_uint16 * 162)()
        lens = (C.c_uint8 * 162)()
        n = lib().rso_huff_symbols(self.h, codes, lens)
        return [(codes[i], lens[i]) for i in range(n)]

    def decode(self, data, n, order=JPEG):
        p, sz = _u8(data)
        out = (C.c_int32 * n)()
        e = Err()
        rc = lib().rso_huff_decode(self.h, order, p, sz, n, out, C.byref(e))
        e.check(rc)
        return list(out)

    def dht(self):
        d = Dht()
        for i in range(16):
            d.ncpl[i] = self.ncpl[i]
        for i, v in enumerate(self.values):
            d.values[i] = v
        d.nvalues = len(self.values)
        return d


def _hts(hts):
    arr = (C.c_void_p * len(hts))(*[h.h for h in hts])
    return arr


def unpack(data, img, w, cpp, crop, in_pitch, bps, order):
    """UncompressedDecompressor(...).readUncompressedRaw() into img (in place)."""
    p, n = _u8(data)
    im = _img(img, w, cpp)
    e = Err()
    rc = lib().rso_unpack(p, C.c_uint32(n), C.byref(im), crop[0], crop[1], crop[2],
                          crop[3], in_pitch, bps, order, C.byref(e))
    e.check(rc)
    return img


FORM_READ, FORM_8BIT, FORM_8BIT_UNCORRECTED = 0, 1, 2
FORM_12BIT_CONTROL_BE, FORM_12BIT_CONTROL_LE, FORM_12BIT_LEFT_BE, FORM_12BIT_LEFT_LE = 3, 4, 5, 6


def new_image_f32(w, h, cpp=1, fill=0xA5A5A5A5):
    """Uncropped 32-bit buffer of an F32 RawImage (bpp = 4*cpp, pitch rounded to 16),
    kept as uint32 so bit patterns (NaN payloads) compare exactly."""
    pitch = (w * cpp * 4 + 15) // 16 * 16
    return np.full((h, pitch // 4), fill, dtype=np.uint32)


def build_table(curve, dither):
    """TableLookUp::setTable (common/TableLookUp.cpp:48-85): the storage of table 0.
    Non-dithered: 65536 entries t[i] = curve[min(i, n-1)].  Dithered: 2*65536 entries,
    t[2i] = clampBits(center - (upper-lower+2)/4, 16), t[2i+1] = upper-lower."""
    curve = np.asarray(curve, dtype=np.int64)
    n = curve.size
    assert 0 < n <= 65536
    if not dither:
        idx = np.minimum(np.arange(65536), n - 1)
        return curve[idx].astype(np.uint16)
    t = np.zeros(2 * 65536, dtype=np.uint16)
    center = curve
    lower = np.concatenate([curve[:1], curve[:-1]])
    upper = np.concatenate([curve[1:], curve[-1:]])
    lower = np.minimum(lower, center)
    upper = np.maximum(upper, center)
    delta = upper - lower
    t[0:2 * n:2] = np.clip(center - ((upper - lower + 2) // 4), 0, 65535).astype(np.uint16)
    t[1:2 * n:2] = delta.astype(np.uint16)
    t[2 * n::2] = np.uint16(curve[-1])
    return t


def unpack_form(data, img, w, cpp, crop, in_pitch, bps, order, form, table=None,
                dither=False):
    """The other UncompressedDecompressor members (rso_unpack_form); `img` is a
    uint16 image, or a uint32 array from new_image_f32 for the F32 forms."""
    p, n = _u8(data)
    is_f32 = img.dtype == np.uint32
    assert img.flags.c_contiguous
    im = _img(img, w, cpp)
    tp = None
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.uint16)
        tp = table.ctypes.data_as(C.POINTER(C.c_uint16))
    e = Err()
    L = lib()
    L.rso_unpack_form.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(Image), C.c_int] + \
        [C.c_int] * 8 + [C.POINTER(C.c_uint16), C.c_int, C.POINTER(Err)]
    rc = L.rso_unpack_form(p, C.c_uint32(n), C.byref(im), int(is_f32), crop[0], crop[1],
                           crop[2], crop[3], in_pitch, bps, order, form, tp, int(dither),
                           C.byref(e))
    e.check(rc)
    return img


def nikon_tree(sel):
    """NikonDecompressor::nikon_tree[sel] as (ncpl[16], values)."""
    ncpl = (C.c_uint8 * 16)()
    vals = (C.c_uint8 * 16)()
    L = lib()
    L.rso_nikon_tree.argtypes = [C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
    n = L.rso_nikon_tree(sel, ncpl, vals)
    assert n > 0
    return bytes(ncpl), bytes(vals[:n])


def nikon_setup(meta, meta_be, bits, w, h):
    """What the NikonDecompressor constructor derives from the maker note:
    dict(curve, pup = [pUp00, pUp01, pUp10, pUp11], huff_select, split)."""
    mp, mn = _u8(meta)
    curve = np.zeros(32770, dtype=np.uint16)
    nc, hs, sp = C.c_int(0), C.c_int(0), C.c_int(0)
    pup = (C.c_int * 4)()
    e = Err()
    L = lib()
    L.rso_nikon_setup.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_uint16), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(Err)]
    rc = L.rso_nikon_setup(mp, mn, int(meta_be), bits, w, h,
                           curve.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(nc), pup,
                           C.byref(hs), C.byref(sp), C.byref(e))
    e.check(rc)
    return dict(curve=curve[:nc.value].copy(), pup=list(pup), huff_select=hs.value, split=sp.value)


def nikon_decompress(img, w, meta, meta_be, bits, data, uncorrected=False):
    """NikonDecompressor(img, meta, bits).decompress(data, uncorrected) into img (in place)."""
    mp, mn = _u8(meta)
    p, n = _u8(data)
    im = _img(img, w, 1)
    e = Err()
    L = lib()
    L.rso_nikon_decompress.argtypes = [C.POINTER(Image), C.c_char_p, C.c_int, C.c_int, C.c_int,
                                       C.c_char_p, C.c_uint32, C.c_int, C.POINTER(Err)]
    rc = L.rso_nikon_decompress(C.byref(im), mp, mn, int(meta_be), bits, p, C.c_uint32(n),
                                int(uncorrected), C.byref(e))
    e.check(rc)
    return img


def _strips(strips):
    n = len(strips)
    off = (C.c_uint64 * n)(*[s[0] for s in strips])
    ln = (C.c_uint32 * n)(*[s[1] for s in strips])
    rown = (C.c_int32 * n)(*[s[2] for s in strips])
    return off, ln, rown, n


def hasselblad_decompress(img, w, ht, init_pred, data):
    """HasselbladDecompressor(img, {ht, init_pred}, data).decompress() into img; returns the
    stream position.  ht: Huff(ncpl, values, full=False)."""
    p, n = _u8(data)
    im = _img(img, w, 1)
    consumed = C.c_uint32(0)
    e = Err()
    L = lib()
    L.rso_hasselblad_decompress.argtypes = [C.POINTER(Image), C.c_void_p, C.c_uint16, C.c_char_p,
                                            C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(Err)]
    rc = L.rso_hasselblad_decompress(C.byref(im), ht.h, init_pred, p, C.c_uint32(n),
                                     C.byref(consumed), C.byref(e))
    e.check(rc)
    return consumed.value


def phaseone(img, w, file, strips):
    """PhaseOneDecompressor(img, strips).decompress(); strips: [(offset, size, row)]."""
    p, n = _u8(file)
    im = _img(img, w, 1)
    off, ln, rown, ns = _strips(strips)
    e = Err()
    L = lib()
    L.rso_phaseone.argtypes = [C.POINTER(Image), C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64),
                               C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_int,
                               C.POINTER(Err)]
    rc = L.rso_phaseone(C.byref(im), p, C.c_uint64(n), off, ln, rown, ns, C.byref(e))
    e.check(rc)
    return img


def panasonic_v4(img, w, data, zero_is_not_bad=True, split=0, cap=1 << 16):
    """PanasonicV4Decompressor(img, data, zero_is_not_bad, split).decompress(); returns the
    sorted list of bad (zero) pixel positions (row << 16 | col)."""
    p, n = _u8(data)
    im = _img(img, w, 1)
    z = (C.c_uint32 * cap)()
    nz = C.c_uint32(0)
    e = Err()
    L = lib()
    L.rso_panasonic_v4.argtypes = [C.POINTER(Image), C.c_char_p, C.c_uint32, C.c_int, C.c_uint32,
                                   C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32),
                                   C.POINTER(Err)]
    rc = L.rso_panasonic_v4(C.byref(im), p, C.c_uint32(n), int(zero_is_not_bad), split, z, cap,
                            C.byref(nz), C.byref(e))
    e.check(rc)
    return sorted(z[:min(nz.value, cap)])


def panasonic(version, img, w, data, bps=14):
    """PanasonicV{5,6,7}Decompressor(img, data[, bps]).decompress() into img (in place)."""
    p, n = _u8(data)
    im = _img(img, w, 1)
    e = Err()
    L = lib()
    L.rso_panasonic.argtypes = [C.c_int, C.POINTER(Image), C.c_char_p, C.c_uint32, C.c_int,
                                C.POINTER(Err)]
    rc = L.rso_panasonic(version, C.byref(im), p, C.c_uint32(n), bps, C.byref(e))
    e.check(rc)
    return img


def scale_uses_sse2(black_sep, white):
    """RawImageDataU16::scaleValues' path choice on x86 (app_scale < 63 -> SSE2)."""
    b = (C.c_int * 4)(*[int(v) for v in black_sep])
    L = lib()
    L.rso_scale_uses_sse2.argtypes = [C.POINTER(C.c_int), C.c_int]
    return bool(L.rso_scale_uses_sse2(b, int(white)))


def scale_values(img, w, crop, black_sep, white, dither=True, sse2=None):
    """RawImageDataU16::scaleValues over crop = (off_x, off_y, crop_w, crop_h), in place;
    sse2=None picks the path the reference picks on x86."""
    if sse2 is None:
        sse2 = scale_uses_sse2(black_sep, white)
    im = _img(img, w, 1)
    b = (C.c_int * 4)(*[int(v) for v in black_sep])
    e = Err()
    L = lib()
    L.rso_scale_values.argtypes = [C.POINTER(Image)] + [C.c_int] * 4 + [C.POINTER(C.c_int)] + \
        [C.c_int] * 3 + [C.POINTER(Err)]
    rc = L.rso_scale_values(C.byref(im), crop[0], crop[1], crop[2], crop[3], b, int(white),
                            int(dither), int(sse2), C.byref(e))
    e.check(rc)
    return img


def dng_opcodes(img, w, cpp, crop, data, cap=1 << 20):
    """DngOpcodes(ri, data) + applyOpCodes(ri) in place; img: uint16 image or uint32 array of
    an F32 image; crop = [off_x, off_y, crop_w, crop_h].  Returns (crop after TrimBounds,
    mBadPixelPositions).  On an exception `dng_opcodes.partial` holds (crop, list, applied)."""
    p, n = _u8(data)
    im = _img(img, w, cpp)
    cr = (C.c_int * 4)(*[int(v) for v in crop])
    bad = (C.c_uint32 * cap)()
    nbad = C.c_uint32(0)
    applied = C.c_int(0)
    e = Err()
    L = lib()
    L.rso_dng_opcodes.argtypes = [C.POINTER(Image), C.POINTER(C.c_int), C.c_char_p, C.c_uint32,
                                  C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_int), C.POINTER(Err)]
    rc = L.rso_dng_opcodes(C.byref(im), cr, p, C.c_uint32(n), bad, cap, C.byref(nbad),
                           C.byref(applied), C.byref(e))
    dng_opcodes.partial = (list(cr), list(bad[:min(nbad.value, cap)]), applied.value)
    e.check(rc)
    return list(cr), list(bad[:min(nbad.value, cap)])


def sixteen_bit_lookup(img, w, cpp, table, dither):
    """RawImageData::sixteenBitLookup() in place; table = build_table(curve, dither) storage."""
    im = _img(img, w, cpp)
    tp = None
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.uint16)
        tp = table.ctypes.data_as(C.POINTER(C.c_uint16))
    e = Err()
    L = lib()
    L.rso_sixteen_bit_lookup.argtypes = [C.POINTER(Image), C.POINTER(C.c_uint16), C.c_int, C.POINTER(Err)]
    e.check(L.rso_sixteen_bit_lookup(C.byref(im), tp, int(dither), C.byref(e)))
    return img


def fix_bad_pixels(img, w, cpp, positions, is_cfa=True):
    """RawImageData::fixBadPixels() with mBadPixelPositions = positions, in place."""
    im = _img(img, w, cpp, is_cfa)
    pos = np.ascontiguousarray(positions, dtype=np.uint32)
    e = Err()
    L = lib()
    L.rso_fix_bad_pixels.argtypes = [C.POINTER(Image), C.c_void_p, C.c_uint32, C.POINTER(Err)]
    e.check(L.rso_fix_bad_pixels(C.byref(im), pos.ctypes.data, pos.size, C.byref(e)))
    return img


class BlackArea(C.Structure):
    _fields_ = [("offset", C.c_uint32), ("size", C.c_uint32), ("is_vertical", C.c_int)]


def scale_black_white(img, w, crop, black_level=-1, black_sep=None, white=None, areas=(),
                      dither=True, is_cfa=True, sse2=None, cpp=1):
    """RawImageDataU16::scaleBlackWhite() in place; areas: [(is_vertical, offset, size)].
    Returns (black_sep or None when no scaling happened, white)."""
    im = _img(img, w, cpp)
    im.is_cfa = int(is_cfa)
    b = (C.c_int * 4)(*([int(v) for v in black_sep] if black_sep is not None else [-7] * 4))
    wh = C.c_int(int(white) if white is not None else 0)
    ar = (BlackArea * max(1, len(areas)))(*[BlackArea(o, s, int(v)) for v, o, s in areas])
    e = Err()
    L = lib()
    L.rso_scale_black_white.argtypes = [C.POINTER(Image)] + [C.c_int] * 5 + [C.POINTER(C.c_int), C.c_int,
                                        C.POINTER(C.c_int), C.c_int, C.POINTER(BlackArea)] + \
        [C.c_int] * 3 + [C.POINTER(Err)]
    rc = L.rso_scale_black_white(C.byref(im), crop[0], crop[1], crop[2], crop[3], int(black_level),
                                 b, int(black_sep is not None), C.byref(wh), int(white is not None),
                                 ar, len(areas), int(dither), -1 if sse2 is None else int(sse2),
                                 C.byref(e))
    e.check(rc)
    sep = list(b)
    return (None if sep == [-7] * 4 else sep), wh.value


def sony_arw2(img, w, data, table=None, dither=False):
    """SonyArw2Decompressor(img, data).decompress() into img (in place); `table` = the
    storage build_table() returns (None: the image has no table)."""
    p, n = _u8(data)
    im = _img(img, w, 1)
    tp = None
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.uint16)
        tp = table.ctypes.data_as(C.POINTER(C.c_uint16))
    e = Err()
    L = lib()
    L.rso_sony_arw2.argtypes = [C.POINTER(Image), C.c_char_p, C.c_uint32, C.POINTER(C.c_uint16),
                                C.c_int, C.POINTER(Err)]
    rc = L.rso_sony_arw2(C.byref(im), p, C.c_uint32(n), tp, int(dither), C.byref(e))
    e.check(rc)
    return img


def ljpeg_decompress(img, w, cpp, img_frame, mcu, frame_dim, hts, init_pred,
                     rows_per_restart, data):
    p, n = _u8(data)
    im = _img(img, w, cpp)
    fr = Frame(mcu[0], mcu[1], frame_dim[0], frame_dim[1])
    ip = (C.c_uint16 * len(init_pred))(*init_pred)
    consumed = C.c_uint32(0)
    e = Err()
    rc = lib().rso_ljpeg_decompress(C.byref(im), img_frame[0], img_frame[1],
                                    img_frame[2], img_frame[3], fr, _hts(hts), ip,
                                    len(hts), rows_per_restart, p, C.c_uint32(n),
                                    C.byref(consumed), C.byref(e))
    e.check(rc)
    return consumed.value


def ljpeg_decode(blob, img, w, cpp, off, size, max_dim, fix16=False):
    p, n = _u8(blob)
    im = _img(img, w, cpp)
    e = Err()
    rc = lib().rso_ljpeg_decode(p, C.c_uint32(n), C.byref(im), off[0], off[1],
                                size[0], size[1], max_dim[0], max_dim[1],
                                int(fix16), C.byref(e))
    e.check(rc)
    return img


def dng_decompress(file_bytes, tile_off, tile_len, img, w, cpp, tile_w, tile_h,
                   compression, fix_ljpeg=False, bps=14, big_endian=False,
                   nthreads=1):
    p, n = _u8(file_bytes)
    im = _img(img, w, cpp)
    offs = (C.c_uint64 * len(tile_off))(*tile_off)
    lens = (C.c_uint32 * len(tile_len))(*tile_len)
    e = Err()
    rc = lib().rso_dng_decompress(p, C.c_uint64(n), offs, lens, len(tile_off),
                                  C.byref(im), tile_w, tile_h, compression,
                                  int(fix_ljpeg), bps, int(big_endian), nthreads,
                                  C.byref(e))
    e.check(rc)
    return img


def pentax_decompress(img, w, data, meta=None, meta_be=True):
    """PentaxDecompressor(img, meta).decompress(data) into img (in place)."""
    p, n = _u8(data)
    im = _img(img, w, 1)
    mp, mn = (None, 0) if meta is None else _u8(meta)
    e = Err()
    L = lib()
    L.rso_pentax_decompress.argtypes = [C.POINTER(Image), C.c_char_p, C.c_int, C.c_int,
                                        C.c_char_p, C.c_uint32, C.POINTER(Err)]
    rc = L.rso_pentax_decompress(C.byref(im), mp, mn, int(meta_be), p, C.c_uint32(n),
                                 C.byref(e))
    e.check(rc)
    return img


def pentax_table(meta=None, meta_be=True):
    """(ncpl[16], values) the PentaxDecompressor constructor builds."""
    mp, mn = (None, 0) if meta is None else _u8(meta)
    ncpl = (C.c_uint8 * 16)()
    vals = (C.c_uint8 * 16)()
    e = Err()
    L = lib()
    L.rso_pentax_table.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint8),
                                   C.POINTER(C.c_uint8), C.POINTER(Err)]
    L.rso_pentax_table.restype = C.c_int
    n = L.rso_pentax_table(mp, mn, int(meta_be), ncpl, vals, C.byref(e))
    if e.code != OK:
        raise_for(e.code, e.msg.decode("utf-8", "replace"))
    return list(ncpl), list(vals)[:n]


def encode_diffs_plain(diffs, ht):
    """Huffman-encode diffs with one table into a plain MSB stream (test inputs)."""
    d = np.ascontiguousarray(diffs, dtype=np.int32)
    cap = d.size * 5 + 64
    out = np.empty(cap, dtype=np.uint8)
    L = lib()
    L.rso_encode_diffs_plain.restype = C.c_int64
    # (the handle is a pointer: without argtypes ctypes would pass a Python int as a C int)
    L.rso_encode_diffs_plain.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
    n = L.rso_encode_diffs_plain(d.ctypes.data_as(C.c_void_p), C.c_uint64(d.size), ht.h,
                                 out.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
    if n < 0:
        raise ValueError("encode_diffs_plain failed (%d)" % n)
    return out[:n].copy()


def sraw_interpolate(inp, in_w, out, out_w, sub, coeffs, hue, version):
    """Cr2sRawInterpolator(out, inp, coeffs, hue).interpolate(version).
    inp: uint16 array (rows, pitch/2) of which in_w columns are the subsampled data;
    out: uint16 3-component image buffer from new_image(out_w, out_h, 3)."""
    assert inp.dtype == np.uint16 and inp.flags.c_contiguous
    im = _img(out, out_w, 3, False, sub)
    k = (C.c_int * 3)(*coeffs)
    e = Err()
    L = lib()
    L.rso_sraw_interpolate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Image),
                                       C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(Err)]
    rc = L.rso_sraw_interpolate(inp.ctypes.data, in_w, inp.shape[0], inp.shape[1] * 2,
                                C.byref(im), k, hue, version, C.byref(e))
    e.check(rc)
    return out


def cr2_decompress(img, w, fmt, frame, slicing, hts, init_pred, data, is_cfa=True):
    p, n = _u8(data)
    im = _img(img, w, 1, is_cfa)
    ip = (C.c_uint16 * len(init_pred))(*init_pred)
    consumed = C.c_uint32(0)
    e = Err()
    rc = lib().rso_cr2_decompress(C.byref(im), fmt[0], fmt[1], fmt[2], frame[0],
                                  frame[1], slicing[0], slicing[1], slicing[2],
                                  _hts(hts), ip, len(hts), p, C.c_uint32(n),
                                  C.byref(consumed), C.byref(e))
    e.check(rc)
    return consumed.value


def cr2_ljpeg_decode(blob, img, w, slicing, is_cfa=True, sub=(1, 1)):
    p, n = _u8(blob)
    im = _img(img, w, 1, is_cfa, sub)
    e = Err()
    rc = lib().rso_cr2_ljpeg_decode(p, C.c_uint32(n), C.byref(im), slicing[0],
                                    slicing[1], slicing[2], C.byref(e))
    e.check(rc)
    return img


# ---------------- test-input tooling ----------------
def encode_diffs(diffs, hts, comp_of):
    d = np.ascontiguousarray(diffs, dtype=np.int32)
    cap = d.size * 5 + 64
    out = np.empty(cap, dtype=np.uint8)
    co = (C.c_uint8 * len(comp_of))(*comp_of)
    n = lib().rso_encode_diffs(d.ctypes.data_as(C.c_void_p), C.c_uint64(d.size),
                               _hts(hts), co, len(comp_of),
                               out.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
    if n < 0:
        raise ValueError("encode_diffs failed (%d)" % n)
    return out[:n].tobytes()


def _dhts(tabs):
    arr = (Dht * len(tabs))()
    for i, t in enumerate(tabs):
        arr[i] = t.dht() if isinstance(t, Huff) else t
    return arr


def ljpeg_encode(samples, frame_w, frame_h, mcu, prec, tabs, tab_of_comp,
                 restart_rows=0, fix16=False):
    """samples: 2-D uint16 array, rows = frame_h*mcu_y, cols >= frame_w*mcu_x."""
    s = np.ascontiguousarray(samples, dtype=np.uint16)
    assert s.shape[0] >= frame_h * mcu[1] and s.shape[1] >= frame_w * mcu[0]
    cap = int(frame_w) * frame_h * mcu[0] * mcu[1] * 5 + 4096
    out = np.empty(cap, dtype=np.uint8)
    toc = (C.c_uint8 * len(tab_of_comp))(*tab_of_comp)
    n = lib().rso_ljpeg_encode(s.ctypes.data_as(C.c_void_p), s.shape[1], frame_w,
                               frame_h, mcu[0], mcu[1], prec, _dhts(tabs), len(tabs),
                               toc, restart_rows, int(fix16),
                               out.ctypes.data_as(C.c_void_p), C.c_uint64(cap))
    if n < 0:
        raise ValueError("ljpeg_encode failed (%d)" % n)
    return out[:n].copy()


def cr2_encode(img, w, fmt, frame, slicing, prec, tabs, tab_of_comp, is_cfa=True):
    im = _img(img, w, 1, is_cfa)
    cap = int(img.shape[0]) * w * 5 + 4096
    out = np.empty(cap, dtype=np.uint8)
    toc = (C.c_uint8 * len(tab_of_comp))(*tab_of_comp)
    n = lib().rso_cr2_encode(C.byref(im), fmt[0], fmt[1], fmt[2], frame[0], frame[1],
                             slicing[0], slicing[1], slicing[2], prec, _dhts(tabs),
                             len(tabs), toc, out.ctypes.data_as(C.c_void_p),
                             C.c_uint64(cap))
    if n < 0:
        raise ValueError("cr2_encode failed (%d)" % n)
    return out[:n].copy()
