"""Minimal synthetic DNG / TIFF container writer (test infrastructure): wraps tile or strip
payloads produced by oracle.synth into a file the reference's RawParser -> DngDecoder accepts
(little-endian classic TIFF, one IFD, the tags DngDecoder::decodeRawInternal reads:
decoders/DngDecoder.cpp:448-534, tile table :301-359)."""
import struct

import numpy as np

BYTE, ASCII, SHORT, LONG = 1, 2, 3, 4
_SIZE = {BYTE: 1, ASCII: 1, SHORT: 2, LONG: 4}


def _pack(typ, values):
    if typ == ASCII:
        return values if isinstance(values, bytes) else values.encode() + b"\0"
    if typ == BYTE:
        return bytes(values)
    fmt = {SHORT: "<%dH", LONG: "<%dI"}[typ] % len(values)
    return struct.pack(fmt, *values)


def build_tiff(entries, payload):
    """entries: list of (tag, type, values) -- values referring to payload offsets are callables
    f(base) -> list; payload: bytes placed behind the IFD and the out-of-line values."""
    entries = sorted(entries, key=lambda e: e[0])
    n = len(entries)
    ifd_off = 8
    data_off = ifd_off + 2 + 12 * n + 4
    # first pass: sizes of the out-of-line values
    blobs = []
    extra = 0
    for tag, typ, vals in entries:
        v = vals(0) if callable(vals) else vals
        raw = _pack(typ, v)
        if len(raw) > 4:
            extra += (len(raw) + 1) & ~1
    base = (data_off + extra + 15) & ~15  # payload starts here
    out = bytearray(b"II*\0" + struct.pack("<I", ifd_off))
    out += struct.pack("<H", n)
    tail = bytearray()
    for tag, typ, vals in entries:
        v = vals(base) if callable(vals) else vals
        raw = _pack(typ, v)
        count = len(raw) // _SIZE[typ]
        if len(raw) <= 4:
            field = raw + b"\0" * (4 - len(raw))
        else:
            field = struct.pack("<I", data_off + len(tail))
            tail += raw + (b"\0" if len(raw) & 1 else b"")
        out += struct.pack("<HHI", tag, typ, count) + field
    out += struct.pack("<I", 0)
    out += tail
    out += b"\0" * (base - len(out))
    out += bytes(payload)
    return np.frombuffer(bytes(out), dtype=np.uint8).copy()


def dng_common(w, h, bps, compression, cpp=1, cfa=True, version=(1, 4, 0, 0)):
    e = [
        (254, LONG, [0]),                    # NewSubFileType: the raw image
        (256, LONG, [w]), (257, LONG, [h]),
        (258, SHORT, [bps] * cpp),
        (259, SHORT, [compression]),
        (262, SHORT, [32803 if cfa else 34892]),
        (271, ASCII, "rsb200"), (272, ASCII, "synthetic"),
        (277, SHORT, [cpp]),
        (50706, BYTE, list(version)),        # DNGVersion
        (50708, ASCII, "rsb200 synthetic"),  # UniqueCameraModel
    ]
    if cfa:
        e += [(33421, SHORT, [2, 2]), (33422, BYTE, [0, 1, 1, 2])]
    return e


def make_dng_tiles(w, h, bps, tile_w, tile_h, blob, offsets, lengths, compression=7, cpp=1, cfa=True,
                   version=(1, 4, 0, 0)):
    """Tiled DNG: `blob` holds the tiles at `offsets` / `lengths` (row-major tile order)."""
    e = dng_common(w, h, bps, compression, cpp, cfa, version)
    offs = [int(o) for o in offsets]
    e += [(322, LONG, [tile_w]), (323, LONG, [tile_h]),
          (324, LONG, lambda base: [base + o for o in offs]),
          (325, LONG, [int(n) for n in lengths])]
    return build_tiff(e, blob)


def make_dng_strips(w, h, bps, rows_per_strip, data, pitch, cpp=1, cfa=True):
    """Uncompressed DNG in strips of rows_per_strip rows (compression 1)."""
    e = dng_common(w, h, bps, 1, cpp, cfa)
    nstrips = (h + rows_per_strip - 1) // rows_per_strip
    offs = [s * rows_per_strip * pitch for s in range(nstrips)]
    lens = [min(rows_per_strip, h - s * rows_per_strip) * pitch for s in range(nstrips)]
    e += [(278, LONG, [rows_per_strip]),
          (273, LONG, lambda base: [base + o for o in offs]),
          (279, LONG, lens)]
    return build_tiff(e, data)
