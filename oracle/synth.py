"""Seeded synthetic inputs for the hot path.
TEST INFRASTRUCTURE ONLY -- inputs for tests and bench, never product code."""
from concurrent.futures import ThreadPoolExecutor
import os

import struct

import numpy as np

from . import port

# DHT used by the synthetic LJPEG streams: 17 codes, lengths 2..14.
DEFAULT_NCPL = bytes([0, 1, 5, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0])
DEFAULT_VALUES = bytes([6, 4, 5, 7, 3, 8, 2, 9, 1, 10, 0, 11, 12, 13, 14, 15, 16])
# A second, different table (for multi-table scans): lengths 2..15, other order.
ALT_NCPL = bytes([0, 2, 2, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0])
ALT_VALUES = bytes([5, 6, 4, 7, 3, 8, 2, 9, 1, 10, 0, 11, 12, 13, 14, 15, 16])


def lcg_u32(n, seed):
    """s_{k+1} = s_k*1664525 + 1013904223 (mod 2^32); returns s_1..s_n."""
    a = np.uint32(1664525)
    c = np.uint32(1013904223)
    out = np.empty(n, dtype=np.uint32)
    block = 1 << 16
    # closed form inside a block, sequential across blocks
    apow = np.cumprod(np.full(block, a, dtype=np.uint32), dtype=np.uint32)  # a^1..a^B
    geo = np.empty(block, dtype=np.uint32)  # 1 + a + ... + a^(k-1), k=1..B
    geo[0] = 1
    geo[1:] = (np.cumsum(apow[:-1], dtype=np.uint32) + np.uint32(1)).astype(np.uint32)
    s = np.uint32(seed)
    with np.errstate(over="ignore"):
        for lo in range(0, n, block):
            m = min(block, n - lo)
            out[lo:lo + m] = apow[:m] * s + c * geo[:m]
            s = out[lo + m - 1]
    return out


def image_model_c(w, h, seed):
    """image_model(w, h, seed) (non-wild, 14 bit) by the C oracle library, with the two checksums
    bench.py compares on the GPU: (image, sum, weighted sum)."""
    import ctypes as C
    L = port.lib()
    L.rso_image_model.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint64)]
    L.rso_image_model.restype = None
    img = np.empty((h, w), dtype=np.uint16)
    sums = (C.c_uint64 * 2)()
    L.rso_image_model(w, h, seed, img.ctypes.data, sums)
    return img, int(sums[0]), int(sums[1])


def lcg_bytes(n, seed):
    """byte_k = s_k >> 24 (SURVEY 8d C1/C2 input)."""
    return (lcg_u32(n, seed) >> np.uint32(24)).astype(np.uint8)


def packed_frame(w, h, bps, seed, pitch=None):
    """Random packed frame of `h` rows, pitch bytes per row (default minimal)."""
    row_bytes = w * bps // 8
    pitch = pitch or row_bytes
    return lcg_bytes(pitch * h, seed), pitch


def image_model(w, h, seed=12345, wild=False, bits=14):
    """px(x,y) = (2000 + ((7x+3y)&1023) + noise6 - 32) & mask  (SURVEY 8d C3);
    `wild`: full-range noise (long codes, diff lengths up to 15/16)."""
    mask = (1 << bits) - 1
    r = lcg_u32(w * h, seed).reshape(h, w)
    if wild:
        return ((r >> np.uint32(32 - bits)) & np.uint32(mask)).astype(np.uint16)
    y, x = np.mgrid[0:h, 0:w].astype(np.uint32)
    noise6 = (r >> np.uint32(26)) & np.uint32(63)
    v = (np.uint32(2000) + ((np.uint32(7) * x + np.uint32(3) * y) & np.uint32(1023))
         + noise6 - np.uint32(32)) & np.uint32(mask)
    return v.astype(np.uint16)


def default_tables(n=1):
    tabs = [port.Huff(DEFAULT_NCPL, DEFAULT_VALUES)]
    if n > 1:
        tabs.append(port.Huff(ALT_NCPL, ALT_VALUES))
    return tabs


class DngTiles:
    """A synthetic tiled-DNG payload: `blob` holds the per-tile LJPEG streams."""

    def __init__(self, blob, offsets, lengths, w, h, cpp, tile_w, tile_h):
        self.blob, self.offsets, self.lengths = blob, offsets, lengths
        self.w, self.h, self.cpp = w, h, cpp
        self.tile_w, self.tile_h = tile_w, tile_h


def make_dng_ljpeg(img, tile_w, tile_h, ncomp=2, prec=14, tabs=None,
                   tab_of_comp=None, restart_rows=0, fix16=False, mcu=None,
                   align=1, threads=None, cpp=1):
    """Encode `img` (h x (w*cpp) uint16, logical pixels only) as DNG LJPEG tiles
    (edge tiles encoded full-size with edge replication, as DNG writers do)."""
    h, wc = img.shape
    w = wc // cpp
    tabs = tabs or default_tables(1)
    mcu = mcu or (ncomp, 1)
    tab_of_comp = tab_of_comp or [0] * (mcu[0] * mcu[1])
    tiles_x = (w + tile_w - 1) // tile_w
    tiles_y = (h + tile_h - 1) // tile_h
    frame_w = tile_w * cpp // mcu[0]
    frame_h = tile_h // mcu[1]

    def enc(n):
        ty, tx = divmod(n, tiles_x)
        ys = np.minimum(np.arange(ty * tile_h, (ty + 1) * tile_h), h - 1)
        xs = np.minimum(np.arange(tx * tile_w, (tx + 1) * tile_w), w - 1)
        if cpp == 1:
            tile = img[np.ix_(ys, xs)]
        else:
            cols = (xs[:, None] * cpp + np.arange(cpp)[None, :]).reshape(-1)
            tile = img[np.ix_(ys, cols)]
        return port.ljpeg_encode(np.ascontiguousarray(tile), frame_w, frame_h, mcu,
                                 prec, tabs, tab_of_comp, restart_rows, fix16)

    n_tiles = tiles_x * tiles_y
    threads = threads or min(32, os.cpu_count() or 1)
    if n_tiles > 4 and threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            blobs = list(ex.map(enc, range(n_tiles)))
    else:
        blobs = [enc(n) for n in range(n_tiles)]
    offsets, lengths, pos = [], [], 0
    for b in blobs:
        pos = (pos + align - 1) // align * align
        offsets.append(pos)
        lengths.append(len(b))
        pos += len(b)
    blob = np.zeros(pos, dtype=np.uint8)
    for o, b in zip(offsets, blobs):
        blob[o:o + len(b)] = b
    return DngTiles(blob, offsets, lengths, w, h, cpp, tile_w, tile_h)


def md5_of_row_md5s(img, row_bytes=None):
    """rstest's image hash: md5 of the concatenated per-row md5 digests of the
    uncropped buffer (src/utilities/rstest/rstest.cpp:131-146)."""
    import hashlib
    rows = []
    for r in range(img.shape[0]):
        b = img[r].tobytes()
        rows.append(hashlib.md5(b if row_bytes is None else b[:row_bytes]).digest())
    return hashlib.md5(b"".join(rows)).hexdigest()


def pentax_modern_meta(big_endian=True):
    """A "modern" Pentax table description (PentaxDecompressor.cpp:83-141): depth 15 codes,
    canonical, lengths {2,2,3,3,3,4,5,6,7,8,9,10,11,12,12}."""
    lens = [3, 3, 3, 2, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 12]   # length of the code of value i
    order = sorted(range(15), key=lambda i: (lens[i], i))
    code, prev, codes = 0, 0, {}
    for i in order:
        code <<= (lens[i] - prev)
        prev = lens[i]
        codes[i] = code
        code += 1
    v0 = [codes[i] << (12 - lens[i]) for i in range(15)]
    out = bytearray()
    u16 = (lambda v: [v >> 8, v & 255]) if big_endian else (lambda v: [v & 255, v >> 8])
    out += bytes(u16(15 - 12))
    out += bytes(12)
    for v in v0:
        out += bytes(u16(v))
    out += bytes(lens)
    return bytes(out)


def make_pentax(img, table):
    """Encode a uint16 image (values < 32768, even width) the way PentaxDecompressor
    decodes it: per-parity left predictor, row starts from two rows up, one plain MSB
    stream.  `table` = (ncpl, values) as returned by port.pentax_table()."""
    from . import port
    h, w = img.shape
    a = img.astype(np.int32)
    d = np.zeros((h, w), dtype=np.int32)
    d[:, 2:] = a[:, 2:] - a[:, :-2]
    d[2:, :2] = a[2:, :2] - a[:-2, :2]
    d[:2, :2] = a[:2, :2]
    ht = port.Huff(table[0], table[1])
    return port.encode_diffs_plain(d.reshape(-1), ht)


def arw2_frame(w, h, seed):
    """Random SonyArw2 payload (one byte per pixel, 128-bit blocks): LCG bytes with the
    one invalid pattern (imax == imin, SonyArw2Decompressor.cpp:88-89) patched away."""
    assert w % 32 == 0
    a = lcg_bytes(w * h, seed).reshape(-1, 16).copy()
    # bits 22..25 = imax, 26..29 = imin (LSB-first fields: max 0..10, min 11..21)
    v = a[:, 2].astype(np.uint32) | (a[:, 3].astype(np.uint32) << 8)   # bits 16..31
    imax = (v >> 6) & 15
    imin = (v >> 10) & 15
    same = imax == imin
    imin = np.where(same, (imin + 1) & 15, imin)
    v = (v & ~np.uint32(15 << 10)) | (imin.astype(np.uint32) << 10)
    a[:, 2] = (v & 255).astype(np.uint8)
    a[:, 3] = (v >> 8).astype(np.uint8)
    return a.reshape(-1)


def sony_curve():
    """The curve ArwDecoder::SonyDecodeCurve-style tables look like: 0x4001 entries,
    identity up to a knee, then steeper segments (values <= 0x3fff -> 16 bit)."""
    x = np.arange(0x4001, dtype=np.int64)
    y = np.where(x < 2048, x, np.where(x < 4096, 2048 + (x - 2048) * 2,
                 np.where(x < 8192, 6144 + (x - 4096) * 3, 18432 + (x - 8192) * 4)))
    return np.minimum(y, 65535).astype(np.uint16)


def nikon_meta(kind, bits, pup=(2048, 2050, 2040, 2046), big_endian=True, split=0):
    """A Nikon maker-note block as NikonDecompressor's constructor reads it
    (NikonDecompressor.cpp:478-511, createCurve :380-441).  kind:
      "lossless"  v0=70 (identity curve, trees 2 / 5)
      "table"     v0=68, v1=16: csize curve values follow (trees 0 / 3)
      "segments"  v0=68, v1=32: csize knots, linear interpolation, split at offset 562
      "z7"        v0=68, v1=64: as segments with the 2-bit shorter curve
      "skip"      v0=73: 2110 bytes skipped first, then as "table"."""
    u16 = (lambda v: [v >> 8, v & 255]) if big_endian else (lambda v: [v & 255, v >> 8])
    out = bytearray()
    v0, v1 = {"lossless": (70, 48), "table": (68, 16), "segments": (68, 32), "z7": (68, 64),
              "skip": (73, 16)}[kind]
    out += bytes([v0, v1])
    if kind == "skip":
        out += bytes(2110)
    for v in pup:   # order in the stream: pUp[0][0], pUp[1][0], pUp[0][1], pUp[1][1]
        out += bytes(u16(v))
    if kind == "lossless":
        out += bytes(u16(0))
    elif kind in ("table", "skip"):
        n = 1 << bits
        xs = np.arange(n, dtype=np.int64)
        ys = np.minimum((xs * 3) // 2 + (xs * xs) // (n * 2), 65535)
        out += bytes(u16(n))
        for y in ys:
            out += bytes(u16(int(y)))
    else:
        cb = bits - 2 if kind == "z7" else bits
        n = ((1 << cb) & 0x7fff) + 1
        csize = 33     # (csize - 1) * step == n - 1 with step = (n - 1) / 32
        out += bytes(u16(csize))
        for k in range(csize):
            x = k * ((n - 1) // 32)
            out += bytes(u16(min(65535, x * 2 + (x * x) // (n // 2 + 1))))
        out += bytes(max(0, 562 - len(out)))
        out[562:564] = bytes(u16(split))
    return bytes(out)


def make_nikon(img, sel, pup):
    """Encode a uint16 image (pre-curve values, even width) the way NikonDecompressor
    decodes it (:513-538): per-parity left predictor, rows start from pUp[row & 1] which
    the first two pixels of every row update.  pup = [pUp00, pUp01, pUp10, pUp11]."""
    h, w = img.shape
    a = img.astype(np.int32)
    d = np.zeros((h, w), dtype=np.int32)
    d[:, 2:] = a[:, 2:] - a[:, :-2]
    d[2:, :2] = a[2:, :2] - a[:-2, :2]
    d[0, 0], d[0, 1] = a[0, 0] - pup[0], a[0, 1] - pup[1]
    if h > 1:
        d[1, 0], d[1, 1] = a[1, 0] - pup[2], a[1, 1] - pup[3]
    ncpl, values = port.nikon_tree(sel)
    return port.encode_diffs_plain(d.reshape(-1), port.Huff(ncpl, values))


_P1_LENGTH = [8, 7, 6, 9, 11, 10, 5, 12, 14, 13]


def phaseone_row(vals):
    """Encode one image row the way PhaseOneDecompressor::decompressStrip reads it
    (PhaseOneDecompressor.cpp:85-135): an MSB32 bit stream (32-bit little-endian chunks
    consumed MSB first); per 8 pixels two code lengths, chosen as the smallest that holds
    the 4 differences of that parity.  Returns bytes (multiple of 4, + 8 bytes slack)."""
    w = len(vals)
    bits = []

    def put(v, n):
        for k in range(n - 1, -1, -1):
            bits.append((v >> k) & 1)

    pred = [0, 0]
    ln = [0, 0]
    lim = w & ~7
    for col in range(w):
        if col >= lim:
            ln = [14, 14]
        elif col % 8 == 0:
            for t in (0, 1):
                p = pred[t]
                need = 5
                for c in range(col + t, col + 8, 2):
                    d = int(vals[c]) - p
                    p = int(vals[c])
                    n = 5
                    while n < 14 and not (0 <= d - 1 + (1 << (n - 1)) < (1 << n)):
                        n += 1
                    need = max(need, n)
                if col == 0:
                    need = max(need, 13)   # at column 0 a 1 bit in the prefix is an error: only 5 zeros
                idx = _P1_LENGTH.index(need)
                j, b = idx // 2 + 1, idx % 2
                put(0, j)
                if j < 5:
                    put(1, 1)
                put(b, 1)
                ln[t] = need
        i = ln[col & 1]
        v = int(vals[col])
        if i == 14:
            put(v & 0xFFFF, 16)
        else:
            put(v - pred[col & 1] - 1 + (1 << (i - 1)), i)
        pred[col & 1] = v
    while len(bits) % 32:
        bits.append(0)
    a = np.array(bits, dtype=np.uint8).reshape(-1, 32)
    words = (a.astype(np.uint64) << np.arange(31, -1, -1, dtype=np.uint64)).sum(axis=1).astype("<u4")
    return words.tobytes() + bytes(8)


def make_phaseone(img, shuffle_seed=None, gap=0):
    """file bytes + strips [(offset, size, row)] for PhaseOneDecompressor: one strip per
    row, stored in shuffled order with `gap` unused bytes between them."""
    h, w = img.shape
    rows = [phaseone_row(img[r]) for r in range(h)]
    order = list(range(h))
    if shuffle_seed is not None:
        rng = np.random.default_rng(shuffle_seed)
        rng.shuffle(order)
    blob = bytearray()
    strips = []
    for r in order:
        blob += bytes(gap)
        strips.append((len(blob), len(rows[r]), r))
        blob += rows[r]
    return np.frombuffer(bytes(blob), dtype=np.uint8).copy(), strips


def make_hasselblad(img, ht, init_pred):
    """Encode a uint16 image (even width) the way HasselbladDecompressor reads it
    (HasselbladDecompressor.cpp:72-100): one MSB32 bit stream, per pixel pair
    [len1 code][len2 code][len1 bits][len2 bits], both predictors reset to init_pred at
    every row.  ht: port.Huff over difference lengths 0..16 (any mode; only its codes are
    used).  Returns bytes (multiple of 4, + 16 bytes slack)."""
    codes = {v: cl for v, cl in zip(ht.values, ht.symbols())}
    bits = []

    def put(v, n):
        for k in range(n - 1, -1, -1):
            bits.append((v >> k) & 1)

    def cat(d):
        """(length, bits) of difference d (T.81 F.12 inverse; -32768 = length 16, all ones)"""
        if d == 0:
            return 0, 0
        if d == -32768:
            return 16, 0xFFFF
        n = abs(d).bit_length()
        return n, (d if d > 0 else d + (1 << n) - 1)

    h, w = img.shape
    for r in range(h):
        p = [init_pred, init_pred]
        for c in range(0, w, 2):
            enc = []
            for t in (0, 1):
                d = (int(img[r, c + t]) - p[t] + 32768) % 65536 - 32768   # difference mod 2^16
                enc.append(cat(d))
                p[t] = int(img[r, c + t])
            for n, _ in enc:
                code, cl = codes[n]
                put(code, cl)
            for n, b in enc:
                if n:
                    put(b, n)
    while len(bits) % 32:
        bits.append(0)
    a = np.array(bits, dtype=np.uint8).reshape(-1, 32)
    words = (a.astype(np.uint64) << np.arange(31, -1, -1, dtype=np.uint64)).sum(axis=1).astype("<u4")
    return np.frombuffer(words.tobytes() + bytes(16), dtype=np.uint8).copy()


def make_hasselblad_fast(img, ht, init_pred):
    """make_hasselblad() with numpy (same bytes; for frames of megapixels)."""
    codes = {v: cl for v, cl in zip(ht.values, ht.symbols())}
    h, w = img.shape
    a = img.astype(np.int64).reshape(h, w // 2, 2)
    prev = np.concatenate([np.full((h, 1, 2), init_pred, dtype=np.int64), a[:, :-1, :]], axis=1)
    d = (a - prev + 32768) % 65536 - 32768                      # differences mod 2^16, per component
    n = np.zeros(d.shape, dtype=np.int64)
    ad = np.abs(d)
    nz = ad > 0
    n[nz] = np.floor(np.log2(ad[nz])).astype(np.int64) + 1
    m = np.where(d > 0, d, d + (1 << n) - 1)
    m = np.where(d == -32768, 0xFFFF, m)
    n = np.where(d == -32768, 16, n)
    m = np.where(n == 0, 0, m)
    code_v = np.zeros(17, dtype=np.int64)
    code_l = np.zeros(17, dtype=np.int64)
    for v, (c, l) in codes.items():
        code_v[v], code_l[v] = c, l
    # tokens per pair in stream order: code1, code2, mantissa1, mantissa2
    val = np.stack([code_v[n[..., 0]], code_v[n[..., 1]], m[..., 0], m[..., 1]], axis=-1).reshape(-1)
    nb = np.stack([code_l[n[..., 0]], code_l[n[..., 1]], n[..., 0], n[..., 1]], axis=-1).reshape(-1)
    pos = np.concatenate([[0], np.cumsum(nb)])
    total = int(pos[-1])
    pos = pos[:-1]
    nwords = (total + 31) // 32 + 1
    off = pos & 31
    widx = pos >> 5
    v64 = val.astype(np.uint64) << (64 - off - nb).astype(np.uint64)  # left aligned in 64 bits at its offset
    hi = (v64 >> np.uint64(32)).astype(np.float64)
    lo = (v64 & np.uint64(0xFFFFFFFF)).astype(np.float64)
    words = np.bincount(widx, weights=hi, minlength=nwords + 1)[:nwords + 1] + \
        np.concatenate([[0.0], np.bincount(widx, weights=lo, minlength=nwords)[:nwords]])
    words = words[:(total + 31) // 32].astype(np.uint64).astype("<u4")
    return np.frombuffer(words.tobytes() + bytes(16), dtype=np.uint8).copy()


def hasselblad_ljpeg_container(w, h, data, ncpl, values, prec=16, predictor=1, frame_w=None, frame_h=None,
                               dri=None):
    """The LJPEG container HasselbladLJpegDecoder reads (HasselbladLJpegDecoder.cpp:50-77,
    AbstractLJpegDecoder.cpp:65-230): SOI, DHT (class 0, id 0), SOF3 (one component), [DRI], SOS, the
    pair stream as it is (no byte stuffing in this format), EOI."""
    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    fw, fh = frame_w or w, frame_h or h
    dht = bytes([0x00]) + bytes(ncpl) + bytes(values)
    sof = bytes([prec]) + fh.to_bytes(2, "big") + fw.to_bytes(2, "big") + bytes([1, 1, 0x11, 0])
    sos = bytes([1, 1, 0x00, predictor, 0, 0])
    out = b"\xFF\xD8" + seg(0xC4, dht) + seg(0xC3, sof)
    if dri is not None:
        out += seg(0xDD, int(dri).to_bytes(2, "big"))
    return np.frombuffer(out + seg(0xDA, sos) + bytes(data) + b"\xFF\xD9", dtype=np.uint8).copy()


def _canonical(ncpl, values):
    """value -> (code, length) by T.81 C.1/C.2 (the first occurrence of a value wins)."""
    out, code, k = {}, 0, 0
    for ln in range(1, 17):
        for _ in range(ncpl[ln - 1]):
            out.setdefault(values[k], (code, ln))
            code += 1
            k += 1
        code <<= 1
    return out


def make_nikon_split(img_top, sel, pup, rows_after, seed):
    """A Nikon stream WITH a split: rows of `img_top` coded with nikon_tree[sel] like
    make_nikon, then `rows_after` rows of random symbols of nikon_tree[sel + 1] in the
    NikonLASDecompressor format (code, then len - shl bits; NikonDecompressor.cpp:315-377).
    Differential-fuzz input: the second part is not the encoding of a chosen image."""
    h, w = img_top.shape
    t1 = _canonical(*port.nikon_tree(sel))
    n2, v2 = port.nikon_tree(sel + 1)
    t2 = _canonical(n2, v2)
    bits = []

    def put(v, n):
        for k in range(n - 1, -1, -1):
            bits.append((v >> k) & 1)

    a = img_top.astype(np.int64)
    for r in range(h):
        for c in range(w):
            if c >= 2:
                p = a[r, c - 2]
            elif r >= 2:
                p = a[r - 2, c]
            else:
                p = pup[2 * r + c]
            d = int(a[r, c] - p)
            n = 0 if d == 0 else abs(d).bit_length()
            code, cl = t1[n]
            put(code, cl)
            if n:
                put(d if d > 0 else d + (1 << n) - 1, n)
    rng = np.random.default_rng(seed)
    vals = sorted(t2)
    for _ in range(rows_after * w):
        rv = vals[int(rng.integers(0, len(vals)))]
        code, cl = t2[rv]
        put(code, cl)
        if rv != 16:
            nb = (rv & 15) - (rv >> 4)
            if nb > 0:
                put(int(rng.integers(0, 1 << nb)), nb)
    while len(bits) % 8:
        bits.append(0)
    by = np.packbits(np.array(bits, dtype=np.uint8))
    return np.concatenate([by, np.zeros(16, dtype=np.uint8)])


# ---- DNG opcode lists (common/DngOpcodes.cpp:666-726; big endian) -----------------------------

def _be32(*v):
    return b"".join(struct.pack(">I", int(x) & 0xFFFFFFFF) for x in v)


def dng_opcode_list(ops):
    """ops: [(code, payload bytes[, flags])] -> the OpcodeList blob."""
    out = _be32(len(ops))
    for op in ops:
        code, payload = op[0], op[1]
        flags = op[2] if len(op) > 2 else 0
        out += _be32(code, 0x01030000, flags, len(payload)) + payload
    return np.frombuffer(out, dtype=np.uint8).copy()


def dng_roi(top, left, bottom, right):
    return _be32(top, left, bottom, right)


def dng_pixel_area(roi, first_plane=0, planes=1, row_pitch=1, col_pitch=1):
    return dng_roi(*roi) + _be32(first_plane, planes, row_pitch, col_pitch)


def dng_fix_bad_constant(value, phase=0):
    return (4, _be32(value, phase))


def dng_fix_bad_list(points=(), rects=(), phase=0):
    """points: [(y, x)], rects: [(top, left, bottom, right)] in uncropped coordinates."""
    b = _be32(phase, len(points), len(rects))
    for y, x in points:
        b += _be32(y, x)
    for r in rects:
        b += dng_roi(*r)
    return (5, b)


def dng_trim_bounds(top, left, bottom, right):
    return (6, dng_roi(top, left, bottom, right))


def dng_map_table(area, table):
    t = np.asarray(table, dtype=">u2")
    return (7, area + _be32(t.size) + t.tobytes())


def dng_map_polynomial(area, coeffs):
    return (8, area + _be32(len(coeffs) - 1) + b"".join(struct.pack(">d", float(c)) for c in coeffs))


def dng_delta(code, area, values):
    """code 10 DeltaPerRow, 11 DeltaPerColumn, 12 ScalePerRow, 13 ScalePerColumn."""
    v = np.asarray(values, dtype=">f4")
    return (code, area + _be32(v.size) + v.tobytes())
