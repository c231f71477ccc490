"""Byte-level mutation fuzz of whole lossless-JPEG tiles (markers + entropy-coded data) and of
packed strips: the oracle against the compiled reference on CORRUPT input -- same success /
exception class, and the same pixels whenever the decode still succeeds (the reference-side
driver hands the image back only then).  (The happy path is
covered by tests/test_oracle_vs_ref.py; this is the unhappy one: bad Huffman codes, early
markers, truncated streams, broken SOF / DHT / SOS fields.)"""
import numpy as np
import pytest

from oracle import port, ref, synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libref.so not built")


def _mutate(rng, blob, lo=0):
    blob = blob.copy()
    for _ in range(int(rng.integers(1, 4))):
        kind = int(rng.integers(0, 5))
        i = int(rng.integers(lo, blob.size))
        if kind == 0:
            blob[i] = rng.integers(0, 256)
        elif kind == 1:
            blob[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 2:
            blob[i] = 0xFF
        elif kind == 3 and blob.size > lo + 8:
            blob = blob[:int(rng.integers(lo + 4, blob.size))].copy()
        else:
            j = int(rng.integers(lo, blob.size))
            blob[i], blob[j] = blob[j], blob[i]
    return blob


def _both(fn_ref, fn_port, shape_img):
    out = []
    for fn in (fn_ref, fn_port):
        im = shape_img.copy()
        try:
            fn(im)
            out.append(("ok", im))
        except Exception as ex:   # noqa: BLE001
            out.append((type(ex).__name__, im))
    return out


@pytest.mark.parametrize("seed", range(150))
def test_mutated_ljpeg_tile(seed):
    rng = np.random.default_rng(7000 + seed)
    ncomp = int(rng.choice([1, 2, 3, 4]))
    tw, th = 8 * ncomp * int(rng.integers(1, 5)), int(rng.integers(2, 12))
    img = synth.image_model(tw, th, seed, bits=14)
    two_tabs = ncomp >= 2 and bool(rng.integers(0, 2))
    t = synth.make_dng_ljpeg(img, tw, th, ncomp=ncomp, tabs=synth.default_tables(2 if two_tabs else 1),
                             tab_of_comp=[c % 2 if two_tabs else 0 for c in range(ncomp)],
                             restart_rows=int(rng.choice([0, 0, 1, 2])))
    blob = _mutate(rng, t.blob)
    base = port.new_image(tw, th)
    (ra, ia), (rb, ib) = _both(lambda im: ref.ljpeg_decode(blob, im, tw, 1, (0, 0), (tw, th), (tw, th)),
                               lambda im: port.ljpeg_decode(blob, im, tw, 1, (0, 0), (tw, th), (tw, th)), base)
    assert ra == rb, (ra, rb)
    if ra == "ok":
        assert np.array_equal(ia, ib)


@pytest.mark.parametrize("seed", range(60))
def test_truncated_or_short_packed_strips(seed):
    rng = np.random.default_rng(8000 + seed)
    bps = int(rng.choice([8, 10, 12, 14, 16]))
    order = int(rng.integers(0, 4))
    w = 8 * int(rng.integers(1, 30))
    h = int(rng.integers(1, 9))
    in_pitch = w * bps // 8 + int(rng.choice([0, 0, 3]))
    n = in_pitch * h
    data = synth.lcg_bytes(n, seed)
    cut = int(rng.integers(max(0, n - 2 * in_pitch), n + 1))
    data = data[:cut].copy()
    base = port.new_image(w, h)
    (ra, ia), (rb, ib) = _both(lambda im: ref.unpack(data, im, w, 1, (0, 0, w, h), in_pitch, bps, order),
                               lambda im: port.unpack(data, im, w, 1, (0, 0, w, h), in_pitch, bps, order), base)
    assert ra == rb, (ra, rb)
    if ra == "ok":
        assert np.array_equal(ia, ib)
