"""The S0/S1 seam: the reference's own consumer API -- RawParser(Buffer).getDecoder()
-> checkSupport(empty CameraMetaData) -> decodeRaw() -- over whole synthetic FILES, once through
the unmodified reference (oracle/_ref/libref_full.so: all 86 translation units of
src/librawspeed) and once through the reference with the four hot-path method bodies replaced by
calls into the rawspeed_b200 C ABI (oracle/_ref/libdropin.so: the same objects, the four
decompressor units compiled with the replaced method renamed, plus
rawspeed_b200/csrc/dropin/dropin_bodies.cpp).  oracle/Makefile.dropin builds both where
/root/reference exists; the GPU box uses the prebuilt libraries.

CPU: both libraries load and export the driver; the unmodified build decodes the synthetic files
and agrees with the oracle (this pins the file writer and the full build).  GPU: the drop-in
build gives the same bytes over the whole uncropped RawImage."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import port, synth
import dngfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "..", "oracle", "_ref")


def _lib(name):
    path = os.path.join(REFDIR, name)
    if os.path.exists("/root/reference/src/librawspeed/decoders/DngDecoder.cpp"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "..", "oracle"), "-f", "Makefile.dropin",
                               "-j8"], stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference)" % name)
    lib = C.CDLL(path)
    lib.rs_file_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_int32),
                                   C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    return lib


_OUT = {}


def decode_file(lib, file_bytes, threads=4):
    info = (C.c_int32 * 8)()
    err = C.create_string_buffer(512)
    if "buf" not in _OUT:
        _OUT["buf"] = np.zeros(256 << 20, dtype=np.uint8)
    out = _OUT["buf"]
    f = np.ascontiguousarray(file_bytes)
    rc = lib.rs_file_decode(f.ctypes.data, f.size, out.ctypes.data, out.size, info, err, 512, threads, 0, 0)
    if rc != 0:
        raise RuntimeError("rc %d: %s" % (rc, err.value.decode(errors="replace")))
    w, h, cpp, pitch = info[0], info[1], info[2], info[3]
    return out[:pitch * h].view(np.uint16).reshape(h, pitch // 2).copy(), (w, h, cpp, pitch, info[6])


def ljpeg_dng(w, h, tile_w, tile_h, seed, **kw):
    img = synth.image_model(w, h, seed)
    t = synth.make_dng_ljpeg(img, tile_w, tile_h, **kw)
    return img, dngfile.make_dng_tiles(w, h, 14, tile_w, tile_h, t.blob, t.offsets, t.lengths)


def packed_dng(w, h, bps, rows_per_strip, seed):
    data, pitch = synth.packed_frame(w, h, bps, seed=seed)
    return data, pitch, dngfile.make_dng_strips(w, h, bps, rows_per_strip, data, pitch)


def test_reference_build_decodes_the_synthetic_files():
    ref = _lib("libref_full.so")
    img, f = ljpeg_dng(600, 200, 256, 64, 7)
    got, (w, h, cpp, pitch, nerr) = decode_file(ref, f)
    assert (w, h, cpp, nerr) == (600, 200, 1, 0)
    assert np.array_equal(got[:, :600], img)
    data, pitch_in, f = packed_dng(512, 96, 12, 32, 3)
    got, (w, h, cpp, pitch, nerr) = decode_file(ref, f)
    want = port.new_image(512, 96)
    port.unpack(data, want, 512, 1, (0, 0, 512, 96), pitch_in, 12, port.MSB)   # DNG: not 8/16/32 bit -> big endian
    assert nerr == 0 and np.array_equal(got[:, :512], want[:, :512])


def test_dropin_library_links_and_loads():
    """Every caller of the four methods inside the reference binds to the replacement bodies
    (the library is linked with --no-undefined) and the library loads without a GPU."""
    drop = _lib("libdropin.so")
    assert hasattr(drop, "rs_file_decode")
    out = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(REFDIR, "libdropin.so")],
                         capture_output=True, text=True).stdout
    for sym in ("rawspeed::LJpegDecompressor::decode() const",
                "rawspeed::LJpegDecompressor::decode_cpu() const",
                "rawspeed::AbstractDngDecompressor::decompress() const",
                "rawspeed::UncompressedDecompressor::readUncompressedRaw()",
                "rawspeed::UncompressedDecompressor::readUncompressedRaw_cpu()"):
        assert sym in out, sym


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ljpeg_tiles", "ljpeg_ragged", "ljpeg_restart", "ljpeg_big", "packed12", "u16"])
def test_dropin_gives_the_reference_image(case):
    ref, drop = _lib("libref_full.so"), _lib("libdropin.so")
    if case == "ljpeg_tiles":
        _, f = ljpeg_dng(1024, 512, 256, 256, 11)
    elif case == "ljpeg_ragged":
        _, f = ljpeg_dng(600, 200, 256, 64, 7)
    elif case == "ljpeg_restart":
        _, f = ljpeg_dng(640, 192, 320, 96, 13, restart_rows=8)
    elif case == "ljpeg_big":
        _, f = ljpeg_dng(4096, 3072, 256, 256, 17)
    elif case == "packed12":
        _, _, f = packed_dng(2048, 300, 12, 64, 5)
    else:
        _, _, f = packed_dng(1024, 200, 16, 200, 9)
    want, wi = decode_file(ref, f)
    got, gi = decode_file(drop, f)
    assert wi == gi
    assert np.array_equal(got, want)   # whole uncropped buffer, padding included


@pytest.mark.gpu
def test_dropin_reports_a_corrupt_tile_like_the_reference():
    """One tile with an unassigned Huffman code: AbstractDngDecompressor::decompress() gives up
    with RawDecoderException "Too many errors encountered" (AbstractDngDecompressor.cpp:246-251);
    the same exception class and message head through the drop-in."""
    ref, drop = _lib("libref_full.so"), _lib("libdropin.so")
    img = synth.image_model(512, 128, 23, wild=True)
    t = synth.make_dng_ljpeg(img, 256, 64)
    blob = t.blob.copy()
    p = int(t.offsets[1]) + 400
    blob[p:p + 9] = [0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFE]   # an unassigned code in tile 1
    f = dngfile.make_dng_tiles(512, 128, 14, 256, 64, blob, t.offsets, t.lengths)
    for lib in (ref, drop):
        with pytest.raises(RuntimeError, match="rc 1: .*Too many errors"):
            decode_file(lib, f)
