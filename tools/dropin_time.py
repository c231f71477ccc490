"""Where does the drop-in call spend its time?  One 8256x5504 DNG LJPEG FILE through
RawParser -> decodeRaw() of (a) the unmodified reference (libref_full.so, all host cores) and
(b) the reference with the hot-path bodies replaced (libdropin.so); plus the pieces of (b)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_dropin as td
import dngfile
import bench
from oracle import synth

W, H = 8256, 5504
img = bench.frame_image(W, H, 12345)
t = synth.make_dng_ljpeg(img, 256, 256)
f = dngfile.make_dng_tiles(W, H, 14, 256, 256, t.blob, t.offsets, t.lengths)
ref, drop = td._lib("libref_full.so"), td._lib("libdropin.so")
ncpu = os.cpu_count() or 1


import ctypes as C
OUT = np.zeros(128 << 20, dtype=np.uint8)   # touched once: no page faults inside the timed calls


def timed(lib, threads, n=7):
    info = (C.c_int32 * 8)()
    err = C.create_string_buffer(512)
    ts = []
    for i in range(n + 2):
        t0 = time.perf_counter()
        rc = lib.rs_file_decode(f.ctypes.data, f.size, OUT.ctypes.data, OUT.size, info, err, 512, threads, 0, 0)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0, err.value
    got = OUT[:info[3] * info[1]].view(np.uint16).reshape(info[1], info[3] // 2).copy()
    ts = sorted(ts[2:])
    return ts[len(ts) // 2], got


ms_ref, want = timed(ref, ncpu)
print("reference, %d threads:      %8.2f ms/file = %8.0f MPix/s" % (ncpu, ms_ref, W * H / ms_ref / 1e3))
for th in (1, 16, ncpu):
    ms, got = timed(drop, th)
    print("drop-in, %3d host threads:   %8.2f ms/file = %8.0f MPix/s  exact=%s" % (
        th, ms, W * H / ms / 1e3, bool(np.array_equal(got, want))))
# (the numbers include RawParser, TIFF walk, RawImage::createData of 91 MB and the copy of the
#  decoded image into the test's numpy buffer, for both libraries alike)
