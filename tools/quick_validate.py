"""Fast GPU check of the kernels developed against CPU replays (K9 scaling, K10 DNG
opcodes, K11 bad pixels, K12 table lookup, Panasonic V4): every scenario of their test files
through the C++ host mirror (-> C ABI -> kernel; no torch import, the library owns the device
buffers), compared with the oracle.  Prints one line per case and a summary; exit code =
number of failing cases.  Seconds, not minutes:

    gpurun --timeout 120 -- 'python tools/quick_validate.py > gpurun_out/quick_validate.log 2>&1; tail -40 gpurun_out/quick_validate.log'
"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import port, synth          # noqa: E402  (the checker)
from rawspeed_b200 import host          # noqa: E402

T0 = time.time()
fails, passed = [], 0


def check(name, fn):
    global passed
    t = time.time()
    try:
        ok = bool(fn())
    except Exception:   # noqa: BLE001
        ok = False
        traceback.print_exc()
    print("%-58s %s  %.2fs" % (name, "ok" if ok else "FAIL", time.time() - t), flush=True)
    if ok:
        passed += 1
    else:
        fails.append(name)


def rnd(w, h, cpp, seed, hi=65536):
    rng = np.random.default_rng(seed)
    a = port.new_image(w, h, cpp)
    a[:, :] = rng.integers(0, hi, size=a.shape, dtype=np.uint16)
    return a


# ---- K12 table lookup --------------------------------------------------------------------------
from test_oracle_lookup import CASES as LUT_CASES, curve   # noqa: E402
for k, (w, h, cpp, crop, ncurve) in enumerate(LUT_CASES):
    for dither in (False, True):
        def f(w=w, h=h, cpp=cpp, ncurve=ncurve, dither=dither, k=k):
            a = rnd(w, h, cpp, k)
            want = a.copy()
            cv = curve(ncurve, 10 + k)
            port.sixteen_bit_lookup(want, w, cpp, port.build_table(cv, dither), dither)
            host.sixteen_bit_lookup(a, w, cpp, cv, dither)
            return np.array_equal(a[:, :w * cpp], want[:, :w * cpp])
        check("K12 lookup case %d dither=%d" % (k, dither), f)

# ---- K9 scaling --------------------------------------------------------------------------------
SCALE = [(64, 16, (0, 0, 64, 16), dict(black_sep=[256] * 4, white=16383)),
         (70, 11, (3, 1, 61, 9), dict(black_sep=[60, 64, 68, 72], white=4095)),
         (37, 9, (2, 3, 30, 5), dict(black_sep=[1000, 1010, 990, 1024], white=15000)),
         (1000, 6, (8, 0, 980, 6), dict(black_sep=[512] * 4, white=16383)),
         (33, 7, (5, 2, 20, 4), dict(black_sep=[100, 200, 300, 400], white=1023)),
         (530, 9, (11, 2, 515, 6), dict(black_sep=[10, 20, 30, 40], white=900)),
         (96, 40, (16, 8, 80, 32), dict(white=15000, areas=[(1, 0, 16)])),
         (8256, 37, (8, 1, 8240, 35), dict(black_sep=[1008, 1010, 1009, 1011], white=16383))]
for k, (w, h, crop, kw) in enumerate(SCALE):
    for dither in (True, False):
        def f(w=w, h=h, crop=crop, kw=kw, dither=dither, k=k):
            a = rnd(w, h, 1, 40 + k, 16384)
            a[:, :16] = 512
            want = a.copy()
            rw = port.scale_black_white(want, w, crop, dither=dither, **kw)
            rg = host.scale_black_white(a, w, crop, dither=dither, **kw)
            return rg == rw and np.array_equal(a[:, :w], want[:, :w])
        check("K9 scale case %d dither=%d" % (k, dither), f)

# ---- K10 DNG opcodes ---------------------------------------------------------------------------
from test_oracle_dngopcodes import scenarios as dng_scenarios   # noqa: E402
for name, img, w, cpp, crop, blob in dng_scenarios():
    def f(img=img, w=w, cpp=cpp, crop=crop, blob=blob):
        want, got = img.copy(), img.copy()
        try:
            wr = port.dng_opcodes(want, w, cpp, crop, blob)
            werr = None
        except Exception as ex:   # noqa: BLE001
            werr, wr = ex, tuple(port.dng_opcodes.partial[:2])
        try:
            gr = host.dng_opcodes(got, w, cpp, crop, blob)
            gerr = None
        except Exception as ex:   # noqa: BLE001
            gerr, gr = ex, tuple(host.dng_opcodes.partial)
        return (np.array_equal(got, want) and tuple(gr) == tuple(wr) and
                type(gerr).__name__ == type(werr).__name__)
    check("K10 opcodes %s" % name, f)

# ---- K11 bad pixels ----------------------------------------------------------------------------
from test_oracle_badpixels import scenarios as bad_scenarios, pos   # noqa: E402
for k, (name, w, h, cpp, cfa, points) in enumerate(bad_scenarios()):
    if cpp != 1:
        continue

    def f(w=w, h=h, cfa=cfa, points=points, k=k):
        a = rnd(w, h, 1, k)
        want = a.copy()
        port.fix_bad_pixels(want, w, 1, pos(points), cfa)
        host.fix_bad_pixels(a, w, 1, pos(points), cfa)
        return np.array_equal(a[:, :w], want[:, :w])
    check("K11 bad pixels %s" % name, f)

# ---- Panasonic V4 ------------------------------------------------------------------------------
from test_pana4_emu import v4_payload   # noqa: E402
for w, h, split, zero_ok in [(14, 1, 0, True), (28, 3, 0, False), (1400, 25, 0x1FF8, True),
                             (2800, 13, 0x1FF8, False), (1414, 9, 0, False), (4200, 6, 0x2008, False),
                             (1428, 11, 0x4000, False), (1428, 30, 0x1235, False), (5600, 12, 0x3FFF, False)]:
    def f(w=w, h=h, split=split, zero_ok=zero_ok):
        data = v4_payload(w, h, split, w + h)
        want, got = port.new_image(w, h), port.new_image(w, h)
        zw = port.panasonic_v4(want, w, data, zero_ok, split)
        zg = host.panasonic_v4(got, w, data, zero_ok, split)
        return zg == zw and np.array_equal(got[:, :w], want[:, :w])
    check("V4 %dx%d split=0x%x zero_ok=%d" % (w, h, split, zero_ok), f)

print("\n%d passed, %d failed, %.1f s" % (passed, len(fails), time.time() - T0))
for n in fails:
    print("FAILED:", n)
sys.exit(min(len(fails), 100))
