"""Torch-free device timing of the post-decode kernels (K9 scaling, K10 DNG opcodes, K11 bad
pixels, K12 table lookup) and Panasonic V4 at full frame size: CUDA events through ctypes on
libcudart, plans through the C ABI, first run of every leg checked bit for bit against the
oracle.  A development probe (seconds on the box) -- the numbers of record come from bench.py.

    gpurun --timeout 120 -- 'python tools/quick_time.py > gpurun_out/quick_time.log 2>&1; tail -30 gpurun_out/quick_time.log'
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rawspeed_b200 as rs              # noqa: E402
from rawspeed_b200 import host          # noqa: E402
from oracle import port, synth          # noqa: E402  (the checker)

rt = None
for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
    try:
        rt = C.CDLL(name)
        break
    except OSError:
        pass
assert rt is not None, "libcudart not found"
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaEventSynchronize.argtypes = [C.c_void_p]
rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
rt.cudaFree.argtypes = [C.c_void_p]
H2D, D2H, D2D = 1, 2, 3


def ck(rc):
    assert rc == 0, "cuda error %d" % rc


def dmalloc(n):
    p = C.c_void_p()
    ck(rt.cudaMalloc(C.byref(p), n))
    return p


def timed(fn, reps):
    e0, e1 = C.c_void_p(), C.c_void_p()
    ck(rt.cudaEventCreate(C.byref(e0)))
    ck(rt.cudaEventCreate(C.byref(e1)))
    for _ in range(3):
        fn()
    ck(rt.cudaDeviceSynchronize())
    ck(rt.cudaEventRecord(e0, None))
    for _ in range(reps):
        fn()
    ck(rt.cudaEventRecord(e1, None))
    ck(rt.cudaEventSynchronize(e1))
    ms = C.c_float(0)
    ck(rt.cudaEventElapsedTime(C.byref(ms), e0, e1))
    return ms.value / reps


ctx = rs.Context(0)
PEAK = 6569.6     # MEASURED_PEAKS.json hbm_gbs of this pool (burst), GB/s
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:   # noqa: BLE001
    pass
W, H = 8256, 5504
pitch = rs.image_pitch(W)
rng = np.random.default_rng(9)
base = port.new_image(W, H)
base[:, :] = rng.integers(0, 16384, size=base.shape, dtype=np.uint16)
nbytes = base.nbytes
d_src, d_img = dmalloc(nbytes), dmalloc(nbytes)
ck(rt.cudaMemcpy(d_src, base.ctypes.data, nbytes, H2D))
results = {}


def leg(name, plan, want, reps=20, restore=True):
    ck(rt.cudaMemcpy(d_img, d_src, nbytes, D2D))
    plan.run(None, (d_img.value, nbytes), stream=0)
    got = np.empty_like(base)
    ck(rt.cudaMemcpy(got.ctypes.data, d_img, nbytes, D2H))
    exact = bool(np.array_equal(got, want))

    def step():
        if restore:
            rt.cudaMemcpy(d_img, d_src, nbytes, D2D)   # every run starts from the same pixels
        plan.run(None, (d_img.value, nbytes), stream=0)
    ms = timed(step, reps)
    ms_copy = timed(lambda: rt.cudaMemcpy(d_img, d_src, nbytes, D2D), reps) if restore else 0.0
    per = ms - ms_copy
    in_b, out_b, pixels = plan.bytes()
    results[name] = {"ms": round(per, 4), "GPix/s": round(pixels / per / 1e6, 1),
                     "GB/s": round((in_b + out_b) / per / 1e6, 1),
                     "frac_of_hbm_peak": round((in_b + out_b) / per / 1e6 / PEAK, 3), "bit_exact": exact,
                     "ms_with_restoring_copy": round(ms, 4), "ms_copy_alone": round(ms_copy, 4)}
    print("%-52s %8.4f ms  %8.1f GPix/s  %7.1f GB/s  exact=%s" % (name, per, pixels / per / 1e6,
                                                               (in_b + out_b) / per / 1e6, exact), flush=True)


t0 = time.time()
for label, black, white in (("SSE2 loop", (1008, 1010, 1009, 1011), 16383), ("plain loop", (64,) * 4, 1000)):
    j = rs.ScaleJob()
    j.offset, j.pitch, j.width, j.height, j.cpp = 0, pitch, W, H, 1
    j.crop_x, j.crop_y, j.crop_w, j.crop_h = 8, 8, W - 16, H - 16
    for i in range(4):
        j.black_separate[i] = black[i]
    j.white_point, j.dither, j.path = white, 1, 0
    want = base.copy()
    port.scale_values(want, W, (8, 8, W - 16, H - 16), black, white)
    leg("K9 scaleBlackWhite, %s, dither" % label, rs.scale_plan(ctx, [j]), want)
for dither in (False, True):
    lj = rs.LookupJob()
    lj.offset, lj.pitch, lj.width, lj.height, lj.cpp, lj.table = 0, pitch, W, H, 1, 0
    t = port.build_table(synth.sony_curve(), dither)
    want = base.copy()
    port.sixteen_bit_lookup(want, W, 1, t, dither)
    leg("K12 sixteenBitLookup, %s" % ("dithered" if dither else "plain"), rs.lookup_plan(ctx, [lj], t, dither), want)
area = synth.dng_pixel_area((0, 0, H, W))
blob = synth.dng_opcode_list([
    synth.dng_delta(12, area, rng.random(H, dtype=np.float32) + 0.5),
    synth.dng_delta(13, synth.dng_pixel_area((0, 0, H, W), 0, 1, 1, 2), rng.random(W // 2, dtype=np.float32) + 0.5),
    synth.dng_delta(10, synth.dng_pixel_area((1, 1, H, W), 0, 1, 2, 2), (rng.random(H // 2, dtype=np.float32) - 0.5) * 0.01),
    synth.dng_delta(11, area, (rng.random(W, dtype=np.float32) - 0.5) * 0.01),
    synth.dng_map_polynomial(area, [0.0, 0.8, 0.3, -0.1]),
    synth.dng_map_table(synth.dng_pixel_area((0, 1, H, W), 0, 1, 2, 2), (np.arange(65536) ^ 1).astype(np.uint16)),
    synth.dng_delta(13, synth.dng_pixel_area((8, 8, H - 8, W - 8), 0, 1, 1, 16), rng.random((W - 16 + 15) // 16, dtype=np.float32) + 0.25),
    synth.dng_delta(12, synth.dng_pixel_area((0, 0, H, W), 0, 1, 4, 1), rng.random(H // 4, dtype=np.float32) + 0.75)])
low = host.dngop_lower(base, W, 1, [0, 0, W, H], blob)
dj = rs.DngOpJob()
dj.offset, dj.pitch, dj.width, dj.height, dj.cpp, dj.is_f32 = 0, pitch, W, H, 1, 0
dj.first_op, dj.num_ops = 0, len(low["ops"])
want = base.copy()
port.dng_opcodes(want, W, 1, [0, 0, W, H], blob)
leg("K10 DngOpcodes, 8 opcodes in one pass", rs.dngop_plan(ctx, [dj], low["ops"], low["tables"], low["deltas"]), want)
n = 20000
p = ((rng.integers(0, H, n).astype(np.uint32) << 16) | rng.integers(0, W, n).astype(np.uint32))
bj = rs.BadPixJob()
bj.offset, bj.pitch, bj.width, bj.height, bj.is_cfa = 0, pitch, W, H, 1
bj.first_position, bj.num_positions, bj.prior_map = 0, n, None
want = base.copy()
port.fix_bad_pixels(want, W, 1, p, True)
leg("K11 fixBadPixels, 20000 defects", rs.badpix_plan(ctx, [bj], p), want, restore=False)
# Panasonic V4: 4 frames of 4592x3448 per launch
w, h, split = 4592 // 14 * 14, 3448, 0x2008
nb = (w * h // 14 * 16 + 0x3FFF) // 0x4000 * 0x4000
data = synth.lcg_bytes(nb, 44)
opitch = rs.image_pitch(w)
nf, fb, ob = 4, (nb + 255) // 256 * 256, (h * opitch + 255) // 256 * 256
jobs = []
for f in range(nf):
    pj = rs.PanaJob()
    pj.in_offset, pj.in_size, pj.out_offset, pj.out_pitch = f * fb, nb, f * ob, opitch
    pj.width, pj.height, pj.version, pj.bps = w, h, 4, 12
    pj.zero_is_not_bad, pj.section_split_offset = 0, split
    jobs.append(pj)
plan = rs.pana_plan(ctx, jobs)
d_in, d_out = dmalloc(nf * fb + 64), dmalloc(nf * ob)
for f in range(nf):
    ck(rt.cudaMemcpy(C.c_void_p(d_in.value + f * fb), data.ctypes.data, nb, H2D))
plan.run((d_in.value, nf * fb), (d_out.value, nf * ob), stream=0)
want = port.new_image(w, h)
zwant = port.panasonic_v4(want, w, data, False, split, cap=1 << 22)
got = np.empty((h, opitch // 2), dtype=np.uint16)
ck(rt.cudaMemcpy(got.ctypes.data, C.c_void_p(d_out.value + (nf - 1) * ob), h * opitch, D2H))
nz, zl = plan.bad_pixels(nf - 1, cap=1 << 22)
exact = bool(np.array_equal(got[:, :w], want[:, :w])) and sorted(zl) == zwant
ms = timed(lambda: plan.run((d_in.value, nf * fb), (d_out.value, nf * ob), stream=0), 20)
in_b, out_b, pixels = plan.bytes()
results["PanasonicV4 4 x %dx%d" % (w, h)] = {"ms": round(ms, 4), "GPix/s": round(pixels / ms / 1e6, 1),
                                           "GB/s": round((in_b + out_b) / ms / 1e6, 1),
                                           "frac_of_hbm_peak": round((in_b + out_b) / ms / 1e6 / PEAK, 3),
                                           "bit_exact": exact, "bad_pixels": nz}
print("%-52s %8.4f ms  %8.1f GPix/s  %7.1f GB/s  exact=%s" % ("PanasonicV4 4 frames", ms, pixels / ms / 1e6,
                                                           (in_b + out_b) / ms / 1e6, exact), flush=True)
print("QUICK_TIME " + json.dumps({"peak_GBps": PEAK, "frame": "%dx%d uint16" % (W, H), "legs": results,
                                  "timing": "CUDA events (ctypes libcudart) on the default stream, 3 warm-up + 20 timed "
                                            "runs; in-place kernels: run + restoring D2D copy minus the copy alone",
                                  "wall_s": round(time.time() - t0, 1)}))
