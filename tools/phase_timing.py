"""Per-phase cycle accounting of k2_fused_kernel (profiling only).

    python tools/phase_timing.py build          # here: builds tools/_timing/librawspeed_b200_timing.so
    RSB200_LIB=tools/_timing/librawspeed_b200_timing.so python tools/phase_timing.py run [ljpeg|ljpeg2]

The timing build adds clock64() reads by thread 0 of every CTA at the phase
boundaries (after the barrier that ends the phase), summed over all CTAs."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "_timing", "librawspeed_b200_timing.so")
NAMES = ["wait TMA", "B unstuff", "C sync", "D prefix + write pass", "E1 prefix sums",
         "E2 row constants", "E3 stores", "E4 carry/keep", "chunk carry"]

if sys.argv[1] == "build":
    sys.path.insert(0, ROOT)
    from rawspeed_b200 import build as b
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-DRSB200_PHASE_TIMING", "-o", OUT,
                          os.path.join(ROOT, "rawspeed_b200", "csrc", "rsb200.cu")], cwd=ROOT)
    print(OUT)
else:
    sys.path.insert(0, ROOT)
    os.environ.setdefault("RSB200_LIB", OUT)
    sys.argv = [sys.argv[0]] + (sys.argv[2:] or ["ljpeg"]) + ["3"]
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import runpy
    from rawspeed_b200 import _abi
    L = _abi.load()
    L.rsb200_debug_phase_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    runpy.run_path(os.path.join(ROOT, "tools", "prof_workload.py"), run_name="__main__")
    buf = (ctypes.c_ulonglong * 16)()
    L.rsb200_debug_phase_cycles(buf, 1)
    tot = sum(buf) or 1
    for i, n in enumerate(NAMES):
        print("%-24s %6.2f %%  %12d cycles" % (n, 100.0 * buf[i] / tot, buf[i]))
