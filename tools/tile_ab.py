"""A/B timing of the LJPEG tile decode (development tool, not the benchmark): configs[2] (one
8256x5504 DNG frame, 726 tiles) and batches of it, through the variants of the plan:

    fused            round-1 k2_fused_kernel (RSB200_LJPEG_PATH=fused)
    tile R=1 / R=2   k2_tile_kernel<R> with the plan's parameters or overrides
    thread           K2C + K2T (batches only)

Every variant is checked bit for bit against the encoder's input before it is timed.
Usage: python tools/tile_ab.py [frames,frames,...]     (default 1,8,32)
       TILE_AB_PHASES=1 with RSB200_LIB=<lib built with -DRSB200_PHASE_TIMING>: per-phase shares."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = [
    ("fused", dict(RSB200_LJPEG_PATH="fused")),
    ("tile_r1", dict(RSB200_TILE_R="1")),
    ("tile_r1_pre352", dict(RSB200_TILE_R="1", RSB200_TILE_PREROLL="352")),
    ("tile_r1_np160", dict(RSB200_TILE_R="1", RSB200_TILE_NPIECES="160")),
    ("tile_r2", dict(RSB200_TILE_R="2")),
    ("tile_r2_pre0", dict(RSB200_TILE_R="2", RSB200_TILE_PREROLL="0")),
    ("tile_r2_pre512", dict(RSB200_TILE_R="2", RSB200_TILE_PREROLL="512")),
    ("tile_r2_np320", dict(RSB200_TILE_R="2", RSB200_TILE_NPIECES="320")),
]
KEYS = ("RSB200_LJPEG_PATH", "RSB200_TILE_R", "RSB200_TILE_PREROLL", "RSB200_TILE_NPIECES")
PHASES = ["wait TMA", "B unstuff(rest)", "C sync", "D decode", "finish", "E1/E2 sums", "E3 stores",
          "E4 carry", "chunk carry", "B1 pass1", "B2 vote/marker", "B3 scan", "B4 pass2a", "B5 pass2b", "B6 barrier"]


def main():
    import numpy as np
    import torch
    import rawspeed_b200 as rs
    from rawspeed_b200 import _abi
    from oracle import port, synth  # synthetic inputs only
    from helpers import dng_ljpeg_scans

    frames = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,8,32").split(",")]
    only = os.environ.get("TILE_AB_ONLY")
    ctx = rs.Context(0)
    W, H = 8256, 5504
    img = synth.image_model(W, H, 12345)
    out_pitch = rs.image_pitch(W)
    t = synth.make_dng_ljpeg(img, 256, 256)
    tabs, scans = dng_ljpeg_scans(t, out_pitch)
    fb = (t.blob.size + 255) // 256 * 256
    ob = (H * out_pitch + 255) // 256 * 256
    L = _abi.load()
    phases = os.environ.get("TILE_AB_PHASES") == "1" and hasattr(L, "rsb200_debug_tile_phase_cycles")
    if phases:
        L.rsb200_debug_tile_phase_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]

    def timeit(fn, reps, warm):
        for _ in range(warm):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    res = {}
    for NB in frames:
        d_in = torch.zeros(NB * fb + 64, dtype=torch.uint8, device="cuda")
        blob_t = torch.from_numpy(t.blob).cuda()
        scans_b = []
        for f in range(NB):
            d_in[f * fb:f * fb + t.blob.size] = blob_t
            for s0 in scans:
                s1 = rs.LJpegScan.from_buffer_copy(s0)
                s1.in_offset = s0.in_offset + f * fb
                s1.out_offset = s0.out_offset + f * ob
                scans_b.append(s1)
        d_out = torch.zeros(NB * ob, dtype=torch.uint8, device="cuda")
        variants = list(VARIANTS) + ([("thread", dict(RSB200_LJPEG_PATH="thread"))] if NB >= 8 else [])
        for name, env in variants:
            if only and name not in only.split(","):
                continue
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            plan = rs.ljpeg_plan(ctx, tabs.tabs, scans_b)
            d_out.zero_()
            plan.run((d_in.data_ptr(), NB * fb), d_out)
            st = plan.results()
            ok = all(s == 0 for s, _ in st)
            for f in sorted({0, NB - 1}):
                g = d_out[f * ob:f * ob + H * out_pitch].cpu().numpy().view(np.uint16).reshape(H, out_pitch // 2)
                ok = ok and bool(np.array_equal(g[:, :W], img))
            if phases:
                buf = (ctypes.c_ulonglong * 16)()
                L.rsb200_debug_tile_phase_cycles(buf, 1)
            reps = 20 if NB == 1 else (5 if NB <= 8 else 3)
            ms = timeit(lambda: plan.run((d_in.data_ptr(), NB * fb), d_out), reps, 3 if NB == 1 else 1)
            r = {"ms": round(ms, 4), "exact": ok, "GPix/s": round(NB * W * H / ms / 1e6, 1),
                 "launches": plan.launches}
            if phases and name.startswith("tile"):
                buf = (ctypes.c_ulonglong * 16)()
                L.rsb200_debug_tile_phase_cycles(buf, 1)
                tot = sum(buf) or 1
                r["phases_pct"] = {PHASES[i]: round(100.0 * buf[i] / tot, 1) for i in range(len(PHASES))}
            res["%dx %s" % (NB, name)] = r
            print("%3d frames  %-16s %9.4f ms  %8.1f GPix/s  exact=%s" % (NB, name, ms, r["GPix/s"], ok),
                  flush=True)
            if "phases_pct" in r:
                print("            " + json.dumps(r["phases_pct"]), flush=True)
            del plan
        del d_in, d_out
    for k in KEYS:
        os.environ.pop(k, None)
    print("TILE_AB " + json.dumps(res))


if __name__ == "__main__":
    main()
