"""A/B timing of the LJPEG decode kernels for library variants (development tool, not
the benchmark).

    python tools/ab_ljpeg.py build NAME [-DFLAG ...]   # here: tools/_ab/NAME.so
    python tools/ab_ljpeg.py run NAME [NAME ...]       # on the GPU box: one process per variant
    python tools/ab_ljpeg.py one                       # (internal) time the library in RSB200_LIB

Each variant is timed on: one 8256x5504 DNG frame (726 LJPEG tiles), the same frame
with two Huffman tables (phase-carrying synchronisation), LJPEG batches and the
6720x4480 CR2 stream; every output is checked bit for bit against the encoder's input.

Environment: AB_FRAMES=8,32,128   batch sizes (frames of the 45 MP DNG in one plan)
             AB_PATHS=fused,thread|auto   kernel path(s) per batch (RSB200_LJPEG_PATH)
             AB_KERNELS=1         per-kernel durations of one run (torch profiler / CUPTI)
             AB_ONLY=batch        only the single-table frame and its batches (no second table, no CR2)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "_ab")


def build(name, flags):
    sys.path.insert(0, ROOT)
    from rawspeed_b200 import build as b
    os.makedirs(AB, exist_ok=True)
    out = os.path.join(AB, name + ".so")
    subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + list(flags) + ["-o", out,
                          os.path.join(ROOT, "rawspeed_b200", "csrc", "rsb200.cu"), "-ldl", "-lgomp"], cwd=ROOT)
    print(out)


def one():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rawspeed_b200 as rs
    from oracle import port, synth  # synthetic inputs only
    from helpers import dng_ljpeg_scans, TableSet

    ctx = rs.Context(0)
    W, H = 8256, 5504
    res = {}

    def timeit(fn, reps=10, warm=3):
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

    img = synth.image_model(W, H, 12345)
    out_pitch = rs.image_pitch(W)
    only_batch = os.environ.get("AB_ONLY") == "batch"
    for label, kw in (("dng1", {}), ("dng1_2tab", dict(tabs=synth.default_tables(2), tab_of_comp=[0, 1]))):
        if only_batch and label != "dng1":
            continue
        t = synth.make_dng_ljpeg(img, 256, 256, **kw)
        tabs, scans = dng_ljpeg_scans(t, out_pitch)
        plan = rs.ljpeg_plan(ctx, tabs.tabs, scans)
        d_in = torch.zeros(t.blob.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:t.blob.size] = torch.from_numpy(t.blob)
        d_out = torch.zeros(H * out_pitch, dtype=torch.uint8, device="cuda")
        plan.run((d_in.data_ptr(), t.blob.size), d_out)
        st = plan.results()
        got = d_out.cpu().numpy().view(np.uint16).reshape(H, out_pitch // 2)
        ok = bool(np.array_equal(got[:, :W], img)) and all(s == 0 for s, _ in st)
        ms = timeit(lambda: plan.run((d_in.data_ptr(), t.blob.size), d_out))
        res[label] = {"ms": round(ms, 4), "exact": ok}
        if label == "dng1":
            import copy
            fb = (t.blob.size + 255) // 256 * 256
            ob = (H * out_pitch + 255) // 256 * 256
            for NB in [int(x) for x in os.environ.get("AB_FRAMES", "8").split(",")]:
                d_inb = torch.zeros(NB * fb + 64, dtype=torch.uint8, device="cuda")
                scans_b = []
                for f in range(NB):
                    d_inb[f * fb:f * fb + t.blob.size] = d_in[:t.blob.size]
                    for s0 in scans:
                        s1 = rs.LJpegScan.from_buffer_copy(s0)
                        s1.in_offset = s0.in_offset + f * fb
                        s1.out_offset = s0.out_offset + f * ob
                        scans_b.append(s1)
                d_outb = torch.zeros(NB * ob, dtype=torch.uint8, device="cuda")
                for path in os.environ.get("AB_PATHS", "auto").split(","):
                    os.environ["RSB200_LJPEG_PATH"] = path
                    planb = rs.ljpeg_plan(ctx, tabs.tabs, scans_b)
                    d_outb.zero_()
                    planb.run((d_inb.data_ptr(), NB * fb), d_outb)
                    stb = planb.results()
                    okb = all(s == 0 for s, _ in stb)
                    for f in (0, NB - 1):
                        gb = d_outb[f * ob:f * ob + H * out_pitch].cpu().numpy().view(np.uint16).reshape(H, out_pitch // 2)
                        okb = okb and bool(np.array_equal(gb[:, :W], img))
                    msb = timeit(lambda: planb.run((d_inb.data_ptr(), NB * fb), d_outb), reps=3, warm=1)
                    if os.environ.get("AB_KERNELS"):
                        from torch.profiler import profile, ProfilerActivity
                        with profile(activities=[ProfilerActivity.CUDA]) as prof:
                            planb.run((d_inb.data_ptr(), NB * fb), d_outb)
                            torch.cuda.synchronize()
                        for ev in prof.key_averages():
                            if "k2_" in ev.key:
                                print("KERNEL dng%d_%s %s %.3f ms" % (NB, path, ev.key.split("(")[0],
                                                                      ev.device_time_total / 1e3))
                    res["dng%d_%s" % (NB, path)] = {"ms": round(msb, 4), "exact": okb,
                                                    "launches": planb.launches,
                                                    "GPix/s": round(NB * W * H / msb / 1e6, 1)}
                    del planb
                os.environ.pop("RSB200_LJPEG_PATH", None)
                del d_inb, d_outb
        del plan, d_in, d_out
    from test_gpu_cr2 import cr2_job
    cw, ch = 6720, 4480
    cimg = port.new_image(cw, ch)
    cimg[:, :cw] = synth.image_model(cw, ch, 4)
    # one shared table (positions-only synchronisation) and two tables
    for label, hts, sel in (("cr2_1tab", synth.default_tables(1), [0, 0]),
                            ("cr2_2tab", synth.default_tables(2), [0, 1])):
        if only_batch:
            break
        blob = port.cr2_encode(cimg, cw, (2, 1, 1), (3360, 4480), (3, 2240, 2240), 14, hts, sel)
        ts = TableSet()
        job = cr2_job(blob, cw, ch, (2, 1, 1), (3, 2240, 2240), cimg.shape[1] * 2, ts)
        plan = rs.cr2_plan(ctx, ts.tabs, [job])
        d_in = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:blob.size] = torch.from_numpy(blob)
        d_out = torch.zeros(cimg.size * 2, dtype=torch.uint8, device="cuda")
        plan.run((d_in.data_ptr(), blob.size), d_out)
        st = plan.results()
        got = d_out.cpu().numpy().view(np.uint16).reshape(cimg.shape)
        ok = bool(np.array_equal(got[:, :cw], cimg[:, :cw])) and st[0][0] == 0
        ms = timeit(lambda: plan.run((d_in.data_ptr(), blob.size), d_out), reps=5, warm=2)
        res[label] = {"ms": round(ms, 4), "exact": ok}
        del plan, d_in, d_out
    print("AB " + os.path.basename(os.environ.get("RSB200_LIB", "default")) + " " + json.dumps(res))


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "build":
        build(sys.argv[2], sys.argv[3:])
    elif cmd == "run":
        for name in sys.argv[2:]:
            env = dict(os.environ)
            if name != "default":
                env["RSB200_LIB"] = os.path.join(AB, name + ".so")
            subprocess.call([sys.executable, os.path.abspath(__file__), "one"], env=env)
    else:
        one()
