"""Small single-plan driver for ncu captures (profiling only, not a benchmark).
    python tools/prof_workload.py ljpeg|ljpeg2|unpack|cr2|cr2_4 [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rawspeed_b200 as rs  # noqa: E402
from oracle import port, synth  # noqa: E402  (synthetic inputs only)
from helpers import dng_ljpeg_scans, TableSet  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ljpeg"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = rs.Context(0)
W, H = 8256, 5504
if what in ("ljpeg", "ljpeg2"):
    img = synth.image_model(W, H, 12345)
    kw = {}
    if what == "ljpeg2":
        kw = dict(tabs=synth.default_tables(2), tab_of_comp=[0, 1])
    t = synth.make_dng_ljpeg(img, 256, 256, **kw)
    tabs, scans = dng_ljpeg_scans(t, rs.image_pitch(W))
    plan = rs.ljpeg_plan(ctx, tabs.tabs, scans)
    d_in = torch.zeros(t.blob.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:t.blob.size] = torch.from_numpy(t.blob)
    d_out = torch.zeros(H * rs.image_pitch(W), dtype=torch.uint8, device="cuda")
    args = ((d_in.data_ptr(), t.blob.size), d_out)
elif what == "unpack":
    data, pitch = synth.packed_frame(W, H, 14, seed=2)
    F = 4
    j = []
    fb = (pitch * H + 255) // 256 * 256
    ob = (rs.image_pitch(W) * H + 255) // 256 * 256
    for f in range(F):
        u = rs.UnpackJob()
        u.in_offset, u.in_size, u.out_offset = f * fb, pitch * H, f * ob
        u.out_pitch, u.rows, u.samples = rs.image_pitch(W), H, W
        u.in_pitch, u.bps, u.order = pitch, 14, rs.MSB
        j.append(u)
    plan = rs.unpack_plan(ctx, j)
    d_in = torch.zeros(F * fb, dtype=torch.uint8, device="cuda")
    for f in range(F):
        d_in[f * fb:f * fb + data.size] = torch.from_numpy(data)
    d_out = torch.zeros(F * ob, dtype=torch.uint8, device="cuda")
    args = (d_in, d_out)
elif what in ("cr2", "cr2_4"):
    from test_gpu_cr2 import cr2_job
    cw, ch = 6720, 4480
    cimg = port.new_image(cw, ch)
    cimg[:, :cw] = synth.image_model(cw, ch, 4)
    fmt, frame = ((2, 1, 1), (3360, 4480)) if what == "cr2" else ((4, 1, 1), (1680, 4480))
    blob = port.cr2_encode(cimg, cw, fmt, frame, (3, 2240, 2240), 14, synth.default_tables(2),
                           [0, 1, 0, 1][:fmt[0]])
    ts = TableSet()
    job = cr2_job(blob, cw, ch, fmt, (3, 2240, 2240), cimg.shape[1] * 2, ts)
    plan = rs.cr2_plan(ctx, ts.tabs, [job])
    d_in = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:blob.size] = torch.from_numpy(blob)
    d_out = torch.zeros(cimg.size * 2, dtype=torch.uint8, device="cuda")
    args = ((d_in.data_ptr(), blob.size), d_out)
else:
    raise SystemExit("unknown workload")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
torch.cuda.synchronize()
for i in range(reps):
    ev[i].record()
    plan.run(*args)
ev[reps].record()
torch.cuda.synchronize()
print(what, "ms per run:", [round(ev[i].elapsed_time(ev[i + 1]), 4) for i in range(reps)])
