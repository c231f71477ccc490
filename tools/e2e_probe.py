"""Host-buffer runs (rsb200_plan_run_host, pinned buffers) of 1 and N configs[2] frames with a
timeline of the pipeline (RSB200_PIPE_TRACE) and a few group sizes.  Tooling, not a benchmark."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rawspeed_b200 as rs  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic inputs only)
from helpers import dng_ljpeg_scans  # noqa: E402

W, H = 8256, 5504


def main():
    nframes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,16").split(",")]
    ctx = rs.Context(0)
    img = synth.image_model(W, H, 12345)
    pitch = rs.image_pitch(W)
    t = synth.make_dng_ljpeg(img, 256, 256)
    tabs, scans = dng_ljpeg_scans(t, pitch)
    fb = (t.blob.size + 255) // 256 * 256
    ob = (H * pitch + 255) // 256 * 256
    for n in nframes:
        h_in = torch.zeros(n * fb + 64, dtype=torch.uint8, pin_memory=True)
        h_out = torch.zeros(n * ob, dtype=torch.uint8, pin_memory=True)
        sc = []
        for f in range(n):
            h_in[f * fb:f * fb + t.blob.size] = torch.from_numpy(t.blob)
            for s0 in scans:
                s1 = rs.LJpegScan.from_buffer_copy(s0)
                s1.in_offset = s0.in_offset + f * fb
                s1.out_offset = s0.out_offset + f * ob
                sc.append(s1)
        for gmb in os.environ.get("E2E_GROUPS", "0").split(","):
            if gmb != "0":
                os.environ["RSB200_GROUP_MB"] = gmb
            else:
                os.environ.pop("RSB200_GROUP_MB", None)
            plan = rs.ljpeg_plan(ctx, tabs.tabs, sc)
            a_in, a_out = h_in.numpy()[:n * fb], h_out.numpy()
            for _ in range(2):
                plan.run_host(a_in, a_out)
            torch.cuda.synchronize()
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                plan.run_host(a_in, a_out)
            ms = (time.perf_counter() - t0) / reps * 1e3
            ok = bool(np.array_equal(a_out[(n - 1) * ob:(n - 1) * ob + H * pitch].view(np.uint16).reshape(H, -1)[:, :W], img))
            print("%3d frames  group %s MB  %8.3f ms/run  %7.2f GPix/s  D2H %.1f GB/s  exact=%s  [%s]" % (
                n, gmb, ms, n * W * H / ms / 1e6, n * ob / ms / 1e6, ok, plan.kernels), flush=True)
            if os.environ.get("E2E_TRACE"):
                os.environ["RSB200_PIPE_TRACE"] = "1"
                plan.run_host(a_in, a_out)
                os.environ.pop("RSB200_PIPE_TRACE")
            del plan
        del h_in, h_out


if __name__ == "__main__":
    main()
