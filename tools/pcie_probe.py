"""What the PCIe link of this box does with pinned buffers: H2D alone, D2H alone, both at once
(two streams), for a few copy sizes.  Tooling for the e2e section of bench.py: the pipelined
host-buffer run cannot be faster than the concurrent figure printed here."""
import json
import sys
import time

import torch


def run(nbytes, reps, mode):
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()

    def go():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return nbytes / dt / 1e9


def main():
    out = {}
    for mb in (4, 8, 64, 512):
        n = mb << 20
        reps = max(3, 2048 // mb // 4)
        out["%d MB" % mb] = {m: round(run(n, reps, m), 2) for m in ("h2d", "d2h", "both")}
        print(mb, "MB", out["%d MB" % mb], "(GB/s per direction)", flush=True)
    json.dump(out, sys.stdout)
    print()


if __name__ == "__main__":
    main()
