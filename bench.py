#!/usr/bin/env python
"""bench.py -- headline benchmark of the rawspeed_b200 hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): MPixels/s decoded (bit-exact); achieved HBM GB/s vs roofline.
Headline workload (configs[1]): 14-bit packed unpack, 8256x5504 (45 MP) frames,
a batch of --frames frames per step, inputs resident in HBM (`value`) and through
the host-buffer C-ABI call with H2D/D2H inside the timed region (`e2e`).
`others` carries the same device-timed measurement for the LJPEG configs
(configs[2] DNG tiles, configs[3] CR2).

One JSON line on stdout (rank 0).  A "step" = one pass of the hot path over one
batch of synthetic input.  Under torchrun each rank decodes its own batch (the
path shards by frame with no data-path collective -> weak scaling); the optional
NVLink output gather is timed separately (`gather`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, BPS = 8256, 5504, 14
PIX = W * H


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per frame of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, idle = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, cmax, util = float(f[0]), float(f[1]), float(f[7])
            except ValueError:
                continue
            mx.append(cmax)
            if util < 50.0:  # not under load: sampler started before the warm-up
                idle.append(clk)
                continue
            sm.append(clk)
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            sm = idle
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "samples_under_load": len(sm), "samples": len(sm) + len(idle),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------
def unpack_jobs(rs, frames, in_frame_bytes, out_frame_bytes, pitch, out_pitch, order):
    jobs = []
    for f in range(frames):
        j = rs.UnpackJob()
        j.in_offset = f * in_frame_bytes
        j.in_size = pitch * H
        j.out_offset = f * out_frame_bytes
        j.out_pitch = out_pitch
        j.row0, j.rows, j.samples, j.out_col0 = 0, H, W, 0
        j.in_pitch, j.bps, j.order = pitch, BPS, order
        jobs.append(j)
    return jobs


def align(x, a=256):
    return (x + a - 1) // a * a


def time_steps(torch, fn, steps, warmup, dist=None):
    """W untimed + K timed steps, CUDA events on the launching (current) stream,
    barrier + synchronize on both sides, max over ranks.  Returns total ms."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    torch.cuda.synchronize()
    return ms


def wall_steps(torch, fn, steps, warmup, dist=None):
    """Same contract for the host-API path (its timed region is host-driven:
    pinned H2D + kernels + D2H, synchronous); wall clock bracketed by syncs."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def cpu_reference_unpack(sample_frames=1, reps=3):
    """The reference's own CPU path on this box's host cores (bounded sample)."""
    import oracle
    from oracle import port, synth
    ncores = os.cpu_count() or 1
    data, pitch = synth.packed_frame(W, H, BPS, seed=2)
    img = port.new_image(W, H)
    if oracle.HAVE_REF:
        ref = oracle.ref
        # (a) as shipped: UncompressedDecompressor is single threaded
        ms1 = min(ref.unpack(data, img, W, 1, (0, 0, W, H), pitch, BPS, port.MSB, reps=1)
                  for _ in range(reps))
        # (b) the reference's best OpenMP shape: rows split into one strip per core,
        #     fanned out by its own AbstractDngDecompressor (compression 1)
        th = (H + ncores - 1) // ncores
        nt = (H + th - 1) // th
        offs = [n * th * pitch for n in range(nt)]
        lens = [min(th, H - n * th) * pitch for n in range(nt)]
        # tile height th: last tile shorter; AbstractDngDecompressor wants full-size
        # tiles in the buffer only for the rows it reads
        msn = min(ref.dng_decompress(data, offs, lens, img, W, 1, W, th, 1, bps=BPS,
                                     nthreads=ncores, reps=1) for _ in range(reps))
        return {"kind": "reference", "cores": ncores,
                "value": PIX / (msn * 1e-3) / 1e6, "unit": "MPixels/s",
                "single_thread_value": PIX / (ms1 * 1e-3) / 1e6,
                "sample": "1 frame 8256x5504 14-bit MSB, best of %d; value = "
                          "AbstractDngDecompressor(compression 1) over %d row strips with %d "
                          "OpenMP threads; single_thread_value = UncompressedDecompressor as "
                          "shipped (no OpenMP)" % (reps, nt, ncores)}
    t0 = time.perf_counter()
    port.unpack(data, img, W, 1, (0, 0, W, H), pitch, BPS, port.MSB)
    ms = (time.perf_counter() - t0) * 1e3
    return {"kind": "port", "cores": 1, "value": PIX / (ms * 1e-3) / 1e6,
            "unit": "MPixels/s", "sample": "1 frame 8256x5504 14-bit MSB, oracle C port"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = None
    ms_tot = 0.0
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_reference_unpack(reps=1)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    line = {
        "impl": "reference", "metric": "MPixels/s decoded (bit-exact)", "value": v,
        "unit": "MPixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": PIX / v / 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": "configs[1]: 14-bit packed unpack 8256x5504 (45 MP), 1 frame per "
                               "step (bounded sample of the GPU arm's batch)"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "MPixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--frames", type=int, default=8, help="frames per step per GPU")
    ap.add_argument("--ljpeg-frames", type=int, default=8, help="frames in the small LJPEG batch leg")
    ap.add_argument("--ljpeg-big-frames", type=int, default=64,
                    help="frames in the large LJPEG batch leg (one-thread-per-segment path)")
    ap.add_argument("--sustain-s", type=float, default=1.0,
                    help="seconds of the same step back to back after the timed steps "
                         "(clock sampling + sustained figure)")
    ap.add_argument("--c5", action="store_true",
                    help="run BASELINE configs[4]: the 256-frame LJPEG batch sharded over the "
                         "ranks (256/N frames per GPU, strong scaling) + NCCL gather of the outputs")
    ap.add_argument("--skip-others", action="store_true")
    ap.add_argument("--only-unvalidated", action="store_true",
                    help="with --unvalidated: skip the other (validated) secondary legs (short runs under ncu)")
    ap.add_argument("--unvalidated", action="store_true",
                    help="also time the kernels that have not passed their first GPU parity run yet "
                         "(K9 scaling, K10 DNG opcodes, K11 bad pixels, K12 table lookup, Panasonic V4); each leg "
                         "checks bit-exactness against the oracle before timing")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    torch.cuda.set_device(local)
    import rawspeed_b200 as rs
    from oracle import port, synth  # checker + synthetic inputs only
    ctx = rs.Context(local)

    F = args.frames
    peak, peak_src = measured_peaks()

    # ---------------- headline: 14-bit packed unpack ----------------
    data, pitch = synth.packed_frame(W, H, BPS, seed=2 + rank)
    out_pitch = rs.image_pitch(W)
    in_fb = align(pitch * H)
    out_fb = align(out_pitch * H)
    h_in = torch.empty(F * in_fb, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(F * out_fb, dtype=torch.uint8).pin_memory()
    hv = h_in.numpy()
    for f in range(F):
        # distinct frames: frame f = frame 0 with its bytes rotated by f
        hv[f * in_fb:f * in_fb + pitch * H] = np.roll(data, f * 7919)
    d_in = h_in.cuda()
    d_out = torch.zeros(F * out_fb, dtype=torch.uint8, device="cuda")
    plan = rs.unpack_plan(ctx, unpack_jobs(rs, F, in_fb, out_fb, pitch, out_pitch, rs.MSB))
    in_b, out_b, pixels = plan.bytes()

    # parity gate (not timed): frame 0 and the last frame against the oracle
    plan.run(d_in, d_out)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    bit_exact = True
    for f in (0, F - 1):
        want = port.new_image(W, H)
        port.unpack(hv[f * in_fb:f * in_fb + pitch * H], want, W, 1, (0, 0, W, H), pitch, BPS, port.MSB)
        g = got[f * out_fb:f * out_fb + out_pitch * H].view(np.uint16).reshape(H, out_pitch // 2)
        bit_exact &= bool(np.array_equal(g[:, :W], want[:, :W]))
    if not bit_exact:
        print(json.dumps({"error": "GPU output differs from the oracle; no number reported"}))
        sys.exit(1)

    # Timed region first: W warm-up + K timed steps straight away (the kernel timed
    # alone -> compared with the burst peak).  nvidia-smi (100 ms period) cannot
    # resolve a few-ms region, so the sampler runs from before the warm-up until
    # the end of a follow-on sustained loop of the SAME step (--sustain-s seconds,
    # timed separately and reported as `sustained`); only samples taken under load
    # count for the median.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ms = time_steps(torch, lambda: plan.run(d_in, d_out), args.steps, args.warmup, dist)
    launches = ctx.launches - l0 - args.warmup * plan.launches
    sus_n, sus_ms = 0, 0.0
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.sustain_s:
        sus_ms += time_steps(torch, lambda: plan.run(d_in, d_out), 100, 0, None)
        sus_n += 100
    clocks = sampler.stop() if rank == 0 else None

    ms_per_step = ms / args.steps
    value = world * pixels * args.steps / (ms * 1e-3) / 1e6
    ach = (in_b + out_b) / (ms_per_step * 1e-3) / 1e9  # per GPU, one launch per step
    roofline = {"bound": "hbm", "kernel": "unpack_fast_kernel<14,MSB>", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": in_b + out_b,
                "read_only_frac": (in_b / (ms_per_step * 1e-3) / 1e9) / peak,
                "launches_per_step": plan.launches}
    tr = ncu_traffic("unpack_fast_kernel<14,MSB>")
    if tr:
        roofline["traffic"] = tr["dram_bytes_per_frame"] * F
        roofline["traffic_source"] = tr["source"]
    sustained = None
    if sus_n:
        sp = sus_ms / sus_n
        sustained = {"ms_per_step": sp, "steps": sus_n,
                     "value": world * pixels / (sp * 1e-3) / 1e6,
                     "achieved": (in_b + out_b) / (sp * 1e-3) / 1e9,
                     "frac": (in_b + out_b) / (sp * 1e-3) / 1e9 / peak,
                     "note": "same step back to back for %.1f s after the timed steps (rank-local "
                             "timing); the board reaches its power cap here, so this is the "
                             "sustained figure against the same burst peak" % args.sustain_s}

    # ---------------- e2e: host buffers through the C-ABI call ----------------
    def e2e_step():
        plan.run_host(h_in.numpy(), h_out.numpy())
    e2e_steps = max(3, min(args.steps, 5))
    ms_e = wall_steps(torch, e2e_step, e2e_steps, 1, dist)
    e2e = {"value": world * pixels * e2e_steps / (ms_e * 1e-3) / 1e6, "unit": "MPixels/s",
           "h2d_bytes_per_step": int(F * in_fb), "d2h_bytes_per_step": int(F * out_fb),
           "steps": e2e_steps, "ms_per_step": ms_e / e2e_steps,
           "api": "rsb200_plan_run_host (pinned host buffers, H2D + kernel + D2H per step)"}
    ok = np.array_equal(h_out.numpy()[:out_pitch * H], got[:out_pitch * H])
    e2e["bit_exact"] = bool(ok)

    others = {}
    if not args.skip_others:
        others = bench_others(torch, rs, ctx, port, synth, args, dist, peak)

    gather = None
    if dist is not None:
        gather = bench_gather(torch, dist, d_out, world, rank, plan, d_in, args, F, out_fb)

    if rank == 0:
        cpu = None if args.skip_cpu else cpu_reference_unpack()
        line = {
            "metric": "MPixels/s decoded (bit-exact)", "value": value, "unit": "MPixels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": "configs[1]: 14-bit packed (MSB) unpack, 8256x5504 (45 MP), "
                                   "%d frames per step per GPU" % F,
                       "frames_per_step_per_gpu": F, "bytes_per_step_per_gpu": in_b + out_b,
                       "l2": "inputs+outputs of one step (%.2f GB) exceed the 126 MB L2; no flush "
                             "needed" % ((in_b + out_b) / 1e9),
                       "parallelism": "frames sharded across ranks, no data-path collective"},
            "bit_exact": bit_exact, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "sustained": sustained,
            "others": others,
        }
        if gather:
            line["gather"] = gather
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_gather(torch, dist, d_out, world, rank, plan, d_in, args, frames, out_fb):
    """north_star's NVLink output gather (every rank ends up with all decoded
    frames), timed separately from the decode: rawspeed_b200.shard.gather_frames
    = one NCCL all_gather on the decode stream."""
    from rawspeed_b200 import shard
    local = d_out.view(frames, out_fb)
    gathered = torch.empty((world, frames, out_fb), dtype=torch.uint8, device="cuda")

    def step():
        plan.run(d_in, d_out)
        # copy-free form: preallocated result, the collective's own layout (frame r + k*world at [r, k])
        shard.gather_frames(local, frames * world, dist, out=gathered, reorder=False)
    n = max(2, min(args.steps, 5))
    ms = time_steps(torch, step, n, 1, dist)
    total = frames * world * out_fb
    return {"what": "decode + ncclAllGather of the uint16 outputs over NVLink (all ranks get all frames)",
            "ms_per_step": ms / n, "gathered_bytes_per_rank": int(total),
            "MPixels/s": world * frames * PIX / (ms / n * 1e-3) / 1e6,
            "busbw_GBps": (total * (world - 1) / world) / (ms / n * 1e-3) / 1e9}


def bench_others(torch, rs, ctx, port, synth, args, dist, peak):
    """configs[2] (DNG LJPEG tiles) and configs[3] (CR2): device-timed decode."""
    from helpers import dng_ljpeg_scans, parse_ljpeg, TableSet
    if args.unvalidated and args.only_unvalidated:
        return bench_unvalidated(torch, rs, ctx, port, synth, args, dist, peak)
    out = {}
    steps = max(3, min(args.steps, 10))
    # ---- C3: 8256x5504 DNG, 726 LJPEG tiles of 256x256, 2 components ----
    img = synth.image_model(W, H, 12345)
    t = synth.make_dng_ljpeg(img, 256, 256)
    out_pitch = rs.image_pitch(W)
    tabs, scans = dng_ljpeg_scans(t, out_pitch)
    plan = rs.ljpeg_plan(ctx, tabs.tabs, scans)
    d_in = torch.zeros(t.blob.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:t.blob.size] = torch.from_numpy(t.blob)
    d_out = torch.zeros(H * out_pitch, dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), t.blob.size), d_out)
    res = plan.results()
    got = d_out.cpu().numpy().view(np.uint16).reshape(H, out_pitch // 2)
    exact = bool(np.array_equal(got[:, :W], img)) and all(s == 0 for s, _ in res)
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), t.blob.size), d_out), steps, 3, dist)
    in_b, out_b, pixels = plan.bytes()
    per = ms / steps
    c3 = {
        "MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
        "compressed_bytes_per_pixel": t.blob.size / PIX,
        "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
        "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak,
        "read_only_roofline_frac": in_b / (per * 1e-3) / 1e9 / peak,
        "kernel": "k2_fused_kernel", "launches_per_frame": plan.launches}
    out["configs[2] DNG LJPEG 8256x5504 (726 tiles 256x256)"] = c3
    del plan, d_out
    # ---- C5-style batches: NB frames of C3 resident in HBM, one plan per batch ----
    # 8 frames stay on the block-per-segment kernel (K2F); from ~22 frames (16384 segments)
    # the plan switches to the one-thread-per-segment path (K2C unstuff pre-pass + K2T).
    world = int(os.environ.get("WORLD_SIZE", "1"))
    fb = (t.blob.size + 255) // 256 * 256
    ob = (H * out_pitch + 255) // 256 * 256
    batches = [max(1, 256 // world)] if args.c5 else sorted({args.ljpeg_frames, args.ljpeg_big_frames})
    for NB in batches:
        d_inb = torch.zeros(NB * fb + 64, dtype=torch.uint8, device="cuda")
        scans_b = []
        for f in range(NB):
            d_inb[f * fb:f * fb + t.blob.size] = d_in[:t.blob.size]
            for s0 in scans:
                s1 = rs.LJpegScan.from_buffer_copy(s0)
                s1.in_offset = s0.in_offset + f * fb
                s1.out_offset = s0.out_offset + f * ob
                scans_b.append(s1)
        planb = rs.ljpeg_plan(ctx, tabs.tabs, scans_b)
        d_outb = torch.zeros(NB * ob, dtype=torch.uint8, device="cuda")
        planb.run((d_inb.data_ptr(), NB * fb), d_outb)
        resb = planb.results()
        exact_b = all(s_ == 0 for s_, _ in resb)
        for f in sorted({0, NB // 2, NB - 1}):
            gb = d_outb[f * ob:f * ob + H * out_pitch].cpu().numpy().view(np.uint16).reshape(H, out_pitch // 2)
            exact_b = exact_b and bool(np.array_equal(gb[:, :W], img))
        nst = max(3, min(steps, 5))
        msb = time_steps(torch, lambda: planb.run((d_inb.data_ptr(), NB * fb), d_outb), nst, 3, dist)
        in_bb, out_bb, pix_b = planb.bytes()
        perb = msb / nst
        kern = "k2_fused_kernel" if planb.launches == 1 else "k2_clean_kernel + k2_thread_kernel"
        label = ("configs[4]: 256-frame LJPEG batch, %d frames per GPU x %d GPUs, one plan per GPU"
                 % (NB, world)) if args.c5 else \
            "configs[4]-style batch: %d LJPEG frames of configs[2] per GPU, one plan" % NB
        entb = {
            "MPixels/s_per_gpu": pix_b / (perb * 1e-3) / 1e6, "ms_per_step": perb, "bit_exact": exact_b,
            "kernels": kern, "launches_per_step": planb.launches,
            "achieved_GBps": (in_bb + out_bb) / (perb * 1e-3) / 1e9,
            "roofline_frac": (in_bb + out_bb) / (perb * 1e-3) / 1e9 / peak,
            "read_only_roofline_frac": in_bb / (perb * 1e-3) / 1e9 / peak}
        if args.c5:
            entb["MPixels/s_all_gpus"] = world * pix_b / (perb * 1e-3) / 1e6
            entb["frames"] = NB * world
            entb["note"] = ("the 256 frames are copies of one synthetic frame (same statistics; generating "
                            "256 distinct frames on the host would take minutes); ms_per_step is the max "
                            "over ranks")
            if dist is not None:
                from rawspeed_b200 import shard
                local = d_outb.view(NB, ob)
                gathered = torch.empty((world, NB, ob), dtype=torch.uint8, device="cuda")

                def step_g():
                    planb.run((d_inb.data_ptr(), NB * fb), d_outb)
                    shard.gather_frames(local, NB * world, dist, out=gathered, reorder=False)
                msg = time_steps(torch, step_g, 3, 1, dist)
                entb["decode_plus_gather_ms"] = msg / 3
                entb["decode_plus_gather_MPixels/s"] = world * pix_b / (msg / 3 * 1e-3) / 1e6
                entb["gather"] = ("one ncclAllGather of the uint16 outputs into a preallocated "
                                  "[world, frames_per_gpu, frame] buffer on every rank (frame r + k*world "
                                  "at [r, k]); %.1f GB received per GPU" % ((world - 1) * NB * ob / 1e9))
                del gathered
        out[label] = entb
        del planb, d_inb, d_outb
    del d_in
    if not args.skip_cpu and int(os.environ.get("RANK", "0")) == 0:
        import oracle
        if oracle.HAVE_REF:
            ncores = os.cpu_count() or 1
            tmp = port.new_image(W, H)
            ms_cpu = min(oracle.ref.dng_decompress(t.blob, t.offsets, t.lengths, tmp, W, 1, 256, 256, 7,
                                                   nthreads=ncores, reps=1) for _ in range(3))
            ms_1 = oracle.ref.dng_decompress(t.blob, t.offsets, t.lengths, tmp, W, 1, 256, 256, 7,
                                             nthreads=1, reps=1)
            c3["cpu_reference"] = {"kind": "reference", "cores": ncores,
                                   "MPixels/s": PIX / (ms_cpu * 1e-3) / 1e6,
                                   "single_thread_MPixels/s": PIX / (ms_1 * 1e-3) / 1e6,
                                   "sample": "AbstractDngDecompressor::decompress() (OpenMP over the 726 "
                                             "tiles), 1 frame, best of 3"}
    # ---- C4: CR2 6720x4480, 3 slices, 2 and 4 components ----
    from test_gpu_cr2 import cr2_job
    cw, ch = 6720, 4480
    cimg = port.new_image(cw, ch)
    cimg[:, :cw] = synth.image_model(cw, ch, 4)
    hts = synth.default_tables(2)
    for fmt, frame in [((2, 1, 1), (3360, 4480)), ((4, 1, 1), (1680, 4480))]:
        blob = port.cr2_encode(cimg, cw, fmt, frame, (3, 2240, 2240), 14, hts, [0, 1, 0, 1][:fmt[0]])
        ts = TableSet()
        job = cr2_job(blob, cw, ch, fmt, (3, 2240, 2240), cimg.shape[1] * 2, ts)
        plan = rs.cr2_plan(ctx, ts.tabs, [job])
        d_in = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:blob.size] = torch.from_numpy(blob)
        d_out = torch.zeros(cimg.size * 2, dtype=torch.uint8, device="cuda")
        plan.run((d_in.data_ptr(), blob.size), d_out)
        res = plan.results()
        got = d_out.cpu().numpy().view(np.uint16).reshape(cimg.shape)
        exact = bool(np.array_equal(got[:, :cw], cimg[:, :cw])) and res[0][0] == 0
        ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), blob.size), d_out), 3, 1, dist)
        per = ms / 3
        ent = {"MPixels/s": cw * ch / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
               "compressed_bytes_per_pixel": blob.size / (cw * ch),
               "kernels": "k2_range_count/verify/diffs + k3_column/row"}
        if not args.skip_cpu and int(os.environ.get("RANK", "0")) == 0:
            import oracle
            if oracle.HAVE_REF:
                tmp = port.new_image(cw, ch)
                msr = min(oracle.ref.cr2_ljpeg_decode(blob, tmp, cw, (3, 2240, 2240), reps=1)
                          for _ in range(2))
                ent["cpu_reference"] = {"kind": "reference", "cores": 1,
                                        "MPixels/s": cw * ch / (msr * 1e-3) / 1e6,
                                        "sample": "Cr2LJpegDecoder::decode (single threaded by design)"}
        out["configs[3] CR2 6720x4480 3 slices <%d,1,1>" % fmt[0]] = ent
        del plan, d_in, d_out
    out.update(bench_forms(torch, rs, ctx, port, synth, args, dist, peak))
    out.update(bench_codecs(torch, rs, ctx, port, synth, args, dist, peak))
    if args.unvalidated:
        out.update(bench_unvalidated(torch, rs, ctx, port, synth, args, dist, peak))
    return out


def bench_codecs(torch, rs, ctx, port, synth, args, dist, peak):
    """Canon sRaw interpolation, the Pentax PEF codec, Sony ARW2; device-timed."""
    out = {}
    steps = max(3, min(args.steps, 10))
    rank0 = int(os.environ.get("RANK", "0")) == 0
    # ---- Cr2sRawInterpolator, 4:2:0 version 2, 5040x3360 RGB output (mRAW class) ----
    num_mcus, rows = 2520, 1680
    rng = np.random.default_rng(5)
    in_w = num_mcus * 6
    pitch = (in_w * 2 + 15) // 16 * 16
    inp = np.zeros((rows, pitch // 2), dtype=np.uint16)
    inp[:, :in_w] = rng.integers(0, 16384, (rows, in_w), dtype=np.uint16)
    out_w, out_h = 2 * num_mcus, 2 * rows
    want = port.new_image(out_w, out_h, 3)
    j = rs.SrawJob()
    j.in_offset, j.in_pitch, j.num_mcus, j.in_rows = 0, pitch, num_mcus, rows
    j.sub_x, j.sub_y, j.version = 2, 2, 2
    j.sraw_coeffs[0], j.sraw_coeffs[1], j.sraw_coeffs[2] = 2000, 1024, 1500
    j.hue, j.out_offset, j.out_pitch = 0, 0, want.shape[1] * 2
    plan = rs.sraw_plan(ctx, [j])
    d_in = torch.from_numpy(inp.view(np.int16)).cuda()
    d_out = torch.from_numpy(want.view(np.int16).copy()).cuda()
    plan.run(d_in, d_out)
    torch.cuda.synchronize()
    port.sraw_interpolate(inp, in_w, want, out_w, (2, 2), (2000, 1024, 1500), 0, 2)
    exact = bool(np.array_equal(d_out.cpu().numpy().view(np.uint16), want))
    ms = time_steps(torch, lambda: plan.run(d_in, d_out), steps, 3, dist)
    in_b, out_b, pixels = plan.bytes()
    per = ms / steps
    ent = {"MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
           "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
           "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak, "kernel": "sraw_kernel<2,420>"}
    if not args.skip_cpu and rank0:
        import oracle
        if oracle.HAVE_REF:
            ncores = os.cpu_count() or 1
            tmp = want.copy()
            msr = min(oracle.ref.sraw_interpolate(inp, in_w, tmp, out_w, (2, 2), (2000, 1024, 1500), 0, 2,
                                                  nthreads=ncores) for _ in range(3))
            ent["cpu_reference"] = {"kind": "reference", "cores": ncores,
                                    "MPixels/s": pixels / (msr * 1e-3) / 1e6,
                                    "sample": "Cr2sRawInterpolator::interpolate(2), OpenMP rows, best of 3"}
    out["8(f)2 Cr2sRawInterpolator 4:2:0 -> 5040x3360 RGB"] = ent
    del plan, d_in, d_out
    # ---- PentaxDecompressor, 6016x4000 (K-3 class), legacy table ----
    w, h = 6016, 4000
    table = port.pentax_table(None)
    img = (synth.image_model(w, h, seed=11, bits=12) & 0x0FFF).astype(np.uint16)
    data = synth.make_pentax(img, table)
    got0 = port.new_image(w, h)
    pj = rs.PentaxJob()
    pj.in_offset, pj.in_size, pj.table, pj.width, pj.height = 0, data.size, 0, w, h
    pj.out_offset, pj.out_pitch = 0, got0.shape[1] * 2
    plan = rs.pentax_plan(ctx, [rs.huff_table(table[0], table[1])], [pj])
    d_in = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:data.size] = torch.from_numpy(data)
    d_out = torch.from_numpy(got0.view(np.int16).copy()).cuda()
    plan.run((d_in.data_ptr(), data.size), d_out)
    res = plan.results()
    exact = bool(np.array_equal(d_out.cpu().numpy().view(np.uint16)[:, :w], img)) and res[0][0] == 0
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), data.size), d_out), 3, 1, dist)
    per = ms / 3
    ent = {"MPixels/s": w * h / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
           "compressed_bytes_per_pixel": data.size / (w * h),
           "kernels": "k2_range_count/verify/diffs (plain MSB pump) + k3p_column/row"}
    if not args.skip_cpu and rank0:
        import oracle
        if oracle.HAVE_REF:
            tmp = port.new_image(w, h)
            msr = min(oracle.ref.pentax_decompress(tmp, w, data) for _ in range(2))
            ent["cpu_reference"] = {"kind": "reference", "cores": 1,
                                    "MPixels/s": w * h / (msr * 1e-3) / 1e6,
                                    "sample": "PentaxDecompressor::decompress (single threaded by design)"}
    out["8(f)2 PentaxDecompressor 6016x4000"] = ent
    del plan, d_in, d_out
    # ---- NikonDecompressor (no split), 6032x4032 14-bit, curve + dither ----
    w, h = 6032, 4032
    half = 1 << 13
    pup = [half, half + 2, half - 8, half - 2]
    meta = synth.nikon_meta("table", 14, (pup[0], pup[2], pup[1], pup[3]), True)
    su = port.nikon_setup(meta, True, 14, w, h)
    img = (synth.image_model(w, h, seed=7, bits=14) & 0x3FFF).astype(np.uint16)
    data = synth.make_nikon(img, su["huff_select"], pup)
    ncpl, values = port.nikon_tree(su["huff_select"])
    nj = rs.NikonJob()
    nj.in_offset, nj.in_size, nj.table, nj.width, nj.height = 0, data.size, 0, w, h
    nj.out_offset, nj.out_pitch, nj.lut = 0, rs.image_pitch(w), 0
    for k in range(4):
        nj.pup[k] = pup[k]
    plan = rs.nikon_plan(ctx, [rs.huff_table(ncpl, values)], [nj], port.build_table(su["curve"], True))
    d_in = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:data.size] = torch.from_numpy(data)
    d_out = torch.zeros(h * rs.image_pitch(w), dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), data.size), d_out)
    res = plan.results()
    want = port.new_image(w, h)
    port.nikon_decompress(want, w, meta, True, 14, data)
    exact = bool(np.array_equal(d_out.cpu().numpy().view(np.uint16).reshape(want.shape), want)) and res[0][0] == 0
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), data.size), d_out), 3, 1, dist)
    per = ms / 3
    ent = {"MPixels/s": w * h / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
           "compressed_bytes_per_pixel": data.size / (w * h),
           "kernels": "k2_range_count/verify/diffs (plain MSB pump) + k3n_column/row (curve + dither)"}
    if not args.skip_cpu and rank0:
        import oracle
        if oracle.HAVE_REF:
            tmp = port.new_image(w, h)
            msr = min(oracle.ref.nikon_decompress(tmp, w, meta, True, 14, data) for _ in range(2))
            ent["cpu_reference"] = {"kind": "reference", "cores": 1,
                                    "MPixels/s": w * h / (msr * 1e-3) / 1e6,
                                    "sample": "NikonDecompressor::decompress (single threaded by design)"}
    out["8(f)2 NikonDecompressor 6032x4032 14-bit (curve + dither)"] = ent
    del plan, d_in, d_out
    # ---- PanasonicV5 (14 bit) / V6 (14 bit) / V7, 5184x3888-class frames, 4 frames per launch ----
    for ver, bps, w, h in ((5, 14, 5184, 3888), (6, 14, 5181, 3888), (7, 14, 5184, 3888)):
        npix = (11 if ver == 6 else 128 // bps)
        nunits = w * h // npix
        nbytes = ((nunits + 1023) // 1024) * 0x4000 if ver == 5 else nunits * 16
        data = synth.lcg_bytes(nbytes, 40 + ver)
        opitch = rs.image_pitch(w)
        nf = 4
        fb = (nbytes + 255) // 256 * 256
        ob = (h * opitch + 255) // 256 * 256
        jobs = []
        for f in range(nf):
            pj = rs.PanaJob()
            pj.in_offset, pj.in_size, pj.out_offset, pj.out_pitch = f * fb, nbytes, f * ob, opitch
            pj.width, pj.height, pj.version, pj.bps = w, h, ver, bps
            jobs.append(pj)
        plan = rs.pana_plan(ctx, jobs)
        d_in = torch.zeros(nf * fb + 64, dtype=torch.uint8, device="cuda")
        for f in range(nf):
            d_in[f * fb:f * fb + nbytes] = torch.from_numpy(data)
        d_out = torch.zeros(nf * ob, dtype=torch.uint8, device="cuda")
        plan.run((d_in.data_ptr(), nf * fb), d_out)
        torch.cuda.synchronize()
        want = port.new_image(w, h)
        port.panasonic(ver, want, w, data, bps)
        got = d_out[(nf - 1) * ob:(nf - 1) * ob + h * opitch].cpu().numpy().view(np.uint16).reshape(h, opitch // 2)
        exact = bool(np.array_equal(got[:, :w], want[:, :w]))
        ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), nf * fb), d_out), steps, 3, dist)
        in_b, out_b, pixels = plan.bytes()
        per = ms / steps
        ent = {"MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_step": per, "frames_per_step": nf,
               "bit_exact": exact, "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
               "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak,
               "kernel": "pana_kernel<%d,%d>" % (ver, bps)}
        if not args.skip_cpu and rank0:
            import oracle
            if oracle.HAVE_REF:
                ncores = os.cpu_count() or 1
                tmp = port.new_image(w, h)
                msr = min(oracle.ref.panasonic(ver, tmp, w, data, bps, nthreads=ncores) for _ in range(3))
                ent["cpu_reference"] = {"kind": "reference", "cores": ncores,
                                        "MPixels/s": w * h / (msr * 1e-3) / 1e6,
                                        "sample": "PanasonicV%dDecompressor::decompress (OpenMP), 1 frame, best of 3" % ver}
        out["8(f)4 PanasonicV%dDecompressor %dx%d %d-bit" % (ver, w, h, bps)] = ent
        del plan, d_in, d_out
    # ---- PhaseOneDecompressor, 11608x8708 (IQ3 100MP class): one thread per row ----
    w, h = 11608, 8708
    rowimg = (synth.image_model(w, 4, seed=31, bits=14)).astype(np.uint16)
    rows4 = [np.frombuffer(synth.phaseone_row(rowimg[k]), dtype=np.uint8) for k in range(4)]
    offs, blobs, pos = [], [], 0
    for r in range(h):   # the four encoded rows repeat down the image (rows are independent streams)
        offs.append((pos, rows4[r % 4].size, r))
        blobs.append(rows4[r % 4])
        pos += rows4[r % 4].size
    blob = np.concatenate(blobs)
    pj = rs.PhaseOneJob()
    pj.out_offset, pj.out_pitch, pj.width, pj.height, pj.first_strip = 0, rs.image_pitch(w), w, h, 0
    pstrips = []
    for off, size, row in offs:
        ps = rs.PhaseOneStrip()
        ps.in_offset, ps.in_size, ps.row = off, size, row
        pstrips.append(ps)
    plan = rs.phaseone_plan(ctx, [pj], pstrips)
    d_in = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:blob.size] = torch.from_numpy(blob)
    d_out = torch.zeros(h * rs.image_pitch(w), dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), blob.size), d_out)
    res = plan.results()
    got = d_out.cpu().numpy().view(np.uint16).reshape(h, rs.image_pitch(w) // 2)
    exact = res[0][0] == 0 and all(bool(np.array_equal(got[r, :w], rowimg[r % 4])) for r in (0, 1, 2, 3, h - 1))
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), blob.size), d_out), 3, 1, dist)
    per = ms / 3
    ent = {"MPixels/s": w * h / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": bool(exact),
           "compressed_bytes_per_pixel": blob.size / (w * h),
           "kernel": "p1_kernel (one thread per row: 8708 threads, latency bound)"}
    if not args.skip_cpu and rank0:
        import oracle
        if oracle.HAVE_REF:
            ncores = os.cpu_count() or 1
            tmp = port.new_image(w, h)
            msr = min(oracle.ref.phaseone(tmp, w, blob, offs, nthreads=ncores) for _ in range(2))
            ent["cpu_reference"] = {"kind": "reference", "cores": ncores,
                                    "MPixels/s": w * h / (msr * 1e-3) / 1e6,
                                    "sample": "PhaseOneDecompressor::decompress (OpenMP over rows), best of 2"}
    out["8(f)4 PhaseOneDecompressor 11608x8708"] = ent
    del plan, d_in, d_out
    # ---- SonyArw2Decompressor, 9568x6376 (61 MP, A7R IV class), dithered curve, 4 frames ----
    w, h, nf = 9568, 6376, 4
    data = synth.arw2_frame(w, h, seed=21)
    curve = synth.sony_curve()
    table = port.build_table(curve, True)
    opitch = rs.image_pitch(w)
    fb = (data.size + 255) // 256 * 256
    ob = (h * opitch + 255) // 256 * 256
    jobs = []
    for f in range(nf):
        aj = rs.Arw2Job()
        aj.in_offset, aj.out_offset, aj.out_pitch = f * fb, f * ob, opitch
        aj.width, aj.height, aj.table = w, h, 0
        jobs.append(aj)
    plan = rs.arw2_plan(ctx, jobs, table, True)
    d_in = torch.zeros(nf * fb + 64, dtype=torch.uint8, device="cuda")
    for f in range(nf):
        d_in[f * fb:f * fb + data.size] = torch.from_numpy(data)
    d_out = torch.zeros(nf * ob, dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), nf * fb), d_out)
    res = plan.results()
    want = port.new_image(w, h)
    port.sony_arw2(want, w, data, table, True)
    got = d_out[(nf - 1) * ob:(nf - 1) * ob + h * opitch].cpu().numpy().view(np.uint16).reshape(h, opitch // 2)
    exact = bool(np.array_equal(got, want)) and all(s_ == 0 for s_, _ in res)
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), nf * fb), d_out), steps, 3, dist)
    in_b, out_b, pixels = plan.bytes()
    per = ms / steps
    ent = {"MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_step": per, "frames_per_step": nf,
           "bit_exact": exact, "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
           "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak,
           "algorithmic_bytes_per_pixel": 3.0, "kernel": "arw2_kernel<dither>"}
    if not args.skip_cpu and rank0:
        import oracle
        if oracle.HAVE_REF:
            ncores = os.cpu_count() or 1
            tmp = port.new_image(w, h)
            msr = min(oracle.ref.sony_arw2(tmp, w, data, curve, True, nthreads=ncores) for _ in range(3))
            ms1 = oracle.ref.sony_arw2(tmp, w, data, curve, True, nthreads=1)
            ent["cpu_reference"] = {"kind": "reference", "cores": ncores,
                                    "MPixels/s": w * h / (msr * 1e-3) / 1e6,
                                    "single_thread_MPixels/s": w * h / (ms1 * 1e-3) / 1e6,
                                    "sample": "SonyArw2Decompressor::decompress (OpenMP over rows), 1 frame, best of 3"}
    out["8(f)4 SonyArw2Decompressor 9568x6376 (dithered curve)"] = ent
    del plan, d_in, d_out
    return out


def bench_unvalidated(torch, rs, ctx, port, synth, args, dist, peak):
    """SURVEY 8(f)3 (+ Panasonic V4): kernels written after round 1's GPU budget was spent.  Off by
    default (--unvalidated); every leg first checks the result against the oracle."""
    out = {}
    steps = max(3, min(args.steps, 10))
    W, H = 8256, 5504
    pitch = rs.image_pitch(W)
    rng = np.random.default_rng(9)
    base = port.new_image(W, H)
    base[:, :] = rng.integers(0, 16384, size=base.shape, dtype=np.uint16)

    rank0 = int(os.environ.get("RANK", "0")) == 0
    ncores = os.cpu_count() or 1

    def leg(name, plan, want, kernel, restore=True, cpu=None, cpu_threads=None):
        d = torch.from_numpy(base.view(np.int16).copy()).cuda()
        src = d.clone()
        plan.run(None, d)
        torch.cuda.synchronize()
        exact = bool(np.array_equal(d.cpu().numpy().view(np.uint16), want))

        def step():
            if restore:
                d.copy_(src)        # in-place kernels: every timed run starts from the same pixels
            plan.run(None, d)
        ms = time_steps(torch, step, steps, 3, dist)
        ms_copy = time_steps(torch, lambda: d.copy_(src), steps, 3, dist) if restore else 0.0
        in_b, out_b, pixels = plan.bytes()
        per = (ms - ms_copy) / steps
        out[name] = {"MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
                     "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
                     "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak, "kernel": kernel,
                     "timing": "in-place kernel + restoring copy, minus the copy alone"}
        if cpu is not None and not args.skip_cpu and rank0:
            import oracle
            if oracle.HAVE_REF:
                best = 1e30
                for _ in range(3):
                    cpu(base.copy())
                    best = min(best, oracle.ref.last_ms())
                out[name]["cpu_reference"] = {"kind": "reference", "cores": cpu_threads or ncores,
                                              "MPixels/s": W * H / (best * 1e-3) / 1e6, "ms": best,
                                              "sample": "the reference's own member on 1 frame, best of 3 (driver copies excluded)"}

    # K9: black / white scaling, both loops
    for label, black, white in (("SSE2 loop", (1008, 1010, 1009, 1011), 16383), ("plain loop", (64,) * 4, 1000)):
        j = rs.ScaleJob()
        j.offset, j.pitch, j.width, j.height, j.cpp = 0, pitch, W, H, 1
        j.crop_x, j.crop_y, j.crop_w, j.crop_h = 8, 8, W - 16, H - 16
        for i in range(4):
            j.black_separate[i] = black[i]
        j.white_point, j.dither, j.path = white, 1, 0
        want = base.copy()
        port.scale_values(want, W, (8, 8, W - 16, H - 16), black, white)
        leg("8(f)3 scaleBlackWhite 8256x5504 (%s, dither)" % label, rs.scale_plan(ctx, [j]), want,
            "scale_kernel<%d>" % (0 if "SSE2" in label else 1),
            cpu=lambda im, black=black, white=white: __import__("oracle").ref.scale_values(
                im, W, (8, 8, W - 16, H - 16), black, white, nthreads=ncores))
    # K12: whole-image table lookup, Sony curve, plain and dithered
    for dither in (False, True):
        lj = rs.LookupJob()
        lj.offset, lj.pitch, lj.width, lj.height, lj.cpp, lj.table = 0, pitch, W, H, 1, 0
        t = port.build_table(synth.sony_curve(), dither)
        want = base.copy()
        port.sixteen_bit_lookup(want, W, 1, t, dither)
        leg("8(f)3 sixteenBitLookup 8256x5504 (%s)" % ("dithered" if dither else "plain"),
            rs.lookup_plan(ctx, [lj], t, dither), want, "lookup_kernel<%s>" % ("true" if dither else "false"),
            cpu=lambda im, dither=dither: __import__("oracle").ref.sixteen_bit_lookup(
                im, W, 1, [0, 0, W, H], synth.sony_curve(), dither, nthreads=ncores))
    # K10: eight opcodes in one pass
    from rawspeed_b200 import host
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
    leg("8(f)3 DngOpcodes 8256x5504, 8 opcodes in one pass", rs.dngop_plan(ctx, [dj], low["ops"], low["tables"], low["deltas"]),
        want, "dngop_kernel", cpu=lambda im: __import__("oracle").ref.dng_opcodes(im, W, 1, [0, 0, W, H], blob),
        cpu_threads=1)   # applyOpCodes is single threaded in the reference
    # K11: 20 000 defects
    n = 20000
    p = ((rng.integers(0, H, n).astype(np.uint32) << 16) | rng.integers(0, W, n).astype(np.uint32))
    bj = rs.BadPixJob()
    bj.offset, bj.pitch, bj.width, bj.height, bj.is_cfa = 0, pitch, W, H, 1
    bj.first_position, bj.num_positions, bj.prior_map = 0, n, None
    want = base.copy()
    port.fix_bad_pixels(want, W, 1, p, True)
    leg("8(f)3 fixBadPixels 8256x5504, 20000 defects", rs.badpix_plan(ctx, [bj], p), want, "badpix_kernel",
        restore=False,     # idempotent: good pixels are never written
        cpu=lambda im: __import__("oracle").ref.fix_bad_pixels(im, W, 1, p, True, nthreads=ncores))
    # Panasonic V4, 4592x3448-class frames, 4 per launch
    w, h, split = 4592 // 14 * 14, 3448, 0x2008
    nbytes = (w * h // 14 * 16 + 0x3FFF) // 0x4000 * 0x4000
    data = synth.lcg_bytes(nbytes, 44)
    opitch = rs.image_pitch(w)
    nf, fb, ob = 4, (nbytes + 255) // 256 * 256, (h * opitch + 255) // 256 * 256
    jobs = []
    for f in range(nf):
        pj = rs.PanaJob()
        pj.in_offset, pj.in_size, pj.out_offset, pj.out_pitch = f * fb, nbytes, f * ob, opitch
        pj.width, pj.height, pj.version, pj.bps = w, h, 4, 12
        pj.zero_is_not_bad, pj.section_split_offset = 0, split
        jobs.append(pj)
    plan = rs.pana_plan(ctx, jobs)
    d_in = torch.zeros(nf * fb + 64, dtype=torch.uint8, device="cuda")
    for f in range(nf):
        d_in[f * fb:f * fb + nbytes] = torch.from_numpy(data)
    d_out = torch.zeros(nf * ob, dtype=torch.uint8, device="cuda")
    plan.run((d_in.data_ptr(), nf * fb), d_out)
    torch.cuda.synchronize()
    want = port.new_image(w, h)
    zwant = port.panasonic_v4(want, w, data, False, split, cap=1 << 22)
    got = d_out[(nf - 1) * ob:(nf - 1) * ob + h * opitch].cpu().numpy().view(np.uint16).reshape(h, opitch // 2)
    nz, zl = plan.bad_pixels(nf - 1, cap=1 << 22)
    exact = bool(np.array_equal(got[:, :w], want[:, :w])) and sorted(zl) == zwant
    ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), nf * fb), d_out), steps, 3, dist)
    in_b, out_b, pixels = plan.bytes()
    per = ms / steps
    out["8(f)4 PanasonicV4Decompressor %dx%d" % (w, h)] = {
        "MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_step": per, "frames_per_step": nf, "bit_exact": exact,
        "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
        "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak, "kernel": "pana_kernel<4,12>"}
    return out


def bench_forms(torch, rs, ctx, port, synth, args, dist, peak):
    """the fixed-layout UncompressedDecompressor forms, one 8256x5504
    frame each, device-timed like the headline (inputs resident in HBM)."""
    from rawspeed_b200 import formats as F
    out = {}
    steps = max(3, min(args.steps, 10))
    cases = [("decode12BitRawWithControl<big>", F.RAW_12BIT_CONTROL_BE, 12 * W // 8 + (W + 2) // 10,
              port.FORM_12BIT_CONTROL_BE, 12, port.MSB, False),
             ("decode12BitRawUnpackedLeftAligned<little>", F.RAW_12BIT_LEFT_LE, 2 * W,
              port.FORM_12BIT_LEFT_LE, 16, port.LSB, False),
             ("decode8BitRaw<uncorrected>", F.RAW_8BIT, W, port.FORM_8BIT_UNCORRECTED, 8, port.LSB, False),
             ("decodePackedFP<MSB, binary16> -> float", F.RAW_FP16_MSB, 2 * W, port.FORM_READ, 16,
              port.MSB, True)]
    for name, fmt, pitch, form, bps, order, f32 in cases:
        data = synth.lcg_bytes(pitch * H, seed=7)
        want = port.new_image_f32(W, H) if f32 else port.new_image(W, H)
        got0 = want.copy()
        j = rs.RawJob()
        j.in_offset, j.in_size, j.out_offset = 0, data.size, 0
        j.out_pitch = want.shape[1] * want.itemsize
        j.row0, j.rows, j.samples, j.out_col0 = 0, H, W, 0
        j.in_pitch, j.format, j.table = pitch, fmt, 0
        plan = rs.raw_plan(ctx, [j])
        d_in = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:data.size] = torch.from_numpy(data.copy())
        d_out = torch.from_numpy(got0.view(np.uint8).reshape(-1).copy()).cuda()
        plan.run((d_in.data_ptr(), data.size), d_out)
        torch.cuda.synchronize()
        port.unpack_form(data, want, W, 1, (0, 0, W, H), pitch, bps, order, form)
        exact = bool(np.array_equal(d_out.cpu().numpy().view(want.dtype).reshape(want.shape), want))
        ms = time_steps(torch, lambda: plan.run((d_in.data_ptr(), data.size), d_out), steps, 3, dist)
        in_b, out_b, pixels = plan.bytes()
        per = ms / steps
        ent = {"MPixels/s": pixels / (per * 1e-3) / 1e6, "ms_per_frame": per, "bit_exact": exact,
               "achieved_GBps": (in_b + out_b) / (per * 1e-3) / 1e9,
               "roofline_frac": (in_b + out_b) / (per * 1e-3) / 1e9 / peak,
               "kernel": "rawform_kernel<%d>" % fmt,
               "note": "single 45 MP frame per launch (%.0f MB moved): a short launch, below "
                       "the batch figure" % ((in_b + out_b) / 1e6)}
        if not args.skip_cpu and int(os.environ.get("RANK", "0")) == 0:
            import oracle
            if oracle.HAVE_REF:
                tmp = want.copy()
                msr = min(oracle.ref.unpack_form(data, tmp, W, 1, (0, 0, W, H), pitch, bps, order,
                                                 form, reps=1) for _ in range(2))
                ent["cpu_reference"] = {"kind": "reference", "cores": 1,
                                        "MPixels/s": PIX / (msr * 1e-3) / 1e6,
                                        "sample": "1 frame, best of 2 (single threaded by design)"}
        out["8(f)1 " + name + " 8256x5504"] = ent
        del plan, d_in, d_out
    return out


if __name__ == "__main__":
    main()
