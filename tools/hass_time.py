"""Timing of the Hasselblad decode (K2H) and the DNG opcode pass (K10) for the library in RSB200_LIB
(development tool for A/B runs, not the benchmark; bench.py --all-legs has the numbers of record).

    RSB200_LIB=tools/_ab/NAME.so python tools/hass_time.py [hass] [dngop] [p1]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timeit(torch, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import numpy as np
    import torch
    import rawspeed_b200 as rs
    from oracle import port, synth
    what = sys.argv[1:] or ["hass", "dngop"]
    ctx = rs.Context(0)
    res = {}
    if "hass" in what:
        w, h = 8272, 6200
        himg = np.tile(synth.image_model(w, 200, seed=41, bits=14), (h // 200, 1))
        hht = port.Huff(synth.DEFAULT_NCPL, synth.DEFAULT_VALUES, full=False)
        cache = os.path.join(ROOT, "tools", "_ab", "hass_frame_8272x6200.npy")  # (the encoder takes ~50 s)
        if os.path.exists(cache):
            hdata = np.load(cache)
        else:
            hdata = synth.make_hasselblad_fast(himg, hht, 0x8000)
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.save(cache, hdata)
        hj = rs.HasselbladJob()
        hj.in_offset, hj.in_size, hj.width, hj.height = 0, hdata.size, w, h
        hj.out_pitch, hj.out_offset, hj.init_pred, hj.table = rs.image_pitch(w), 0, 0x8000, 0
        plan = rs.hasselblad_plan(ctx, [rs.huff_table(bytes(synth.DEFAULT_NCPL), bytes(synth.DEFAULT_VALUES), False)], [hj])
        d_in = torch.zeros(hdata.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:hdata.size] = torch.from_numpy(hdata)
        d_out = torch.zeros(h * rs.image_pitch(w), dtype=torch.uint8, device="cuda")
        plan.run((d_in.data_ptr(), hdata.size), d_out)
        st = plan.results()
        got = d_out.cpu().numpy().view(np.uint16).reshape(h, rs.image_pitch(w) // 2)
        ok = st[0][0] == 0 and bool(np.array_equal(got[:, :w], himg))
        ms = timeit(torch, lambda: plan.run((d_in.data_ptr(), hdata.size), d_out))
        res["hasselblad_8272x6200"] = {"ms": round(ms, 4), "GPix/s": round(w * h / ms / 1e6, 1), "exact": ok,
                                       "launches": plan.launches}
        del plan, d_in, d_out
    if "dngop" in what:
        from rawspeed_b200 import host
        W, H = 8256, 5504
        rng = np.random.default_rng(7)
        pitch = rs.image_pitch(W)
        base = port.new_image(W, H)
        base[:, :W] = synth.image_model(W, H, 3)
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
        plan = rs.dngop_plan(ctx, [dj], low["ops"], low["tables"], low["deltas"])
        d_base = torch.from_numpy(base.view(np.uint8).reshape(-1)).cuda()
        NB = 16  # every timed run gets a fresh copy (the pass works in place; 1.45 GB: nothing stays in L2)
        d_imgs = [d_base.clone() for _ in range(NB)]
        d_chk = d_base.clone()
        plan.run(None, d_chk)
        got = d_chk.cpu().numpy().view(np.uint16).reshape(base.shape)
        ok = bool(np.array_equal(got, want))
        it = [0]

        def step():
            plan.run(None, d_imgs[it[0] % NB])
            it[0] += 1
        ms = timeit(torch, step, reps=12, warm=4)
        res["dngop_8256x5504_8ops"] = {"ms": round(ms, 4), "GPix/s": round(W * H / ms / 1e6, 1), "exact": ok}
    if "p1" in what:
        w, h = 11608, 8708
        rowimg = synth.image_model(w, 4, seed=31, bits=14).astype(np.uint16)
        rows4 = [np.frombuffer(synth.phaseone_row(rowimg[k]), dtype=np.uint8) for k in range(4)]
        offs, blobs, pos = [], [], 0
        for r in range(h):
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
        d_in = torch.zeros(blob.size + 64, dtype=torch.uint8, device="cuda")
        d_in[:blob.size] = torch.from_numpy(blob)
        d_out = torch.zeros(h * rs.image_pitch(w), dtype=torch.uint8, device="cuda")
        vers = ("3", "3w8", "3w3", "2")
        if os.environ.get("RSB200_P1W"):   # (one variant only: profiling runs)
            vers = ("3w" + os.environ["RSB200_P1W"],)
        for ver in vers:
            os.environ["RSB200_P1"] = ver[0]   # (read when the plan is created)
            os.environ["RSB200_P1W"] = ver[2] if len(ver) == 3 else "0"
            plan = rs.phaseone_plan(ctx, [pj], pstrips)
            d_out.zero_()
            plan.run((d_in.data_ptr(), blob.size), d_out)
            st = plan.results()
            got = d_out.cpu().numpy().view(np.uint16).reshape(h, rs.image_pitch(w) // 2)
            ok = st[0][0] == 0 and all(bool(np.array_equal(
                got[k::4, :w], np.broadcast_to(rowimg[k], (len(range(k, h, 4)), w)))) for k in range(4))
            ms = timeit(torch, lambda: plan.run((d_in.data_ptr(), blob.size), d_out), reps=5, warm=2)
            res["phaseone_11608x8708_v" + ver] = {"ms": round(ms, 4), "GPix/s": round(w * h / ms / 1e6, 1),
                                                  "exact": ok, "launches": plan.launches}
            del plan
        os.environ.pop("RSB200_P1", None)
        os.environ.pop("RSB200_P1W", None)
    print("HT " + os.path.basename(os.environ.get("RSB200_LIB", "default")) + " " + json.dumps(res))


if __name__ == "__main__":
    main()
