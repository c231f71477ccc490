"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

import rawspeed_b200 as rs
from oracle import port, synth


def parse_ljpeg(blob):
    """Minimal SOI..SOS walk of a (synthetic) LJPEG blob -> dict with frame,
    tables, per-component table ids, DRI and the offset of the entropy data."""
    b = bytes(blob)
    assert b[0:2] == b"\xff\xd8"
    pos = 2
    info = {"tables": {}, "dri": 0}
    while True:
        assert b[pos] == 0xFF, "marker expected"
        m = b[pos + 1]
        ln = (b[pos + 2] << 8) | b[pos + 3]
        seg = b[pos + 4:pos + 2 + ln]
        if m == 0xC3:
            info["prec"] = seg[0]
            info["frame_h"] = (seg[1] << 8) | seg[2]
            info["frame_w"] = (seg[3] << 8) | seg[4]
            info["cps"] = seg[5]
            info["comp_ids"] = [seg[6 + 3 * i] for i in range(seg[5])]
            info["hv"] = [seg[7 + 3 * i] for i in range(seg[5])]
        elif m == 0xC4:
            q = 0
            while q < len(seg):
                tid = seg[q] & 0xF
                ncpl = seg[q + 1:q + 17]
                n = sum(ncpl)
                info["tables"][tid] = (bytes(ncpl), bytes(seg[q + 17:q + 17 + n]))
                q += 17 + n
        elif m == 0xDD:
            info["dri"] = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            sel = {}
            for i in range(ns):
                sel[seg[1 + 2 * i]] = seg[2 + 2 * i] >> 4
            info["table_of_comp"] = [sel[c] for c in info["comp_ids"]]
            info["pred"] = seg[1 + 2 * ns]
            info["data_off"] = pos + 2 + ln
            return info
        pos += 2 + ln


def find_restart_markers(blob, start):
    """Offsets (relative to blob) of FF D0..D7 after `start`."""
    a = np.frombuffer(bytes(blob), dtype=np.uint8)
    idx = np.nonzero((a[start:-1] == 0xFF) & (a[start + 1:] >= 0xD0) & (a[start + 1:] <= 0xD7))[0]
    return [int(i) + start for i in idx]


class TableSet:
    """Deduplicated list of rs.HuffTable for a plan."""

    def __init__(self, fix16=False):
        self.keys, self.tabs, self.fix16 = {}, [], fix16

    def add(self, ncpl, values):
        k = (bytes(ncpl), bytes(values))
        if k not in self.keys:
            self.keys[k] = len(self.tabs)
            self.tabs.append(rs.huff_table(k[0], k[1], self.fix16))
        return self.keys[k]


def dng_ljpeg_scans(t, out_pitch, out_offset=0, in_base=0, tabs=None, fix16=False,
                    img_w=None, img_h=None):
    """Scan descriptors for every tile of a synth.DngTiles (what the host layer's
    AbstractDngDecompressor/LJpegDecoder produce)."""
    tabs = tabs or TableSet(fix16)
    W = img_w or t.w
    H = img_h or t.h
    tiles_x = (W + t.tile_w - 1) // t.tile_w
    scans = []
    for n, (off, ln) in enumerate(zip(t.offsets, t.lengths)):
        blob = t.blob[off:off + ln]
        info = parse_ljpeg(blob)
        ty, tx = divmod(n, tiles_x)
        offx, offy = tx * t.tile_w, ty * t.tile_h
        w = min(t.tile_w, W - offx)
        h = min(t.tile_h, H - offy)
        cps = info["cps"]
        # MCU = (cpp*maxDim.x / frame.w, maxDim.y / frame.h) (LJpegDecoder.cpp:128-141)
        mcu_w = t.cpp * t.tile_w // info["frame_w"]
        mcu_h = t.tile_h // info["frame_h"]
        assert mcu_w * mcu_h == cps
        tids = [tabs.add(*info["tables"][k]) for k in info["table_of_comp"]]
        rows_total = h // mcu_h
        rpr = info["dri"] // info["frame_w"] if info["dri"] else info["frame_h"]
        starts = [info["data_off"]]
        if info["dri"]:
            starts += [m + 2 for m in find_restart_markers(blob, info["data_off"])]
        r0 = 0
        k = 0
        while r0 < rows_total:
            rows = min(rpr, rows_total - r0)
            s = rs.LJpegScan()
            s.in_offset = in_base + off + starts[k]
            s.in_size = ln - starts[k]
            s.rows = rows
            s.frame_w = info["frame_w"]
            s.mcu_w, s.mcu_h = mcu_w, mcu_h
            for c in range(cps):
                s.table[c] = tids[c]
                s.init_pred[c] = 1 << (info["prec"] - 1)
            s.out_offset = out_offset
            s.out_pitch = out_pitch
            s.out_x = t.cpp * offx
            s.out_y = offy + r0 * mcu_h
            s.store_w = t.cpp * w
            scans.append(s)
            r0 += rows
            k += 1
    return tabs, scans


def gpu_run(plan, in_np, out_np):
    """Run a plan with device-resident buffers (torch owns the memory)."""
    import torch
    d_in = torch.zeros(in_np.size + 64, dtype=torch.uint8, device="cuda")
    d_in[:in_np.size] = torch.from_numpy(np.ascontiguousarray(in_np))
    d_out = torch.from_numpy(out_np.view(np.int16).copy()).cuda()
    plan.run((d_in.data_ptr(), in_np.size), d_out)
    torch.cuda.synchronize()
    res = plan.results()
    return d_out.cpu().numpy().view(np.uint16), res


def compile_shared(cmd):
    """Run a compiler command whose output file follows "-o": the library is written under a private name and
    renamed into place, so that test processes running side by side (pytest-xdist) never load a half-written
    file or overwrite one another's output in the middle of a link."""
    import os
    import subprocess
    i = cmd.index("-o")
    out = cmd[i + 1]
    tmp = "%s.tmp.%d" % (out, os.getpid())
    try:
        subprocess.check_call(cmd[:i + 1] + [tmp] + cmd[i + 2:])
        os.replace(tmp, out)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
