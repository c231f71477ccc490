"""Summarise an .ncu-rep: key raw metrics + per-phase/per-line instruction and
stall-sample breakdown (CPU-side, reads the report with `ncu -i`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
src_file = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(raw.splitlines()))
hdr, vals = rd[0], (rd[2] if len(rd) > 2 else rd[1])
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum', 'thread_inst_executed_per_inst',
        'issue_stalled', 'launch__waves', 'launch__grid_size', 'shared_mem_per_block_dynamic',
        'dram__throughput.avg.pct', 'lts__t_sector_hit_rate', 'l1tex__t_sector_hit_rate']
for h, v in zip(hdr, vals):
    if any(w in h for w in want) and 'pcsamp' not in h and v:
        print(h, '=', v)
mix = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(mix.splitlines()))
h = None
out = []
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    if len(r) > 8 and r[0] == 'Line No':
        h = r
        ix, tx, sx = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
        continue
    if h and len(r) > ix and r[0] != '':
        try:
            out.append((int(r[sx]), int(r[ix]), int(r[tx]), cur, int(r[0]), r[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(o[1] for o in out) or 1
ts = sum(o[0] for o in out) or 1
print("\ntotal warp-inst", tot, "samples", ts)
print("top lines by stall samples:")
for o in sorted(out, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"smp={100*o[0]/ts:5.1f}% inst={100*o[1]/tot:5.1f}% act={o[2]/max(o[1],1):5.1f} {o[3]}:{o[4]}: {o[5]}")
