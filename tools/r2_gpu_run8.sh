#!/usr/bin/env bash
# round 2, GPU call 8: ncu evidence (launch lists + one --set full per LJPEG kernel) and the
# numbers of record of the secondary kernels (--all-legs --unvalidated).
set -u
OUT=gpurun_out/r2_run8
mkdir -p "$OUT"
# launch list of the default bench command, shortened (32 frames: thread path; + the single-frame section)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 2 --warmup 3 --total-frames 32 --skip-others --skip-cpu > "$OUT/bench_under_ncu.log" 2>&1
echo "launch list exit $?" | tee -a "$OUT/summary.txt"
# one full capture per kernel: the tile kernel on one frame, K2C + K2T on a 32-frame batch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_tile_kernel -c 1 -o "$OUT/k2_tile_1frame" \
    python tools/prof_workload.py ljpeg 2 > "$OUT/ncu_tile.log" 2>&1
echo "ncu tile exit $?" | tee -a "$OUT/summary.txt"
AB_FRAMES=32 AB_PATHS=thread timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2_thread|k2_clean" -c 2 \
    -o "$OUT/k2t_k2c_32frames" python tools/ab_ljpeg.py one > "$OUT/ncu_thread.log" 2>&1
echo "ncu thread exit $?" | tee -a "$OUT/summary.txt"
TILE_AB_ONLY=tile_r1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_tile_kernel -c 1 \
    -o "$OUT/k2_tile_20frames" python tools/tile_ab.py 20 > "$OUT/ncu_tile20.log" 2>&1
echo "ncu tile20 exit $?" | tee -a "$OUT/summary.txt"
for f in k2_tile_1frame k2t_k2c_32frames k2_tile_20frames; do
  if [ -f "$OUT/$f.ncu-rep" ]; then
    ncu -i "$OUT/$f.ncu-rep" --page raw --csv > "$OUT/${f}_raw.csv" 2>/dev/null
  fi
done
ls -la "$OUT" | tail -12
# numbers of record of every secondary leg
( time timeout 1500 python bench.py --steps 5 --warmup 3 --total-frames 16 --all-legs --unvalidated --skip-single > "$OUT/bench_all_legs.json" 2> "$OUT/bench_all_legs.err" ) 2> "$OUT/bench_all_legs.time"
echo "all legs exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench_all_legs.time"; tail -3 "$OUT/bench_all_legs.err"
python - <<'PY' | tee -a gpurun_out/r2_run8/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/r2_run8/bench_all_legs.json").read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and ("MPixels/s" in v or "MPixels/s_per_gpu" in v):
            print("%-90s %9.1f GPix/s frac %s exact %s" % (k[:90], v.get("MPixels/s", v.get("MPixels/s_per_gpu", 0)) / 1e3, v.get("roofline_frac"), v.get("bit_exact")))
except Exception as ex:
    print("unreadable:", ex)
PY
