#!/usr/bin/env bash
# round 2, GPU call 32: state of record of the round -- GPU suite, smoke, default bench, ncu --set full of the
# stream kernel (256 frames), numbers of record of the secondary legs, launch list of the bench command.
set -u
OUT=gpurun_out/r2_run32
mkdir -p "$OUT"
timeout 300 python -m pytest tests -m gpu -q > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/test_gpu_all.log"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/smoke.log"
( time timeout 330 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/bench.err"; cat "$OUT/bench.time"
python - "$OUT/bench.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("single_frame") or {}
    print("value %.0f MPix/s  ms/step %.2f  frac %.4f  launches %s  e2e %.0f MPix/s (%.1f ms)  single %.0f MPix/s e2e %.0f pageable %.0f mirror %.0f" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        s.get("MPixels/s", 0), (s.get("e2e") or {}).get("value", 0), ((s.get("e2e") or {}).get("pageable") or {}).get("value", 0),
        (s.get("e2e_host_mirror") or {}).get("value", 0)))
    print(d["cpu_baseline"]); print(d.get("clocks"))
except Exception as ex:
    print("unreadable:", ex)
PY
AB_ONLY=batch AB_FRAMES=256 AB_PATHS=stream timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k2_stream_kernel" -c 1 \
    -o "$OUT/k2s_256frames" python tools/ab_ljpeg.py one > "$OUT/ncu_256.log" 2>&1
echo "ncu 256 exit $?" | tee -a "$OUT/summary.txt"
if [ -f "$OUT/k2s_256frames.ncu-rep" ]; then
  ncu -i "$OUT/k2s_256frames.ncu-rep" --page raw --csv > "$OUT/k2s_256frames_raw.csv" 2>/dev/null
  rm -f "$OUT/k2s_256frames.ncu-rep"
fi
timeout 200 python bench.py --steps 3 --warmup 3 --total-frames 8 --all-legs --unvalidated --skip-single --skip-cpu > "$OUT/bench_all_legs.json" 2> "$OUT/bench_all_legs.err"
echo "all legs exit $?" | tee -a "$OUT/summary.txt"
python - "$OUT/bench_all_legs.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and "MPixels/s" in v:
            print("%-62s %9.1f GPix/s frac %s exact %s" % (k[:62], v.get("MPixels/s", 0) / 1e3, v.get("roofline_frac"), v.get("bit_exact")))
except Exception as ex:
    print("unreadable:", ex)
PY
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_bench.csv" \
    python bench.py --steps 2 --warmup 3 --total-frames 32 --skip-others --skip-cpu --skip-single > "$OUT/ncu_bench.log" 2>&1
echo "ncu launch list exit $?" | tee -a "$OUT/summary.txt"
