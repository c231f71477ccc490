#!/usr/bin/env bash
# round 2, GPU call 14: Hasselblad kernels (first contact), stream kernel with 256-bit loads/stores by launch size,
# the whole GPU suite, default bench + reference arm, ncu of the stream kernel.
set -u
OUT=gpurun_out/r2_run14
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_hasselblad.py -q -x > "$OUT/test_gpu_hasselblad.log" 2>&1
echo "test_gpu_hasselblad exit $?" | tee -a "$OUT/summary.txt"; tail -12 "$OUT/test_gpu_hasselblad.log"
AB_FRAMES=32,64,128,256 AB_PATHS=stream timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_default.log" 2>&1
echo "ab exit $?" | tee -a "$OUT/summary.txt"; grep "^AB" "$OUT/ab_default.log" | cut -c1-700
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/test_gpu_all.log"
( time timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench.err"
python - "$OUT/bench.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("single_frame") or {}
    print("value %.0f MPix/s  ms/step %.2f  frac %.4f  launches %s  e2e %.0f MPix/s (%.1f ms)  single %.0f MPix/s e2e %.0f pageable %.0f mirror %.0f" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        s.get("MPixels/s", 0), (s.get("e2e") or {}).get("value", 0), ((s.get("e2e") or {}).get("pageable") or {}).get("value", 0),
        (s.get("e2e_host_mirror") or {}).get("value", 0)))
    print(d["roofline"]["kernel"]); print(d["cpu_baseline"])
except Exception as ex:
    print("unreadable:", ex)
PY
( time timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err" ) 2> "$OUT/bench_reference.time"
echo "bench reference exit $?" | tee -a "$OUT/summary.txt"; tail -c 700 "$OUT/bench_reference.json"
AB_FRAMES=256 AB_PATHS=stream timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k2_stream_kernel" -c 1 \
    -o "$OUT/k2s_256frames" python tools/ab_ljpeg.py one > "$OUT/ncu_256.log" 2>&1
echo "ncu 256 exit $?" | tee -a "$OUT/summary.txt"
[ -f "$OUT/k2s_256frames.ncu-rep" ] && ncu -i "$OUT/k2s_256frames.ncu-rep" --page raw --csv > "$OUT/k2s_256frames_raw.csv" 2>/dev/null
[ -f "$OUT/k2s_256frames.ncu-rep" ] && ncu -i "$OUT/k2s_256frames.ncu-rep" --page source --csv > "$OUT/k2s_256frames_source.csv" 2>/dev/null
timeout 900 python bench.py --steps 3 --warmup 3 --total-frames 8 --all-legs --unvalidated --skip-single > "$OUT/bench_all_legs.json" 2> "$OUT/bench_all_legs.err"
echo "all legs exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench_all_legs.err"
python - <<'PY' | tee -a gpurun_out/r2_run14/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/r2_run14/bench_all_legs.json").read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and "Hasselblad" in k:
            print(k, v)
except Exception as ex:
    print("unreadable:", ex)
PY
