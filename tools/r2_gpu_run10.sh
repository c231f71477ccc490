#!/usr/bin/env bash
# round 2, GPU call 10: k2_stream_kernel (first contact) against K2C + K2T; e2e with 8 pipeline streams and
# tile groups for host runs of thread-path plans; the LJPEG GPU tests on all six paths.
set -u
OUT=gpurun_out/r2_run10
mkdir -p "$OUT"
timeout 1200 python -m pytest tests/test_gpu_ljpeg.py -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/test_gpu_ljpeg.log"
AB_FRAMES=32,128,256 AB_PATHS=thread,stream AB_KERNELS=1 timeout 900 python tools/ab_ljpeg.py one > "$OUT/ab_thread_stream.log" 2>&1
echo "ab exit $?" | tee -a "$OUT/summary.txt"; grep "KERNEL\|dng[0-9]*_" "$OUT/ab_thread_stream.log" | tail -30
for k in clean stream; do
  ( time RSB200_THREAD_KERNEL=$k timeout 900 python bench.py --skip-others --skip-cpu > "$OUT/bench_$k.json" 2> "$OUT/bench_$k.err" ) 2> "$OUT/bench_$k.time"
  echo "bench $k exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench_$k.err"
  python - "$OUT/bench_$k.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("single_frame") or {}
    print("value %.0f MPix/s  ms/step %.2f  frac %.4f  launches %s  e2e %.0f MPix/s (%.1f ms)  single %.0f MPix/s e2e %.0f pageable %.0f" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        s.get("MPixels/s", 0), (s.get("e2e") or {}).get("value", 0), ((s.get("e2e") or {}).get("pageable") or {}).get("value", 0)))
except Exception as ex:
    print("unreadable:", ex)
PY
done
AB_FRAMES=256 AB_PATHS=thread,stream timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k2_thread_kernel|k2_stream_kernel" -c 2 \
    -o "$OUT/k2t_k2s_256frames" python tools/ab_ljpeg.py one > "$OUT/ncu_256.log" 2>&1
echo "ncu 256 exit $?" | tee -a "$OUT/summary.txt"
[ -f "$OUT/k2t_k2s_256frames.ncu-rep" ] && ncu -i "$OUT/k2t_k2s_256frames.ncu-rep" --page raw --csv > "$OUT/k2t_k2s_256frames_raw.csv" 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/test_gpu_all.log"
