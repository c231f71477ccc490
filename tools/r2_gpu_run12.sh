#!/usr/bin/env bash
# round 2, GPU call 12: the stream kernel's loads (128-byte L2 lines, L2 prefetch ahead), 7 CTAs/SM; LJPEG tests.
set -u
OUT=gpurun_out/r2_run12
mkdir -p "$OUT"
for v in default s_ld0_pf0 s_ld128_pf0 s_ld0_pf8 s_ld128_pf16 s_lb7; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  AB_FRAMES=32,128,256 AB_PATHS=stream timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_$v.log" 2>&1
  echo "ab $v exit $?" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_$v.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    print('  $v', {k: (v['ms'], v['GPix/s'], v['exact']) for k, v in d.items() if k.startswith('dng') and '_stream' in k})
" | tee -a "$OUT/summary.txt"
done
unset RSB200_LIB
timeout 1200 python -m pytest tests/test_gpu_ljpeg.py -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_ljpeg.log"
( time timeout 900 python bench.py --skip-others --skip-cpu > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
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
    print(d["roofline"]["kernel"])
except Exception as ex:
    print("unreadable:", ex)
PY
