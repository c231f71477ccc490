#!/usr/bin/env bash
# round 2, GPU call 25: first contact of the third Phase One kernel (tests with versions 3 and 2, timing of both),
# the LJPEG suite with the random-table fuzz on every path, A/B of the stream kernel's fill variant,
# per-kernel durations of one Hasselblad frame.
set -u
OUT=gpurun_out/r2_run25
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_phaseone.py -q > "$OUT/test_gpu_phaseone.log" 2>&1
echo "test_gpu_phaseone exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_phaseone.log"
timeout 300 python tools/hass_time.py p1 > "$OUT/ht_p1.log" 2>&1
echo "ht p1 exit $?" | tee -a "$OUT/summary.txt"; grep "^HT" "$OUT/ht_p1.log" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/ht_p1.log"
timeout 400 python -m pytest tests/test_gpu_ljpeg.py -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_ljpeg.log"
for v in default s_p1_fill2; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  AB_ONLY=batch AB_FRAMES=256 AB_PATHS=stream timeout 300 python tools/ab_ljpeg.py one > "$OUT/ab_$v.log" 2>&1
  echo "ab $v exit $?" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_$v.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    print('  $v', {k: (v['ms'], v.get('GPix/s'), v['exact']) for k, v in d.items()})
" | tee -a "$OUT/summary.txt"
done
RSB200_LIB=$PWD/tools/_ab/s_p1_fill2.so timeout 300 python -m pytest tests/test_gpu_ljpeg.py -q -x -k "stream or auto" > "$OUT/test_gpu_ljpeg_fill2.log" 2>&1
echo "test_gpu_ljpeg fill2 exit $?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/test_gpu_ljpeg_fill2.log"
unset RSB200_LIB
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$OUT/hass_launches.csv" \
    python tools/hass_time.py hass > "$OUT/ncu_hass.log" 2>&1
echo "ncu hass exit $?" | tee -a "$OUT/summary.txt"
