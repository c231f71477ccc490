#!/usr/bin/env bash
# round 2, GPU call 16: 32-bit LUT entries (IMAD.HI field extraction) in the stream kernel's straight-line decode.
set -u
OUT=gpurun_out/r2_run16
mkdir -p "$OUT"
for v in default s_lut32; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  AB_FRAMES=32,64,128,256 AB_PATHS=stream timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_$v.log" 2>&1
  echo "ab $v exit $?" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_$v.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    print('  $v', {k: (v['ms'], v['GPix/s'], v['exact']) for k, v in d.items() if k.startswith('dng') and '_stream' in k})
" | tee -a "$OUT/summary.txt"
done
RSB200_LIB=$PWD/tools/_ab/s_lut32.so timeout 1200 python -m pytest tests/test_gpu_ljpeg.py -q -k "stream or auto" > "$OUT/test_gpu_ljpeg_lut32.log" 2>&1
echo "test_gpu_ljpeg lut32 exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_ljpeg_lut32.log"
