#!/usr/bin/env bash
# round 2, GPU call 21: k2_par_kernel with the slices staged in shared memory; K2C vs K2C2 as its pre-pass.
set -u
OUT=gpurun_out/r2_run21
mkdir -p "$OUT"
export RSB200_LIB=$PWD/tools/_ab/par_v3.so
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -q -k "par" -x > "$OUT/test_gpu_ljpeg_par.log" 2>&1
echo "test_gpu_ljpeg par exit $?" | tee -a "$OUT/summary.txt"; tail -8 "$OUT/test_gpu_ljpeg_par.log"
for c in 1 2; do
  RSB200_CLEAN=$c AB_FRAMES=1,8,32 AB_PATHS=par AB_KERNELS=1 timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_par_clean$c.log" 2>&1
  echo "ab clean=$c exit $?" | tee -a "$OUT/summary.txt"
  grep "KERNEL dng.*k2_par_kernel\|KERNEL dng.*k2_clean" "$OUT/ab_par_clean$c.log" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_par_clean$c.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    for k, v in d.items():
        if k.startswith('dng') and '_par' in k:
            print('  %-16s %8.4f ms %7.1f GPix/s exact=%s launches=%s' % (k, v['ms'], v['GPix/s'], v['exact'], v.get('launches')))
" | tee -a "$OUT/summary.txt"
done
AB_FRAMES=8 AB_PATHS=par timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_par_kernel" -c 1 \
    -o "$OUT/k2p_8frames" python tools/ab_ljpeg.py one > "$OUT/ncu_par.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
[ -f "$OUT/k2p_8frames.ncu-rep" ] && ncu -i "$OUT/k2p_8frames.ncu-rep" --page raw --csv > "$OUT/k2p_8frames_raw.csv" 2>/dev/null
[ -f "$OUT/k2p_8frames.ncu-rep" ] && ncu -i "$OUT/k2p_8frames.ncu-rep" --page source --csv > "$OUT/k2p_8frames_source.csv" 2>/dev/null
