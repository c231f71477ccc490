#!/usr/bin/env bash
# round 2, GPU call 18: k2_par_kernel with register windows.
set -u
OUT=gpurun_out/r2_run18
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -q -k "par" -x > "$OUT/test_gpu_ljpeg_par.log" 2>&1
echo "test_gpu_ljpeg par exit $?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/test_gpu_ljpeg_par.log"
AB_FRAMES=1,8,32 AB_PATHS=tile,par AB_KERNELS=1 timeout 900 python tools/ab_ljpeg.py one > "$OUT/ab_par.log" 2>&1
echo "ab exit $?" | tee -a "$OUT/summary.txt"
grep "KERNEL dng1_\|KERNEL dng8_\|KERNEL dng32_" "$OUT/ab_par.log" | head -30 | tee -a "$OUT/summary.txt"
grep "^AB" "$OUT/ab_par.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    for k, v in d.items():
        if k.startswith('dng') and '_' in k and 'tab' not in k:
            print('  %-16s %8.4f ms %7.1f GPix/s exact=%s launches=%s' % (k, v['ms'], v['GPix/s'], v['exact'], v.get('launches')))
" | tee -a "$OUT/summary.txt"
AB_FRAMES=8 AB_PATHS=par timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k2_par_kernel" -c 1 \
    -o "$OUT/k2p_8frames" python tools/ab_ljpeg.py one > "$OUT/ncu_par.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
[ -f "$OUT/k2p_8frames.ncu-rep" ] && ncu -i "$OUT/k2p_8frames.ncu-rep" --page raw --csv > "$OUT/k2p_8frames_raw.csv" 2>/dev/null
[ -f "$OUT/k2p_8frames.ncu-rep" ] && ncu -i "$OUT/k2p_8frames.ncu-rep" --page source --csv > "$OUT/k2p_8frames_source.csv" 2>/dev/null
