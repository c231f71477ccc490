#!/usr/bin/env bash
# round 2, GPU call 17: k2_par_kernel (first contact): parity tests on the par path, A/B against tile / stream.
set -u
OUT=gpurun_out/r2_run17
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -q -k "par" -x > "$OUT/test_gpu_ljpeg_par.log" 2>&1
echo "test_gpu_ljpeg par exit $?" | tee -a "$OUT/summary.txt"; tail -15 "$OUT/test_gpu_ljpeg_par.log"
AB_FRAMES=1,4,8,16,32,64 AB_PATHS=tile,par,stream AB_KERNELS=1 timeout 900 python tools/ab_ljpeg.py one > "$OUT/ab_par.log" 2>&1
echo "ab exit $?" | tee -a "$OUT/summary.txt"
grep "KERNEL dng1_\|KERNEL dng8_\|KERNEL dng32_" "$OUT/ab_par.log" | head -30
grep "^AB" "$OUT/ab_par.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    for k, v in d.items():
        if k.startswith('dng') and '_' in k and 'tab' not in k:
            print('  %-16s %8.4f ms %7.1f GPix/s exact=%s launches=%s' % (k, v['ms'], v['GPix/s'], v['exact'], v.get('launches')))
" | tee -a "$OUT/summary.txt"
