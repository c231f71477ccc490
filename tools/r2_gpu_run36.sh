#!/usr/bin/env bash
# round 2, GPU call 36: the "cadence" form of the Phase One header walk (RSB200_P1W=8): tests of every form, timing.
set -u
OUT=gpurun_out/r2_run36
mkdir -p "$OUT"
timeout 120 python -m pytest tests/test_gpu_phaseone.py -q > "$OUT/test_gpu_phaseone.log" 2>&1
echo "test_gpu_phaseone exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/test_gpu_phaseone.log"
timeout 120 python tools/hass_time.py p1 > "$OUT/ht_p1.log" 2>&1
echo "ht p1 exit $?" | tee -a "$OUT/summary.txt"; grep "^HT" "$OUT/ht_p1.log" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/ht_p1.log" | cut -c1-300
