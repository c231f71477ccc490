#!/usr/bin/env bash
# round 2, GPU call 31: the third Phase One kernel with the second form of its walk (aligned-word windows,
# table-driven length codes): tests of every version, timing, per-kernel durations.
set -u
OUT=gpurun_out/r2_run31
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_phaseone.py -q > "$OUT/test_gpu_phaseone.log" 2>&1
echo "test_gpu_phaseone exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_phaseone.log"
timeout 300 python tools/hass_time.py p1 > "$OUT/ht_p1.log" 2>&1
echo "ht p1 exit $?" | tee -a "$OUT/summary.txt"; grep "^HT" "$OUT/ht_p1.log" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/ht_p1.log"
