#!/usr/bin/env bash
# round 2, first GPU call: open the gates of the round-1 post-decode tests, whole GPU suite, A/Bs.
set -u
OUT=gpurun_out/r2_run1
mkdir -p "$OUT"
export RSB200_UNVALIDATED=1
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_lookup.py tests/test_gpu_dngopcodes.py \
       tests/test_gpu_badpixels.py tests/test_gpu_panasonic_v4.py tests/test_examples.py -m gpu -q > "$OUT/gated_tests.log" 2>&1
echo "gated tests exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/gated_tests.log"
timeout 900 python -m pytest tests -m gpu -q > "$OUT/all_gpu_tests.log" 2>&1
echo "all gpu tests exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/all_gpu_tests.log"
timeout 300 python tools/quick_time.py > "$OUT/quick_time_v1.log" 2>&1
RSB200_LUT_SMEM=1 timeout 300 python tools/quick_time.py > "$OUT/quick_time_lut_smem.log" 2>&1
grep -h "K1[0-2]\|K9" "$OUT/quick_time_v1.log" | sed 's/^/shipped  /' | tee -a "$OUT/summary.txt"
grep -h "K12" "$OUT/quick_time_lut_smem.log" | sed 's/^/LUT_SMEM /' | tee -a "$OUT/summary.txt"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a "$OUT/summary.txt"
lscpu | head -20 > "$OUT/lscpu.txt"; free -g >> "$OUT/lscpu.txt"
