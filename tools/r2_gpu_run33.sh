#!/usr/bin/env bash
# round 2, GPU call 33: the stream kernel with miss-consistent LUTs (several tables keep the straight-line unit):
# LJPEG suite on every path, timing of the 256-frame batch.
set -u
OUT=gpurun_out/r2_run33
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_gpu_ljpeg.py -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/test_gpu_ljpeg.log"
AB_ONLY=batch AB_FRAMES=256 AB_PATHS=stream timeout 200 python tools/ab_ljpeg.py one > "$OUT/ab_default.log" 2>&1
echo "ab exit $?" | tee -a "$OUT/summary.txt"; grep "^AB" "$OUT/ab_default.log" | cut -c1-400 | tee -a "$OUT/summary.txt"
