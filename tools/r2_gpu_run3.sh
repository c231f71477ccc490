#!/usr/bin/env bash
# round 2, GPU call 3: restructured unstuff (cooperative, from the staging), 5 CTAs/SM.
set -u
OUT=gpurun_out/r2_run3
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -x -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"
tail -4 "$OUT/test_gpu_ljpeg.log"
export TILE_AB_ONLY=fused,tile_r1,tile_r1_pre352,tile_r2,tile_r2_pre0
timeout 600 python tools/tile_ab.py 1,8,20 > "$OUT/tile_ab.log" 2>&1
echo "tile_ab exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_ab.log" | tail -40 | tee -a "$OUT/summary.txt"
for v in tile_cta4 tile_big4; do
  RSB200_LIB=tools/_ab/$v.so TILE_AB_ONLY=tile_r1,tile_r1_pre352 timeout 300 python tools/tile_ab.py 1,8,20 > "$OUT/$v.log" 2>&1
  echo "$v exit $?" | tee -a "$OUT/summary.txt"
  grep -v "^TILE_AB" "$OUT/$v.log" | tail -8 | tee -a "$OUT/summary.txt"
done
RSB200_LIB=tools/_ab/tile_phases.so TILE_AB_PHASES=1 TILE_AB_ONLY=tile_r1,tile_r2 timeout 300 python tools/tile_ab.py 1,8 > "$OUT/tile_phases.log" 2>&1
echo "tile_phases exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_phases.log" | tail -20 | tee -a "$OUT/summary.txt"
