#!/usr/bin/env bash
# round 2, GPU call 4: tightened loops; 5-CTA vs 4-CTA geometry; new bench.py first contact.
set -u
OUT=gpurun_out/r2_run4
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -x -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/test_gpu_ljpeg.log"
export TILE_AB_ONLY=fused,tile_r1,tile_r2
timeout 600 python tools/tile_ab.py 1,8,20 > "$OUT/tile_ab.log" 2>&1
echo "tile_ab exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_ab.log" | tail -40 | tee -a "$OUT/summary.txt"
RSB200_LIB=tools/_ab/tile_g4.so TILE_AB_ONLY=tile_r1 timeout 300 python tools/tile_ab.py 1,8,20 > "$OUT/tile_g4.log" 2>&1
echo "tile_g4 exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_g4.log" | tail -8 | tee -a "$OUT/summary.txt"
RSB200_LIB=tools/_ab/tile_g4_phases.so TILE_AB_PHASES=1 TILE_AB_ONLY=tile_r1 timeout 300 python tools/tile_ab.py 1,8 > "$OUT/tile_phases.log" 2>&1
echo "tile_phases exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_phases.log" | tail -20 | tee -a "$OUT/summary.txt"
timeout 1500 python bench.py --steps 5 --warmup 3 --total-frames 32 > "$OUT/bench32.json" 2> "$OUT/bench32.err"
echo "bench32 exit $?" | tee -a "$OUT/summary.txt"
tail -c 3000 "$OUT/bench32.json"; tail -5 "$OUT/bench32.err"
