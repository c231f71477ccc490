#!/usr/bin/env bash
# round 2, GPU call 2: first contact of k2_tile_kernel with a B200: parity, A/B timing, phase shares.
set -u
OUT=gpurun_out/r2_run2
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -x -q > "$OUT/test_gpu_ljpeg.log" 2>&1
echo "test_gpu_ljpeg exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/test_gpu_ljpeg.log"
timeout 600 python tools/tile_ab.py 1,8,32 > "$OUT/tile_ab.log" 2>&1
echo "tile_ab exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_ab.log" | tail -40 | tee -a "$OUT/summary.txt"
RSB200_LIB=tools/_ab/tile_phases.so TILE_AB_PHASES=1 TILE_AB_ONLY=tile_r1,tile_r2,tile_r2_pre0 timeout 300 python tools/tile_ab.py 1,8 > "$OUT/tile_phases.log" 2>&1
echo "tile_phases exit $?" | tee -a "$OUT/summary.txt"
grep -v "^TILE_AB" "$OUT/tile_phases.log" | tail -20 | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_cr2.py tests/test_abi.py -m gpu -x -q > "$OUT/test_gpu_host.log" 2>&1
echo "test_gpu_host exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/test_gpu_host.log"
