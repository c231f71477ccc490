#!/usr/bin/env bash
# round 2, GPU call 9: the whole GPU suite on the current tree, the default bench, NT=512 A/B, PCIe probe.
set -u
OUT=gpurun_out/r2_run9
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/test_gpu.log" 2>&1
echo "pytest -m gpu exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu.log"
timeout 300 python tools/pcie_probe.py > "$OUT/pcie_probe.log" 2>&1
echo "pcie exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/pcie_probe.log"
TILE_AB_ONLY=tile_r1 timeout 600 python tools/tile_ab.py 1,8,20 > "$OUT/tile_nt256.log" 2>&1
echo "tile nt256 exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/tile_nt256.log"
RSB200_LIB=$PWD/tools/_ab/tile_nt512.so TILE_AB_ONLY=tile_r1 timeout 600 python tools/tile_ab.py 1,8,20 > "$OUT/tile_nt512.log" 2>&1
echo "tile nt512 exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/tile_nt512.log"
( time timeout 1500 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err" ) 2> "$OUT/bench_default.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench_default.time"; tail -c 3000 "$OUT/bench_default.json"
( time timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err" ) 2> "$OUT/bench_reference.time"
echo "bench reference exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench_reference.time"; tail -c 1500 "$OUT/bench_reference.json"
