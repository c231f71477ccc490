#!/usr/bin/env bash
# round 2, GPU call 29: ncu --set full of p1_walk_kernel and p1_decode_kernel (one launch each).
set -u
OUT=gpurun_out/r2_run29
mkdir -p "$OUT"
RSB200_P1W=3 timeout 280 ncu --set full --clock-control none --import-source on -k regex:"p1_walk_kernel|p1_decode_kernel" -c 2 \
    -o "$OUT/p1" python tools/hass_time.py p1 > "$OUT/ncu_p1.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
if [ -f "$OUT/p1.ncu-rep" ]; then
  ncu -i "$OUT/p1.ncu-rep" --page raw --csv > "$OUT/p1_raw.csv" 2>/dev/null
  ncu -i "$OUT/p1.ncu-rep" --page source --csv --print-source sass > "$OUT/p1_source.csv" 2>/dev/null
  ls -la "$OUT" | tee -a "$OUT/summary.txt"
fi
