#!/usr/bin/env bash
# round 2, GPU call 35: ncu --set full of p1_walk_kernel in its default ("lines") form.
set -u
OUT=gpurun_out/r2_run35
mkdir -p "$OUT"
RSB200_P1W=7 timeout 100 ncu --set full --clock-control none --import-source on -k regex:"p1_walk_kernel" -c 1 \
    -o "$OUT/p1w" python tools/hass_time.py p1 > "$OUT/ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
if [ -f "$OUT/p1w.ncu-rep" ]; then
  ncu -i "$OUT/p1w.ncu-rep" --page raw --csv > "$OUT/p1w_raw.csv" 2>/dev/null
  ncu -i "$OUT/p1w.ncu-rep" --page source --csv --print-source sass > "$OUT/p1w_source.csv" 2>/dev/null
  rm -f "$OUT/p1w.ncu-rep"
fi
