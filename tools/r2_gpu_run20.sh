#!/usr/bin/env bash
# round 2, GPU call 20: k2_par_kernel with persistent CTAs (2..5 per SM): does the L1 keep the slices?
set -u
OUT=gpurun_out/r2_run20
mkdir -p "$OUT"
export RSB200_LIB=$PWD/tools/_ab/par_persist.so
timeout 900 python -m pytest tests/test_gpu_ljpeg.py -q -k "par" -x > "$OUT/test_gpu_ljpeg_par.log" 2>&1
echo "test_gpu_ljpeg par exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_ljpeg_par.log"
for c in 2 3 4 5; do
  RSB200_PAR_CTAS=$c AB_FRAMES=1,8,32 AB_PATHS=par AB_KERNELS=1 timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_par_c$c.log" 2>&1
  echo "ab ctas=$c exit $?" | tee -a "$OUT/summary.txt"
  grep "KERNEL dng.*k2_par_kernel\|KERNEL dng.*k2_clean" "$OUT/ab_par_c$c.log" | tee -a "$OUT/summary.txt"
done
# scale / lookup kernels with column segments (same variant library): parity + timing
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_lookup.py tests/test_gpu_postdecode.py -q > "$OUT/test_gpu_scale_lookup.log" 2>&1
echo "test scale/lookup exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_scale_lookup.log"
timeout 900 python bench.py --steps 5 --warmup 3 --total-frames 8 --all-legs --unvalidated --skip-single --skip-cpu > "$OUT/bench_all_legs.json" 2> "$OUT/bench_all_legs.err"
echo "all legs exit $?" | tee -a "$OUT/summary.txt"
python - <<'PY' | tee -a gpurun_out/r2_run20/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/r2_run20/bench_all_legs.json").read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and ("scale" in k or "Lookup" in k or "lookup" in k):
            print("%-70s %9.1f GPix/s frac %s exact %s" % (k[:70], v.get("MPixels/s", 0) / 1e3, v.get("roofline_frac"), v.get("bit_exact")))
except Exception as ex:
    print("unreadable:", ex)
PY
