#!/usr/bin/env bash
# round 2, GPU call 24: Hasselblad segment size (main library: 16384 bits / 6 rounds; variants 8192 and 4096 bits / 8 rounds).
set -u
OUT=gpurun_out/r2_run24
mkdir -p "$OUT"
for v in default hass8192 hass4096; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_hasselblad.py -q > "$OUT/test_$v.log" 2>&1
  echo "tests $v exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/test_$v.log"
  timeout 600 python bench.py --steps 5 --warmup 3 --total-frames 8 --all-legs --unvalidated --skip-single --skip-cpu > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
  python - "$OUT/bench_$v.json" "$v" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and "Hasselblad" in k:
            print(sys.argv[2], "%-50s %9.1f GPix/s %.3f ms exact %s" % (k[:50], v.get("MPixels/s", 0) / 1e3, v.get("ms_per_frame", 0), v.get("bit_exact")))
except Exception as ex:
    print("unreadable:", ex)
PY
done
