#!/usr/bin/env bash
# round 2, GPU call 23: state of record after the last changes (opcode walk 2, lookup / scale column segments, NUMA
# pinning from sysfs): GPU suite, smoke, default bench; A/B of the shared-memory-table lookup kernel.
set -u
OUT=gpurun_out/r2_run23
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/test_gpu_all.log"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
( time timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench.err"; cat "$OUT/bench.time"
python - "$OUT/bench.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("single_frame") or {}
    print("value %.0f MPix/s  ms/step %.2f  frac %.4f  launches %s  e2e %.0f MPix/s (%.1f ms)  single %.0f MPix/s e2e %.0f pageable %.0f mirror %.0f" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        s.get("MPixels/s", 0), (s.get("e2e") or {}).get("value", 0), ((s.get("e2e") or {}).get("pageable") or {}).get("value", 0),
        (s.get("e2e_host_mirror") or {}).get("value", 0)))
    print(d["cpu_baseline"]); print(d.get("clocks"))
except Exception as ex:
    print("unreadable:", ex)
PY
for v in 0 1; do
  RSB200_LUT_SMEM=$v timeout 300 python -m pytest tests/test_gpu_lookup.py -q > "$OUT/test_lookup_smem$v.log" 2>&1
  echo "lookup tests smem=$v exit $?" | tee -a "$OUT/summary.txt"
  RSB200_LUT_SMEM=$v timeout 600 python bench.py --steps 5 --warmup 3 --total-frames 8 --all-legs --unvalidated --skip-single --skip-cpu > "$OUT/bench_legs_smem$v.json" 2> "$OUT/bench_legs_smem$v.err"
  python - "$OUT/bench_legs_smem$v.json" "$v" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if isinstance(v, dict) and ("Lookup" in k or "scale" in k or "DngOpcodes" in k or "Hasselblad" in k):
            print("smem=%s %-60s %9.1f GPix/s frac %s exact %s" % (sys.argv[2], k[:60], v.get("MPixels/s", 0) / 1e3, v.get("roofline_frac"), v.get("bit_exact")))
except Exception as ex:
    print("unreadable:", ex)
PY
done
