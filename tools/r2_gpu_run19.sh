#!/usr/bin/env bash
# round 2, GPU call 19: the state of record -- the whole GPU suite, smoke(), the default bench and the reference arm.
set -u
OUT=gpurun_out/r2_run19
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/test_gpu_all.log"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log"
( time timeout 900 python bench.py --impl reference > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err" ) 2> "$OUT/bench_reference.time"
echo "bench reference exit $?" | tee -a "$OUT/summary.txt"; tail -c 600 "$OUT/bench_reference.json"
( time timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench.err"; cat "$OUT/bench.time"
python - "$OUT/bench.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("single_frame") or {}
    print("value %.0f MPix/s  ms/step %.2f  frac %.4f  traffic %s  launches %s  e2e %.0f MPix/s (%.1f ms)  single %.0f MPix/s e2e %.0f pageable %.0f mirror %.0f" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("traffic"), d["gpu_launches"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        s.get("MPixels/s", 0), (s.get("e2e") or {}).get("value", 0), ((s.get("e2e") or {}).get("pageable") or {}).get("value", 0),
        (s.get("e2e_host_mirror") or {}).get("value", 0)))
    print(d["roofline"]["kernel"]); print(d["cpu_baseline"]); print(d.get("clocks"))
except Exception as ex:
    print("unreadable:", ex)
PY
