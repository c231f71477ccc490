#!/usr/bin/env bash
# round 2, GPU call 5: whole GPU suite (drop-in file tests included), default bench (256 frames), reference arm.
set -u
OUT=gpurun_out/r2_run5
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -x > "$OUT/all_gpu_tests.log" 2>&1
echo "all gpu tests exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/all_gpu_tests.log"
( time timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err" ) 2> "$OUT/bench.time"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.time"; tail -3 "$OUT/bench.err"
python - <<'PY' | tee -a gpurun_out/r2_run5/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/r2_run5/bench.json").read().strip().splitlines()[-1])
    print("value %.0f MPix/s  ms/step %.2f  frac %.3f  kernel %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"]))
    print("e2e %.0f MPix/s (%.1f ms/step)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"].get("best"), d["cpu_baseline"].get("worst"))
    s = d["single_frame"]
    print("single: %.0f MPix/s %.3f ms frac %.3f | e2e pinned %.2f ms pageable %.2f ms | mirror %.2f ms" % (
        s["MPixels/s"], s["ms_per_frame"], s["roofline"]["frac"], s["e2e"]["ms_per_frame"], s["e2e"]["pageable"]["ms_per_frame"], s["e2e_host_mirror"]["ms_per_frame"]))
    print("synth s", d["config"]["frame_synthesis_s"])
except Exception as ex:
    print("bench line unreadable:", ex)
PY
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err" ) 2> "$OUT/bench_ref.time"
echo "bench ref exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench_ref.time"; head -c 600 "$OUT/bench_ref.json"
# where the drop-in call spends its time
timeout 300 python tools/dropin_time.py > "$OUT/dropin_time.log" 2>&1
tail -15 "$OUT/dropin_time.log" | tee -a "$OUT/summary.txt"
