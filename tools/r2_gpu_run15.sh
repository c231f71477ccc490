#!/usr/bin/env bash
# round 2, GPU call 15 (N GPUs): bench.py under torchrun exactly as the driver launches it: strong-scaling shard of the
# 256-frame batch + C-ABI NCCL gather; reference arm under torchrun too (rank 0 alone works).
set -u
N=${1:-8}
OUT=gpurun_out/r2_run15
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo_n$N.txt" 2>&1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 --skip-single --skip-others > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err" ) 2> "$OUT/bench_n$N.time"
echo "bench n=$N exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench_n$N.time"; tail -5 "$OUT/bench_n$N.err"
python - "$OUT/bench_n$N.json" <<'PY' | tee -a gpurun_out/r2_run15/summary.txt
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d value %.0f MPix/s  ms/step %.2f  frac %.3f  kernel %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"]))
    print("e2e %.0f MPix/s (%.1f ms/step)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    print("gather", json.dumps(d.get("gather"), indent=1))
    print("clocks", d.get("clocks"))
except Exception as ex:
    print("bench line unreadable:", ex)
PY
