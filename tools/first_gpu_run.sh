#!/usr/bin/env bash
# Follow-up GPU run for the kernels written at the end of round 1 (K9 scaling, K10 DNG opcodes,
# K11 bad pixels, K12 table lookup, Panasonic V4).  They passed their first contact with a B200
# bit for bit (profiles/r1_postdecode_first_gpu_run.md, tools/quick_validate.py + quick_time.py);
# what is still owed: their torch-based pytest files, the bench numbers of record, ncu.  One call:
#
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_run.sh'
#
# 1. their gated parity tests (-x: stop at the first failure, the log says which kernel),
# 2. the whole GPU suite with the gates open (nothing else may have moved),
# 3. bench legs (--unvalidated) next to the validated ones,
# 4. launch list + one `ncu --set full` capture per new kernel (read back here with
#    `ncu -i ... --page raw --csv`; summaries go to profiles/).
# Everything lands in gpurun_out/first_run/.
set -u
OUT=gpurun_out/first_run
mkdir -p "$OUT"
export RSB200_UNVALIDATED=1
python -m pytest tests/test_gpu_scale.py tests/test_gpu_lookup.py tests/test_gpu_dngopcodes.py \
       tests/test_gpu_badpixels.py tests/test_gpu_panasonic_v4.py tests/test_examples.py -m gpu -x -q > "$OUT/gated_tests.log" 2>&1
echo "gated tests exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/gated_tests.log"
python -m pytest tests -m gpu -q > "$OUT/all_gpu_tests.log" 2>&1
echo "all gpu tests exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/all_gpu_tests.log"
python bench.py --steps 10 --warmup 3 --unvalidated --skip-cpu > "$OUT/bench_unvalidated.json" 2> "$OUT/bench_unvalidated.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
python - <<'PY' | tee -a gpurun_out/first_run/summary.txt
import json
try:
    d = json.loads(open("gpurun_out/first_run/bench_unvalidated.json").read().strip().splitlines()[-1])
    for k, v in d.get("others", {}).items():
        if "8(f)3" in k or "V4" in k:
            print("%-62s %9.1f GPix/s  frac %.3f  exact %s" % (k, v["MPixels/s"] / 1e3, v.get("roofline_frac", 0), v["bit_exact"]))
except Exception as ex:   # noqa: BLE001
    print("bench line unreadable:", ex)
PY
# A/B of the second opcode walk (build it here first: python tools/ab_ljpeg.py build dngop_v2 -DRSB200_DNGOP_V2)
python tools/quick_time.py > "$OUT/quick_time_v1.log" 2>&1
grep "K10" "$OUT/quick_time_v1.log" | sed 's/^/shipped  /' | tee -a "$OUT/summary.txt"
if [ -f tools/_ab/dngop_v2.so ]; then
  RSB200_LIB=tools/_ab/dngop_v2.so python tools/quick_time.py > "$OUT/quick_time_v2.log" 2>&1
  grep "K10" "$OUT/quick_time_v2.log" | sed 's/^/DNGOP_V2 /' | tee -a "$OUT/summary.txt"
fi
# A/B of the shared-memory-table lookup (run-time switch, same library)
RSB200_LUT_SMEM=1 python tools/quick_time.py > "$OUT/quick_time_lut_smem.log" 2>&1
grep "K12 sixteenBitLookup, plain" "$OUT/quick_time_v1.log" | sed 's/^/shipped   /' | tee -a "$OUT/summary.txt"
grep "K12 sixteenBitLookup, plain" "$OUT/quick_time_lut_smem.log" | sed 's/^/LUT_SMEM  /' | tee -a "$OUT/summary.txt"
if command -v ncu > /dev/null; then
  for k in scale_kernel lookup_kernel dngop_kernel badpix_kernel "pana_kernel<4"; do
    f=$(echo "$k" | tr -cd 'a-z0-9_')
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${k%%<*}" -c 2 \
        -o "$OUT/ncu_$f" python bench.py --steps 2 --warmup 1 --unvalidated --only-unvalidated --skip-cpu \
        > "$OUT/ncu_$f.log" 2>&1 || true
  done
fi
ls -la "$OUT" | tail -20
