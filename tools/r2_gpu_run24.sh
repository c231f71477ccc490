#!/usr/bin/env bash
# round 2, GPU call 24 (second session): state of the tree after the two-table fix of k2_stream_kernel, the
# straight-line unit without hit bookkeeping and the third opcode walk; A/B of the stream kernel's
# FMA-pipe variants, of the Hasselblad segment size and of the opcode walks.
set -u
OUT=gpurun_out/r2_run24
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/test_gpu_all.log" 2>&1
echo "pytest -m gpu (all) exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/test_gpu_all.log"
for v in base_run23 default s_pipe1 s_pipe2 s_p2_lb7; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  AB_ONLY=batch AB_FRAMES=64,256 AB_PATHS=stream timeout 300 python tools/ab_ljpeg.py one > "$OUT/ab_$v.log" 2>&1
  echo "ab $v exit $?" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_$v.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    print('  $v', {k: (v['ms'], v.get('GPix/s'), v['exact']) for k, v in d.items()})
" | tee -a "$OUT/summary.txt"
done
for v in s_pipe1 s_pipe2 s_p2_lb7; do
  RSB200_LIB=$PWD/tools/_ab/$v.so timeout 300 python -m pytest tests/test_gpu_ljpeg.py -q -x -k "stream or auto" > "$OUT/test_gpu_ljpeg_$v.log" 2>&1
  echo "test_gpu_ljpeg $v exit $?" | tee -a "$OUT/summary.txt"; tail -1 "$OUT/test_gpu_ljpeg_$v.log"
done
for v in default hass8192 hass16384 dngop_v2; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  what="hass"; [ "$v" = default ] && what="hass dngop"; [ "$v" = dngop_v2 ] && what="dngop"
  timeout 300 python tools/hass_time.py $what > "$OUT/ht_$v.log" 2>&1
  echo "ht $v exit $?" | tee -a "$OUT/summary.txt"; grep "^HT" "$OUT/ht_$v.log" | tee -a "$OUT/summary.txt"
done
unset RSB200_LIB
RSB200_LIB=$PWD/tools/_ab/hass8192.so timeout 300 python -m pytest tests/test_gpu_hasselblad.py -q -x > "$OUT/test_hass8192.log" 2>&1
echo "test_gpu_hasselblad hass8192 exit $?" | tee -a "$OUT/summary.txt"
