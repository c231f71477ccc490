#!/usr/bin/env bash
# round 2, GPU call 11: stream kernel variants (straight-line unit, L2 prefetch, 7 CTAs/SM), a timeline of the
# host-buffer pipeline, ncu of the stream kernel at 256 frames, ncu metrics of the post-decode kernels.
set -u
OUT=gpurun_out/r2_run11
mkdir -p "$OUT"
for v in default s_base s_pf4 s_pf8 s_lb7 s_lb7pf4; do
  if [ "$v" = default ]; then unset RSB200_LIB; else export RSB200_LIB=$PWD/tools/_ab/$v.so; fi
  AB_FRAMES=32,128,256 AB_PATHS=stream timeout 600 python tools/ab_ljpeg.py one > "$OUT/ab_$v.log" 2>&1
  echo "ab $v exit $?" | tee -a "$OUT/summary.txt"
  grep "^AB" "$OUT/ab_$v.log" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 2)[2])
    print('  $v', {k: (v['ms'], v['GPix/s'], v['exact']) for k, v in d.items() if k.startswith('dng') and '_stream' in k})
" | tee -a "$OUT/summary.txt"
done
unset RSB200_LIB
E2E_TRACE=1 E2E_GROUPS=0,4,16 timeout 600 python tools/e2e_probe.py 1,16 > "$OUT/e2e_probe.log" 2>&1
echo "e2e probe exit $?" | tee -a "$OUT/summary.txt"; grep -v "^PIPE_TRACE  *[0-9]" "$OUT/e2e_probe.log" | tail -20
AB_FRAMES=256 AB_PATHS=stream timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k2_stream_kernel" -c 1 \
    -o "$OUT/k2s_256frames" python tools/ab_ljpeg.py one > "$OUT/ncu_256.log" 2>&1
echo "ncu 256 exit $?" | tee -a "$OUT/summary.txt"
[ -f "$OUT/k2s_256frames.ncu-rep" ] && ncu -i "$OUT/k2s_256frames.ncu-rep" --page raw --csv > "$OUT/k2s_256frames_raw.csv" 2>/dev/null
[ -f "$OUT/k2s_256frames.ncu-rep" ] && ncu -i "$OUT/k2s_256frames.ncu-rep" --page source --csv > "$OUT/k2s_256frames_source.csv" 2>/dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,launch__grid_size,launch__block_size,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
timeout 1500 ncu --metrics $M --clock-control none -k regex:"scale_kernel|dngop_kernel|lookup_kernel|badpix|pana4|pana_kernel|arw2_kernel|p1_kernel|sraw_kernel|rawform" --csv --log-file "$OUT/postdecode_ncu.csv" \
    python bench.py --steps 2 --warmup 1 --total-frames 8 --all-legs --unvalidated --skip-single --skip-cpu > "$OUT/postdecode_bench_under_ncu.log" 2>&1
echo "ncu postdecode exit $?" | tee -a "$OUT/summary.txt"
wc -l "$OUT/postdecode_ncu.csv"
