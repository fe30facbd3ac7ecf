#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
b() { # name, extra env, args
  env $2 timeout 900 python bench.py $3 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms', round(r['avg_launch_ms'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], (r['dual_direction'] or {}).get('overflow_columns'))"; tail -2 gpurun_out/b_$1.err; }
MID="--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 3 --warmup 2 --no-hub-scores"
b mid_twopass_sync "KB2_WAVE_SYNC=1" "$MID --fused off"
b mid_twopass_nosync "KB2_WAVE_SYNC=0" "$MID --fused off"
b mid_fused_sync "KB2_WAVE_SYNC=1" "$MID --fused on"
b mid_fused_nosync "KB2_WAVE_SYNC=0" "$MID --fused on"
C4="--steps 2 --warmup 1"
b c4_twopass_sync "KB2_WAVE_SYNC=1" "$C4 --fused off"
b c4_twopass_nosync "KB2_WAVE_SYNC=0" "$C4 --fused off"
b c4_fused_sync "KB2_WAVE_SYNC=1" "$C4 --fused on"
b c4_fused_nosync "KB2_WAVE_SYNC=0" "$C4 --fused on"
for m in "dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum"; do
KB2_WAVE_SYNC=1 timeout 300 ncu --metrics $m --clock-control none -k regex:knn_fused -s 1 -c 1 python bench.py $MID --fused on --no-cpu-baseline --no-e2e 2>&1 | grep -E "dram__bytes|hit_rate|duration" | sed 's/^/fused sync /'
KB2_WAVE_SYNC=0 timeout 300 ncu --metrics $m --clock-control none -k regex:knn_fused -s 1 -c 1 python bench.py $MID --fused on --no-cpu-baseline --no-e2e 2>&1 | grep -E "dram__bytes|hit_rate|duration" | sed 's/^/fused nosync /'
done
