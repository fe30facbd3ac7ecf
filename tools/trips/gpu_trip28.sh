#!/bin/bash
# L2 blocking of the dual-direction screen: index range size (default 24 MB)
mkdir -p gpurun_out
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}).get('emitted_per_column_mean'))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 1 --no-hub-scores"
for mb in 6 12 24 48; do
KB2_SCREEN_RANGE_MB=$mb b c4_range$mb "$C4 --fused on --precision screen"
done
KB2_SCREEN_RANGE_MB=12 b c4_twopass_range12 "$C4 --fused off --precision screen"
