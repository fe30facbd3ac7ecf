#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -q -x --timeout 600 -k "screen" > gpurun_out/pytest_screen.log 2>&1; rc=$?; echo "pytest(screen) exit $rc"; tail -25 gpurun_out/pytest_screen.log
if [ $rc -eq 0 ]; then
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -8 gpurun_out/pytest_all.log
fi
b() { # name, extra env, args
  env $2 timeout 900 python bench.py $3 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms', round(r['avg_launch_ms'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}).get('overflow_columns'))"; tail -2 gpurun_out/b_$1.err; }
MID="--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 3 --warmup 2 --no-hub-scores"
b mid_screen_twopass "A=1" "$MID --fused off --precision screen"
b mid_screen_fused "A=1" "$MID --fused on --precision screen"
b mid_x3_twopass "A=1" "$MID --fused off --precision tf32x3"
C4="--steps 2 --warmup 1"
b c4_screen_fused "A=1" "$C4 --fused on --precision screen"
b c4_screen_twopass "A=1" "$C4 --fused off --precision screen"
b c4_x3_fused "A=1" "$C4 --fused on --precision tf32x3"
