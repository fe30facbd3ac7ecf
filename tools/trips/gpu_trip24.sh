#!/bin/bash
# dual-direction pass in row segments with tightening column thresholds: tests + C4 tuning
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_kiez.py -m gpu -q -x --timeout 600 -k "fused or screen or dual" > gpurun_out/pytest_fused.log 2>&1; echo "pytest(fused) exit $?"; tail -15 gpurun_out/pytest_fused.log
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 2"
b c4_seg_default "$C4 --fused on --precision screen"
KB2_FUSED_GROWTH=3 b c4_seg_g3 "$C4 --fused on --precision screen"
KB2_FUSED_SAMPLE_DIV=128 b c4_seg_div128 "$C4 --fused on --precision screen"
KB2_FUSED_SAMPLE_DIV=32 b c4_seg_div32 "$C4 --fused on --precision screen"
KB2_FUSED_GROWTH=1.5 b c4_seg_g15 "$C4 --fused on --precision screen"
b c4_seg_x3 "$C4 --fused on --precision tf32x3"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
