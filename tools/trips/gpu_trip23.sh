#!/bin/bash
# pipelined column emits: correctness (fused tests), C4 A/B, ncu of a C4-like dual launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_kiez.py -m gpu -q -x --timeout 600 -k "fused or screen or dual" > gpurun_out/pytest_fused.log 2>&1; echo "pytest(fused) exit $?"; tail -3 gpurun_out/pytest_fused.log
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms', round(r['avg_launch_ms'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 2"
b c4_screen_fused_pipelined "$C4 --fused on --precision screen"
KB2_EMIT_TARGET=192 b c4_screen_fused_emit192 "$C4 --fused on --precision screen"
KB2_EMIT_TARGET=768 b c4_screen_fused_emit768 "$C4 --fused on --precision screen"
b c4_x3_fused_pipelined "$C4 --fused on --precision tf32x3"
prof() { # name, kernel regex, bench args
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_$1 -f python bench.py $3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"; tail -2 gpurun_out/ncu_$1.log | cut -c1-200; }
prof screen_dual_c4like knn_screen "--workload custom --n 1000000 --m 65536 --d 256 --c 10 --k 10 --fused on --precision screen"
