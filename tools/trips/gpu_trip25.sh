#!/bin/bash
# setmaxnreg + double-buffered tcgen05.ld in both epilogue warpgroups of the dual-direction screen
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_kiez.py -m gpu -q -x --timeout 600 -k "fused or screen or dual" > gpurun_out/pytest_fused.log 2>&1; echo "pytest(fused) exit $?"; tail -5 gpurun_out/pytest_fused.log
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}).get('emitted_per_column_mean'))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 2"
b c4_setmaxnreg "$C4 --fused on --precision screen"
KB2_LIB=/root/repo/build/lib_noreg/libkiez_b200.so b c4_noreg "$C4 --fused on --precision screen"
b c4_setmaxnreg_2 "$C4 --fused on --precision screen"
KB2_LIB=/root/repo/build/lib_noreg/libkiez_b200.so b c4_noreg_2 "$C4 --fused on --precision screen"
