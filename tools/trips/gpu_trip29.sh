#!/bin/bash
# vote-free slow paths (column emits: per-lane masks + one ballot per round; row appends: sparse path)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], [round(s['avg_launch_ms'],1) for s in r['search_launches']])"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 1 --no-hub-scores"
b c4_new "$C4"
KB2_LIB=/root/repo/build/lib_base/libkiez_b200.so b c4_base "$C4"
b c4_new_twopass "$C4 --fused off"
KB2_LIB=/root/repo/build/lib_base/libkiez_b200.so b c4_base_twopass "$C4 --fused off"
b c4_new_x3 "$C4 --precision tf32x3"
b c2_new "--workload c2 --steps 3 --warmup 3"
KB2_LIB=/root/repo/build/lib_base/libkiez_b200.so b c2_base "--workload c2 --steps 3 --warmup 3"
b c3_new "--workload c3 --steps 3 --warmup 3"
KB2_LIB=/root/repo/build/lib_base/libkiez_b200.so b c3_base "--workload c3 --steps 3 --warmup 3"
