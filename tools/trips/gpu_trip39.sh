#!/bin/bash
# screen probe policy: tests, C4 on the hubby distribution, C4 default regression check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -x --timeout 500 -k "probe or screen or fused" > gpurun_out/pytest_probe.log 2>&1; echo "pytest(probe) exit $?"; tail -4 gpurun_out/pytest_probe.log
b() { # name, args
  timeout 400 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), r['kernel'][:30], 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], r['dual_direction'])"; tail -2 gpurun_out/b_$1.err; }
b c4_hubby "--data hubby --steps 1 --warmup 1"
b c4_hubby_forced_screen "--data hubby --steps 1 --warmup 1 --precision screen"
b c4_default_check "--steps 2 --warmup 1"
