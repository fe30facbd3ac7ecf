#!/bin/bash
# closing check of the round: full GPU suite, smoke, bench line sanity, C4 at c=50
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -3 gpurun_out/pytest_all.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
b() { # name, args
  timeout 300 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), r['kernel'][:30], 'frac', round(r['frac'],3), r['traffic'], d['clocks']['sm_mhz'], r['screen'])"; tail -2 gpurun_out/b_$1.err; }
b c4_check "--steps 2 --warmup 1"
b c4_c50 "--steps 1 --warmup 1 --c 50"
