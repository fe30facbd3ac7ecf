#!/bin/bash
# round 2, trip 2: partially resident screen (c = 50 / 100 at d = 256), overlapped upload (e2e),
# on-device bound test; bench lines for C4, C4 c=50, C2, C3, C5 (NICDM)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest2.log 2>&1; echo "pytest exit $?"
grep -E "max \|screen|dual-direction:|probe fractions|c4 full size|passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest2.log | tail -60
b() { # name, args, timeout
  timeout ${3:-400} python bench.py $2 > gpurun_out/r2_b_$1.json 2> gpurun_out/r2_b_$1.err; echo "bench $1 exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_b_$1.json')); r=d['roofline']
    print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), r['kernel'][:28], 'frac', round(r['frac'],3), 'ach', round(r['achieved'],1), 'clk', (d['clocks'] or {}).get('sm_mhz'))
    print('   e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','ms_per_step','pinned','fraction_of_device_value')})
    print('   parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first','seconds')})
    print('   screen', r['screen'], 'launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]])
    if d.get('data_variants'): print('   variants', {k:(round(v['value']), v['screen'], v['parity_check'] and v['parity_check']['mismatch']) for k,v in d['data_variants'].items()})
except Exception as e:
    print('$1 failed', e)
PY
  tail -2 gpurun_out/r2_b_$1.err; }
b c4 "--steps 5 --warmup 3 --no-cpu-baseline"
b c4_c50 "--steps 2 --warmup 1 --c 50 --no-cpu-baseline --no-variants --no-e2e"
b c4_c50_fused "--steps 2 --warmup 1 --c 50 --fused on --no-cpu-baseline --no-variants --no-e2e"
b c2 "--workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-variants --e2e-steps 5"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants --e2e-steps 3"
b c5 "--workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256" 600
