#!/bin/bash
# round 2, trip 16 (1 GPU, short): tunables of the dual-direction pass under the final kernels
mkdir -p gpurun_out
b() { timeout 100 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256 > gpurun_out/r2_b16_$1.json 2> gpurun_out/r2_b16_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b16_$1.json')); r=d['roofline']; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],2), 'parity', d['parity_check'] and d['parity_check']['mismatch'], 'clk', (d.get('clocks') or {}).get('sm_mhz'), [(x['kind'], x['nq'], round(x['avg_launch_ms'],2)) for x in r['search_launches'][:5]], (r.get('dual_direction') or {}).get('emitted_per_column_mean'))"; }
b default
KB2_FUSED_SAMPLE_DIV=48 b div48
KB2_FUSED_GROWTH=4 b growth4
KB2_FUSED_GROWTH=2 b growth2
KB2_SCREEN_SLOTS=12 b slots12
b default2
