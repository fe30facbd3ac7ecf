#!/bin/bash
# re-tune sample / growth after the cheap emits; launch list of the default step
mkdir -p gpurun_out
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], [round(s['avg_launch_ms'],1) for s in r['search_launches']], (r['dual_direction'] or {}).get('emitted_per_column_mean'))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 1 --no-hub-scores"
b c4_d32_g2 "$C4"
KB2_FUSED_SAMPLE_DIV=64 b c4_d64_g2 "$C4"
KB2_FUSED_SAMPLE_DIV=16 b c4_d16_g2 "$C4"
KB2_FUSED_GROWTH=3 b c4_d32_g3 "$C4"
KB2_FUSED_GROWTH=4 KB2_FUSED_SAMPLE_DIV=64 b c4_d64_g4 "$C4"
KB2_FUSED_GROWTH=1 KB2_FUSED_SAMPLE_DIV=16 KB2_FUSED_COL_CAP=1024 b c4_d16_oneseg "$C4"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_c4.log 2>&1; echo "ncu launches exit $?"
