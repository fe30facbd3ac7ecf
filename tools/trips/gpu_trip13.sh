#!/bin/bash
# final round-1 measurements at 1 GPU: default bench (C4), reference arm, ncu launch list
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_c4.json 2> gpurun_out/bench_ref_c4.err; echo "ref exit $?"; cat gpurun_out/bench_ref_c4.json | cut -c1-600
timeout 1200 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench exit $?"; cat gpurun_out/bench_c4.json; tail -3 gpurun_out/bench_c4.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --workload c5 --n 200000 --m 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_small.json 2> gpurun_out/bench_c5_small.err; echo "c5-small exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_small.json')); print('c5small', 'q/s', d['value'], 'ms/step', d['ms_per_step'], 'ms/launch', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], d['clocks'])"; tail -3 gpurun_out/bench_c5_small.err
