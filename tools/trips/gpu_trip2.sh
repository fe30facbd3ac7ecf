#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 exit $?"; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_quick.json 2> gpurun_out/bench_c4_quick.err; echo "c4 exit $?"; cat gpurun_out/bench_c4_quick.json; tail -5 gpurun_out/bench_c4_quick.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c2.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_c2.log
