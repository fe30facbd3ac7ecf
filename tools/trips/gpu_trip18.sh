#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench exit $?"; cat gpurun_out/bench_c4.json | cut -c1-3000; tail -3 gpurun_out/bench_c4.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_fused -s 1 -c 1 -o gpurun_out/prof_knn_fused python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores --fused on > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
