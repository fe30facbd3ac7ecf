#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_knn.py simt tc > gpurun_out/diag.log 2>&1; echo "diag exit $?"; tail -14 gpurun_out/diag.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_all.log
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench_c3.err
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_quick.json 2> gpurun_out/bench_c4_quick.err; echo "c4 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_quick.json')); print('c4', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['clocks'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 2 -c 1 -o gpurun_out/prof_knn_tc python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
