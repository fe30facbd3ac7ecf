#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -x --timeout 300 -k "fused" > gpurun_out/pytest_fused.log 2>&1; echo "pytest(fused) exit $?"; tail -15 gpurun_out/pytest_fused.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_all.log
timeout 900 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_mid_fused.json 2> gpurun_out/bench_mid_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mid_fused.json')); print('mid', 'q/s', d['value'], 'ms/step', d['ms_per_step'], d['roofline'])"; tail -3 gpurun_out/bench_mid_fused.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_fused.json 2> gpurun_out/bench_c4_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_fused.json')); print('c4', 'q/s', d['value'], 'ms/step', d['ms_per_step'], d['roofline'], d['e2e'], d['clocks'])"; tail -3 gpurun_out/bench_c4_fused.err
