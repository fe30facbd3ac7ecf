#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --fused off > gpurun_out/bench_c4_twopass.json 2> gpurun_out/bench_c4_twopass.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_twopass.json')); r=d['roofline']; print('c4 two-pass', 'q/s', d['value'], 'ms/step', d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['clocks'])"; tail -3 gpurun_out/bench_c4_twopass.err
timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_fused.json 2> gpurun_out/bench_c4_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_fused.json')); r=d['roofline']; print('c4 fused', 'q/s', d['value'], 'ms/step', d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['search_launches'], r['dual_direction'], d['e2e'], d['clocks'])"; tail -3 gpurun_out/bench_c4_fused.err
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --fused on > gpurun_out/bench_c2_fused.json 2> gpurun_out/bench_c2_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_fused.json')); r=d['roofline']; print('c2 fused', 'q/s', d['value'], 'ms/step', d['ms_per_step'], r['search_launches'], r['dual_direction'])"; tail -3 gpurun_out/bench_c2_fused.err
timeout 600 python bench.py --workload c5 --n 200000 --m 2000000 --steps 1 --warmup 1 --no-cpu-baseline --fused on > gpurun_out/bench_c5_small_fused.json 2> gpurun_out/bench_c5_small_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_small_fused.json')); r=d['roofline']; print('c5small fused', 'q/s', d['value'], 'ms/step', d['ms_per_step'], r['search_launches'], r['dual_direction'])"; tail -3 gpurun_out/bench_c5_small_fused.err
