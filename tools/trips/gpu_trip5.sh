#!/bin/bash
mkdir -p gpurun_out
for cfg in 256x16 256x32 128x16 128x32; do
  echo "== diag $cfg" >> gpurun_out/diag5.log
  KB2_TC_CONFIG=$cfg timeout 300 python tools/diag_knn.py tc >> gpurun_out/diag5.log 2>&1; echo "diag $cfg exit $?"
done
grep -E "DIAG|bad_rows=[1-9]|Error|error" gpurun_out/diag5.log | head -20
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_all.log
for cfg in 256x16 256x32 128x16 128x32; do
  KB2_TC_CONFIG=$cfg timeout 600 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/bench_mid_$cfg.json 2> gpurun_out/bench_mid_$cfg.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mid_$cfg.json')); print('$cfg', 'ms/launch', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'issued', d['roofline']['issued_tf32_tflops'], d['clocks'])"
done
timeout 600 python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c3.json')); print('c3', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_quick.json 2> gpurun_out/bench_c4_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_quick.json')); print('c4', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['clocks'])"
