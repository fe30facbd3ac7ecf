#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 200 python tools/diag_knn.py tc tc1 simt > gpurun_out/diag.log 2>&1; echo "diag exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror" gpurun_out/diag.log | head
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -3 gpurun_out/pytest_all.log
run() { AB_NAME=$1 timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log; }
run pair_cap16
AB_CAP=56 run pair_cap56
AB_CAP=112 run pair_cap112
for wl in c2 c3; do
timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', 'q/s', d['value'], 'ms/step', d['ms_per_step'], 'ms/launch', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"
done
