#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 200 python tools/diag_knn.py tc tc1 simt > gpurun_out/diag.log 2>&1; echo "diag exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror" gpurun_out/diag.log | head
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -3 gpurun_out/pytest_all.log
run() { AB_NAME=$1 timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log; }
run pair_cap16
AB_CAP=56 run pair_cap56_auto
AB_CAP=56 KB2_TC2_BK=32 run pair_cap56_bk32
AB_CAP=112 run pair_cap112_auto
AB_CAP=112 KB2_TC_MODE=1 run single_cap112
for wl in c2 c3 c4; do
timeout 900 python bench.py --workload $wl --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', 'q/s', d['value'], 'ms/step', d['ms_per_step'], 'ms/launch', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"
done
timeout 600 python tools/bench_kernels.py --n 1000000 --m 1000000 --d 256 --c 10 --k 10 > gpurun_out/kernels_c4.log 2>&1; echo "kernels c4 exit $?"
timeout 600 python tools/bench_kernels.py --n 200000 --m 200000 --d 256 --c 50 --k 10 > gpurun_out/kernels_c50.log 2>&1; echo "kernels c50 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/kernels_n*.json')):
    for r in json.load(open(f)):
        print(f"{f[-22:]:24s} {r['kernel']:24s} {r['ms']:9.3f} ms {r['achieved_gbs']:8.1f} GB/s frac {r['frac']:.3f}")
PY
