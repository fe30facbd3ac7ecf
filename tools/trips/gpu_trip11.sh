#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -3 gpurun_out/pytest_all.log
timeout 600 python tools/bench_kernels.py --n 1000000 --m 1000000 --d 256 --c 10 --k 10 > gpurun_out/kernels_c4.log 2>&1; echo "kernels c4 exit $?"
timeout 600 python tools/bench_kernels.py --n 200000 --m 200000 --d 256 --c 50 --k 10 > gpurun_out/kernels_c50.log 2>&1; echo "kernels c50 exit $?"
timeout 600 python tools/bench_kernels.py --n 100000 --m 100000 --d 256 --c 100 --k 10 > gpurun_out/kernels_c100.log 2>&1; echo "kernels c100 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/kernels_n*.json')):
    for r in json.load(open(f)):
        print(f"{f[-22:]:24s} {r['kernel']:24s} {r['ms']:9.3f} ms {r['achieved_gbs']:8.1f} GB/s frac {r['frac']:.3f}")
PY
for m in "dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum"; do
KB2_TC_MODE=1 timeout 300 ncu --metrics $m --clock-control none -k regex:knn_tc -s 2 -c 1 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores 2>&1 | grep -E "dram__bytes|hit_rate|duration" | sed 's/^/single /'
timeout 300 ncu --metrics $m --clock-control none -k regex:knn_tc -s 2 -c 1 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores 2>&1 | grep -E "dram__bytes|hit_rate|duration" | sed 's/^/pair32 /'
KB2_TC2_BK=16 timeout 300 ncu --metrics $m --clock-control none -k regex:knn_tc -s 2 -c 1 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores 2>&1 | grep -E "dram__bytes|hit_rate|duration" | sed 's/^/pair16 /'
done
