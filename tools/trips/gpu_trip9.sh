#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 120 python tools/diag_knn.py tc > gpurun_out/diag_tc2.log 2>&1; echo "diag tc(pair) exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror|timed out|row " gpurun_out/diag_tc2.log | head -20
KB2_TC2_BK=16 timeout 120 python tools/diag_knn.py tc > gpurun_out/diag_tc2_bk16.log 2>&1; echo "diag tc(pair,bk16) exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror|timed out" gpurun_out/diag_tc2_bk16.log | head
timeout 120 python tools/diag_knn.py tc1 simt > gpurun_out/diag.log 2>&1; echo "diag tc1/simt exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror" gpurun_out/diag.log | head
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
run() { AB_NAME=$1 timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log; }
for rep in 1 2; do
KB2_TC_MODE=1 KB2_TC_CONFIG=256x32 run single_256x32
KB2_TC2_BK=32 run pair_bk32
KB2_TC2_BK=16 run pair_bk16
done
AB_CAP=56 KB2_TC_MODE=1 run single_cap56
AB_CAP=56 run pair_cap56
AB_CAP=112 KB2_TC_MODE=1 run single_cap112
AB_CAP=112 run pair_cap112
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc2 -s 2 -c 1 -o gpurun_out/prof_knn_tc2 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
