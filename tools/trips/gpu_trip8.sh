#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ab.log
timeout 300 python tools/diag_knn.py simt tc > gpurun_out/diag.log 2>&1; echo "diag exit $?"; grep -E "DIAG|bad_rows=[1-9]|rror" gpurun_out/diag.log | head
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -4 gpurun_out/pytest_all.log
V=build/variants
run() { AB_NAME=$1 KB2_LIB=$2 KB2_TC_CONFIG=$3 timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log; }
for rep in 1 2; do
for cfg in 256x32 256x16; do run nobound $PWD/$V/nobound/lib/libkiez_b200.so $cfg; run bound $PWD/kiez_b200/lib/libkiez_b200.so $cfg; done
done
AB_CAP=56 run bound_cap56 $PWD/kiez_b200/lib/libkiez_b200.so 128x32
AB_CAP=56 run bound_cap56 $PWD/kiez_b200/lib/libkiez_b200.so 128x16
AB_CAP=112 run bound_cap112 $PWD/kiez_b200/lib/libkiez_b200.so 128x16
KB2_TC_CONFIG=256x32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc -s 2 -c 1 -o gpurun_out/prof_knn_tc_v3c_256x32 python bench.py --workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_full.log 2>&1; echo "ncu exit $?"
