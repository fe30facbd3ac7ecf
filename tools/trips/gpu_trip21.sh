#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -8 gpurun_out/pytest_all.log
prof() { # name, kernel regex, bench args
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_$1 -f python bench.py $3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"; tail -2 gpurun_out/ncu_$1.log | cut -c1-200; }
prof screen_dual_c4like knn_screen "--workload custom --n 1000000 --m 65536 --d 256 --c 10 --k 10 --fused on --precision screen"
prof screen_dual_mid knn_screen "--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --fused on --precision screen"
prof screen_rows_mid knn_screen "--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --fused off --precision screen"
