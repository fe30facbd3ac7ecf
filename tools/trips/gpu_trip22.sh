#!/bin/bash
# full GPU test suite, default bench line (with e2e + cpu baseline), reference arm, A/B of the
# screen (two passes vs dual-direction), ncu launch list of the bench command, one ncu --set full
# capture of the screening kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -6 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/b_c4_default.json 2> gpurun_out/b_c4_default.err; echo "bench default exit $?"; cut -c1-600 gpurun_out/b_c4_default.json; tail -2 gpurun_out/b_c4_default.err
b() { # name, args
  timeout 900 python bench.py $2 --no-cpu-baseline --no-e2e > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms', round(r['avg_launch_ms'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], r['screen'], (r['dual_direction'] or {}).get('overflow_columns'))"; tail -2 gpurun_out/b_$1.err; }
C4="--steps 2 --warmup 2"
b c4_screen_fused "$C4 --fused on --precision screen"
b c4_screen_twopass "$C4 --fused off --precision screen"
b c4_c50_auto "$C4 --c 50"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/b_c4_reference.json 2> gpurun_out/b_ref.err; echo "reference arm exit $?"; cut -c1-400 gpurun_out/b_c4_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_c4.log 2>&1; echo "ncu launches exit $?"
prof() { # name, kernel regex, bench args
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_$1 -f python bench.py $3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"; tail -2 gpurun_out/ncu_$1.log | cut -c1-200; }
prof screen_rows_mid knn_screen "--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --fused off --precision screen"
prof screen_dual_mid knn_screen "--workload custom --n 131072 --m 262144 --d 256 --c 10 --k 10 --fused on --precision screen"
