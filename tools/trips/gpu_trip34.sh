#!/bin/bash
# final single-GPU evidence: full GPU suite, default bench line, reference arm, launch list, ncu --set full of a big row segment
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -3 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/b_c4_default.json 2> gpurun_out/b_c4_default.err; echo "bench default exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b_c4_default.json')); r=d['roofline']; print('default', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks'], d['e2e'], d['cpu_baseline'], [round(s['avg_launch_ms'],1) for s in r['search_launches']], r['dual_direction'], r['screen'], d['gpu_launches'])"; tail -2 gpurun_out/b_c4_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/b_c4_reference.json 2> gpurun_out/b_ref.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/b_c4_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_c4.log 2>&1; echo "ncu launches exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 3 -c 1 -o gpurun_out/prof_screen_dual_final -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_final.log 2>&1; echo "ncu full exit $?"; tail -2 gpurun_out/ncu_final.log | cut -c1-200
