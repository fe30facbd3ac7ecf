#!/bin/bash
# round 2, trip 9 (1 GPU): index-length-aware append buffers (C2 / C3), ncu --set full of the
# LARGEST row segment of the dual-direction screen (the capture of trip 8 took the first segment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_kiez.py -m gpu -q --timeout 600 -k "not full_size_c4" > gpurun_out/r2_pytest9.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest9.log
b() { timeout ${3:-400} python bench.py $2 > gpurun_out/r2_b9_$1.json 2> gpurun_out/r2_b9_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b9_$1.json')); r=d['roofline']; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'parity', d['parity_check'] and d['parity_check']['mismatch'], r['screen'], [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:4]])"; tail -2 gpurun_out/r2_b9_$1.err; }
b c2 "--workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-variants --no-e2e"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c4 "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c4_c50 "--steps 2 --warmup 1 --c 50 --no-cpu-baseline --no-variants --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen_kernel -s 7 -c 1 -o gpurun_out/r2_prof_dual_large python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores > gpurun_out/r2_prof_dual_large.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof_dual_large.ncu-rep > gpurun_out/r2_ncu_knn_screen_dual_large.txt 2> gpurun_out/r2_ncu_dual_large.err; grep -E "gpu__time_duration|dram__bytes|sm__pipe_tensor_cycles_active|lts__t_sector_hit|sm__warps_active" gpurun_out/r2_ncu_knn_screen_dual_large.txt
rm -f gpurun_out/r2_prof_dual_large.ncu-rep
timeout 200 python tools/bench_kernels.py > gpurun_out/r2_kernels9.log 2>&1; grep -o '"kernel": "[^"]*", "ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_kernels9.log | paste - - | head -12
