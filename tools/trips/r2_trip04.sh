#!/bin/bash
# round 2, trip 4 (1 GPU): full suite with the thread-per-row rescale kernels and the strict
# probe; memory-bound kernel roofline + the reference's torch-eager rescale beside it; ncu launch
# list of the default bench command and --set full captures (dual screen kernel, rescale family);
# C3 with / without hub scores; C5 DisSimLocal
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest4.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest4.log | tail -20
timeout 300 python tools/bench_kernels.py > gpurun_out/r2_kernels_c10.log 2>&1; grep -o '"kernel": "[^"]*", "ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_kernels_c10.log | paste - - | head -20
timeout 300 python tools/bench_kernels.py --n 200000 --m 200000 --c 50 > gpurun_out/r2_kernels_c50.log 2>&1
timeout 600 python tools/bench_rescale_vs_reference.py > gpurun_out/r2_rescale_vs_ref_c10.log 2>&1; python - <<'PY'
import json
for line in open('gpurun_out/r2_rescale_vs_ref_c10.log'):
    if line.startswith('{'):
        r=json.loads(line); m=r['kiez_b200']
        print(r['method'], 'mine ms', round(m['ms'],3), 'launches', m['launches'], 'frac', round(m['frac_of_hbm_peak'],3), {k:(round(v.get('ms',0),2), v.get('launches'), v.get('index_mismatch_rows'), v.get('failed')) for k,v in r.items() if k.startswith('reference')})
PY
timeout 600 python tools/bench_rescale_vs_reference.py --n 200000 --m 200000 --c 50 > gpurun_out/r2_rescale_vs_ref_c50.log 2>&1
b() { timeout ${3:-400} python bench.py $2 > gpurun_out/r2_b4_$1.json 2> gpurun_out/r2_b4_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b4_$1.json')); r=d['roofline']; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['pinned']['value'])), 'parity', d['parity_check'] and d['parity_check']['mismatch'], r['screen'], d.get('data_variants') and {k:(round(v['value']), v['screen']) for k,v in d['data_variants'].items()})"; tail -2 gpurun_out/r2_b4_$1.err; }
b c4 "--steps 5 --warmup 3 --no-cpu-baseline"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c3_nohub "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants --no-e2e --no-hub-scores"
b c5dsl "--workload c5dsl --steps 1 --warmup 1 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256" 600
# ncu: launch list of the default bench command (2 steps), then --set full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 > gpurun_out/r2_launches_c4.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen_kernel -s 5 -c 1 -o gpurun_out/r2_prof_screen_dual python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores > gpurun_out/r2_prof_screen_dual.log 2>&1; echo "ncu dual exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rows_small|row_stats_small|refine_topk|dsl_|k_occurrence|topk|prepare_rows" -c 40 -o gpurun_out/r2_prof_membound python tools/bench_kernels.py --iters 1 > gpurun_out/r2_prof_membound.log 2>&1; echo "ncu membound exit $?"
ls -la gpurun_out/*.ncu-rep
