#!/bin/bash
# round 2, trip 8 (1 GPU): closing evidence -- full GPU suite, smoke, bench lines of every
# BASELINE config, memory-bound kernel roofline, the reference's torch-eager rescale beside the
# fused kernels, ncu launch list + --set full summaries (reports are summarised on the box and
# deleted: gpurun_out must stay below 64 MiB), compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest8.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest8.log | tail -20
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/bench_kernels.py > gpurun_out/r2_kernels_c10.log 2>&1; cp gpurun_out/kernels_n1000000_c10_d256.json gpurun_out/r2_kernels_n1000000_c10_d256.json
grep -o '"kernel": "[^"]*", "ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_kernels_c10.log | paste - - | head -20
timeout 300 python tools/bench_kernels.py --n 200000 --m 200000 --c 50 > gpurun_out/r2_kernels_c50.log 2>&1; cp gpurun_out/kernels_n200000_c50_d256.json gpurun_out/r2_kernels_n200000_c50_d256.json
timeout 600 python tools/bench_rescale_vs_reference.py --out gpurun_out/r2_rescale_vs_reference_n1000000_c10.json > gpurun_out/r2_rescale_vs_ref_c10.log 2>&1; python - <<'PY'
import json
for line in open('gpurun_out/r2_rescale_vs_ref_c10.log'):
    if line.startswith('{'):
        r=json.loads(line); m=r['kiez_b200']
        print(r['method'], 'mine ms', round(m['ms'],3), 'kernel ms', round(m['kernel_ms'],3), 'launches', m['launches'], 'frac', round(m['frac_of_hbm_peak'],3), {k:(round(v.get('ms',0),2), round(v.get('kernel_ms',0),2), v.get('launches'), v.get('index_mismatch_rows'), v.get('failed')) for k,v in r.items() if k.startswith('reference')})
PY
timeout 600 python tools/bench_rescale_vs_reference.py --n 200000 --m 200000 --c 50 --out gpurun_out/r2_rescale_vs_reference_n200000_c50.json > gpurun_out/r2_rescale_vs_ref_c50.log 2>&1
b() { timeout ${3:-500} python bench.py $2 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$1.json')); r=d['roofline']; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']), round(d['e2e']['pinned']['value'])), 'parity', d['parity_check'] and d['parity_check']['mismatch'], r['screen'], d.get('data_variants') and {k:(round(v['value']), v['screen']) for k,v in d['data_variants'].items()}, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))"; tail -2 gpurun_out/r2_bench_$1.err; }
b c4_1gpu "--steps 10 --warmup 3"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c4_reference_arm.json 2> gpurun_out/r2_bench_c4_reference_arm.err; cut -c1-400 gpurun_out/r2_bench_c4_reference_arm.json
b c4_c50 "--steps 3 --warmup 2 --c 50 --no-cpu-baseline"
b c2 "--workload c2 --steps 20 --warmup 3 --no-cpu-baseline"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline"
b c5 "--workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --parity-rows 256" 700
b c5dsl "--workload c5dsl --steps 2 --warmup 1 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256" 700
# ncu: launch list of the default bench command, then --set full captures summarised here
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 > gpurun_out/r2_launches_c4.log 2>&1; echo "ncu launches exit $?"
python tools/launch_summary.py gpurun_out/r2_launches_c4.csv "python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0" > gpurun_out/r2_launches_c4_summary.txt; head -30 gpurun_out/r2_launches_c4_summary.txt
cap() { # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -o gpurun_out/r2_prof_$name "$@" > gpurun_out/r2_prof_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_prof_$name.ncu-rep > gpurun_out/r2_ncu_$name.txt 2> gpurun_out/r2_ncu_$name.err; echo "ncu $name: $(grep -E 'gpu__time_duration|dram__bytes_read.sum |sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed' gpurun_out/r2_ncu_$name.txt | tr -s ' ' | tr '\n' ';')"
  rm -f gpurun_out/r2_prof_$name.ncu-rep; }
cap knn_screen_dual knn_screen_kernel 5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores
cap knn_screen_c50 knn_screen_kernel 3 python bench.py --steps 1 --warmup 1 --c 50 --fused off --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores
for k in rows_small_kernel row_stats_small refine_topk dsl_fit dsl_transform k_occurrence prepare_rows; do cap $k $k 1 python tools/bench_kernels.py --iters 1; done
KB2_SAN_TIMEOUT=300 bash tools/sanitize.sh gpurun_out/r2_sanitizer
du -sh gpurun_out
