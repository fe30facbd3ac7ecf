#!/bin/bash
# round 2, trip 11 (1 GPU): rank-based list merge + packed FFMA2 key finish + cp.async rescale
# kernels + device-side centre / 8 memcpy threads in the upload.  Full GPU suite, A/B against the
# previous library build (build/ab/libkiez_b200_base.so via KB2_LIB) on the same box, bench lines of
# the BASELINE shapes, memory-bound kernels, ncu --set full of the largest dual segment, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest11.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest11.log | tail -20
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
b() { timeout ${3:-500} python bench.py $2 > gpurun_out/r2_b11_$1.json 2> gpurun_out/r2_b11_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b11_$1.json')); r=d['roofline']; e=d.get('e2e') or {}; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'e2e', e and (round(e['value']), round(e['pinned']['value']), e.get('upload_jobs_enqueued_ms')), 'parity', d['parity_check'] and d['parity_check']['mismatch'], r['screen'], 'clk', (d.get('clocks') or {}).get('sm_mhz'), [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:4]], d.get('data_variants') and {k:(round(v['value']), v['screen']) for k,v in d['data_variants'].items()})"; tail -2 gpurun_out/r2_b11_$1.err; }
# same box, alternating: previous build / new build (device value only)
KB2_LIB=$PWD/build/ab/libkiez_b200_base.so b c4_base "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c4_new "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
KB2_LIB=$PWD/build/ab/libkiez_b200_base.so b c4_base2 "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
KB2_FUSED_SAMPLE_DIV=64 b c4_div64 "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
KB2_SCREEN_RANGE_MB=48 b c4_range48 "--steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c4_1gpu "--steps 10 --warmup 3"
KB2_COPY_THREADS=4 b c4_e2e_4threads "--steps 2 --warmup 2 --no-cpu-baseline --no-variants --parity-rows 0"
b c4_c50 "--steps 3 --warmup 2 --c 50 --no-cpu-baseline --no-variants"
KB2_LIB=$PWD/build/ab/libkiez_b200_base.so b c3_base "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants --no-e2e"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants"
b c2 "--workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-variants"
b c5 "--workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256" 700
timeout 300 python tools/bench_kernels.py > gpurun_out/r2_kernels11_c10.log 2>&1; cp gpurun_out/kernels_n1000000_c10_d256.json gpurun_out/r2_kernels11_n1000000_c10_d256.json
grep -o '"kernel": "[^"]*", "ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_kernels11_c10.log | paste - - | head -20
timeout 300 python tools/bench_kernels.py --n 200000 --m 200000 --c 50 > gpurun_out/r2_kernels11_c50.log 2>&1; cp gpurun_out/kernels_n200000_c50_d256.json gpurun_out/r2_kernels11_n200000_c50_d256.json
grep -o '"kernel": "[^"]*", "ms": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2_kernels11_c50.log | paste - - | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen_kernel -s 7 -c 1 -o gpurun_out/r2_prof11_dual_large python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores > gpurun_out/r2_prof11_dual_large.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof11_dual_large.ncu-rep > gpurun_out/r2_ncu11_knn_screen_dual_large.txt 2> gpurun_out/r2_ncu11_dual_large.err; grep -E "gpu__time_duration|dram__bytes|sm__pipe_tensor_cycles_active|lts__t_sector_hit|sm__warps_active" gpurun_out/r2_ncu11_knn_screen_dual_large.txt
ls -la gpurun_out/r2_prof11_dual_large.ncu-rep
[ $(stat -c %s gpurun_out/r2_prof11_dual_large.ncu-rep 2>/dev/null || echo 99999999) -gt 30000000 ] && rm -f gpurun_out/r2_prof11_dual_large.ncu-rep
CS=/usr/local/cuda/bin/compute-sanitizer
for t in racecheck memcheck; do for fam in screen rescale; do
  timeout 300 $CS --tool $t --print-limit 20 --error-exitcode 77 python tools/sanitize_driver.py $fam > gpurun_out/r2_san11_${t}_${fam}.log 2>&1
  echo "$t $fam: rc=$?  $(grep -c '^ok ' gpurun_out/r2_san11_${t}_${fam}.log) checks ok | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_san11_${t}_${fam}.log | tail -1)"
done; done
