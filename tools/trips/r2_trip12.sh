#!/bin/bash
# round 2, trip 12 (1 GPU): ballot-based rank merge, batched list seed, STS appends.  The full GPU
# suite gates everything else; then A/B against the trip-10 library on the same box, the L2 range
# size and the rank / bitonic crossover, the closing bench lines of the BASELINE shapes, the
# reference arm, memory-bound kernels, ncu launch list + --set full of the largest dual segment,
# sanitizer over the screen family
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest12.log 2>&1; rc=$?; echo "pytest exit $rc"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest12.log | tail -20
if [ $rc -ne 0 ]; then grep -B5 -A40 "Error\|assert" gpurun_out/r2_pytest12.log | head -150; exit 1; fi
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
b() { timeout ${3:-500} python bench.py $2 > gpurun_out/r2_b12_$1.json 2> gpurun_out/r2_b12_$1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b12_$1.json')); r=d['roofline']; e=d.get('e2e') or {}; print('$1', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'e2e', e and (round(e['value']), round(e['pinned']['value']), e.get('upload_jobs_enqueued_ms')), 'parity', d['parity_check'] and d['parity_check']['mismatch'], r['screen'], 'clk', (d.get('clocks') or {}).get('sm_mhz'), [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:4]], d.get('data_variants') and {k:(round(v['value']), v['screen']) for k,v in d['data_variants'].items()}, 'cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value']))"; tail -2 gpurun_out/r2_b12_$1.err; }
S="--no-cpu-baseline --no-variants --no-e2e"
KB2_LIB=$PWD/build/ab/libkiez_b200_base.so b c4_base "--steps 3 --warmup 2 $S"
b c4_new "--steps 3 --warmup 2 $S"
KB2_SCREEN_RANGE_MB=48 b c4_range48 "--steps 3 --warmup 2 $S"
b c3 "--workload c3 --steps 5 --warmup 2 --no-cpu-baseline --no-variants"
KB2_RANK_MAX_CNT=64 b c3_rank64 "--workload c3 --steps 5 --warmup 2 $S"
b c2 "--workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-variants"
KB2_RANK_MAX_CNT=64 b c2_rank64 "--workload c2 --steps 20 --warmup 3 $S"
b c4_1gpu "--steps 10 --warmup 3"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_b12_c4_reference_arm.json 2> gpurun_out/r2_b12_c4_reference_arm.err; cut -c1-300 gpurun_out/r2_b12_c4_reference_arm.json
b c4_c50 "--steps 3 --warmup 2 --c 50 --no-cpu-baseline --no-variants"
b c5 "--workload c5 --steps 2 --warmup 1 $S --parity-rows 256" 700
timeout 300 python tools/bench_kernels.py > gpurun_out/r2_kernels12_c10.log 2>&1; cp gpurun_out/kernels_n1000000_c10_d256.json gpurun_out/r2_kernels12_n1000000_c10_d256.json
timeout 300 python tools/bench_kernels.py --n 200000 --m 200000 --c 50 > gpurun_out/r2_kernels12_c50.log 2>&1; cp gpurun_out/kernels_n200000_c50_d256.json gpurun_out/r2_kernels12_n200000_c50_d256.json
# ncu: launch list of the default bench command (shares), then --set full of the largest dual segment
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches12_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 > gpurun_out/r2_launches12.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches12_c4.csv > gpurun_out/r2_launches12_c4_summary.txt 2>&1; head -12 gpurun_out/r2_launches12_c4_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen_kernel -s 7 -c 1 -o gpurun_out/r2_prof12_dual_large python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --parity-rows 0 --no-hub-scores > gpurun_out/r2_prof12_dual_large.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof12_dual_large.ncu-rep > gpurun_out/r2_ncu12_knn_screen_dual_large.txt 2> gpurun_out/r2_ncu12_dual_large.err; grep -E "gpu__time_duration|dram__bytes|sm__pipe_tensor_cycles_active|lts__t_sector_hit|sm__warps_active" gpurun_out/r2_ncu12_knn_screen_dual_large.txt
[ $(stat -c %s gpurun_out/r2_prof12_dual_large.ncu-rep 2>/dev/null || echo 99999999) -gt 30000000 ] && rm -f gpurun_out/r2_prof12_dual_large.ncu-rep
CS=/usr/local/cuda/bin/compute-sanitizer
for t in racecheck memcheck; do for fam in screen; do
  timeout 300 $CS --tool $t --print-limit 20 --error-exitcode 77 python tools/sanitize_driver.py $fam > gpurun_out/r2_san12_${t}_${fam}.log 2>&1
  echo "$t $fam: rc=$?  $(grep -c '^ok ' gpurun_out/r2_san12_${t}_${fam}.log) checks ok | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_san12_${t}_${fam}.log | tail -1)"
done; done
