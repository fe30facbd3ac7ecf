#!/bin/bash
# round 2, trip 3 (2 GPUs): row-sharded dual-direction pass -- 2-GPU parity tests, C4 bench at
# N=2 (rows vs cols mode), per-rank trace; plus single-GPU traces of C3 and C4-c50 (where the
# non-search time goes) and C5 with the dual-direction pass forced
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q --timeout 600 -s > gpurun_out/r2_pytest3.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR|rank [01] " gpurun_out/r2_pytest3.log | tail -30
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
b() { # name, args
  tr bench.py --gpus 2 $2 > gpurun_out/r2_b2_$1.json 2> gpurun_out/r2_b2_$1.err; echo "bench $1 exit $?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_b2_$1.json')); r=d['roofline']
    print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'search share', round(r['all_search_launches_share_of_step'],3))
    print('   e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','ms_per_step','pinned','fraction_of_device_value','d2h_bytes_per_step')})
    print('   parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first')})
    print('   launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]], r['dual_direction'])
    if d.get('data_variants'): print('   variants', {k:(round(v['value']), v['screen'], v['parity_check'] and v['parity_check']['mismatch']) for k,v in d['data_variants'].items()})
except Exception as e:
    print('$1 failed', e)
PY
  tail -3 gpurun_out/r2_b2_$1.err; }
b c4_rows "--steps 5 --warmup 3"
b c4_cols "--steps 3 --warmup 2 --shard-mode cols --no-variants --no-e2e"
tr tools/trace_step.py > gpurun_out/r2_trace_c4_2gpu.txt 2>&1; grep -A3 "^rank" gpurun_out/r2_trace_c4_2gpu.txt | head -20
timeout 300 python tools/trace_step.py --n 100000 --m 100000 --c 100 --hubness MutualProximity --method normal > gpurun_out/r2_trace_c3.txt 2>&1; head -60 gpurun_out/r2_trace_c3.txt
timeout 300 python tools/trace_step.py --c 50 > gpurun_out/r2_trace_c4_c50.txt 2>&1; grep -A30 "device time by" gpurun_out/r2_trace_c4_c50.txt
timeout 600 python bench.py --workload c5 --fused on --steps 1 --warmup 1 --no-cpu-baseline --no-variants --no-e2e --parity-rows 256 > gpurun_out/r2_b_c5_fused.json 2> gpurun_out/r2_b_c5_fused.err; python -c "
import json; d=json.load(open('gpurun_out/r2_b_c5_fused.json')); r=d['roofline']; print('c5 fused', round(d['value']), round(d['ms_per_step'],1), d['parity_check']['mismatch'], [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]])"; tail -3 gpurun_out/r2_b_c5_fused.err
