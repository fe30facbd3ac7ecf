#!/bin/bash
# round 2, trip 7 (8 GPUs): C4 at N = 8 and N = 4 (row-sharded dual-direction pass; parity_check,
# both distributions, e2e), per-rank trace at N = 8, C5 (NICDM and DisSimLocal) at N = 8
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
show() { python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$1.json')); r=d['roofline']
    print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3), 'clk', (d['clocks'] or {}).get('sm_mhz'))
    print('   e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','ms_per_step','pinned','fraction_of_device_value')})
    print('   parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first')}, r['screen'], r['dual_direction'])
    print('   launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]])
    if d.get('data_variants'): print('   variants', {k:(round(v['value']), v['screen'], v['parity_check'] and v['parity_check']['mismatch']) for k,v in d['data_variants'].items()})
except Exception as e:
    print('$1 failed', e)
PY
}
tr 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_b7_c4_8gpu.json 2> gpurun_out/r2_b7_c4_8gpu.err; show r2_b7_c4_8gpu; tail -2 gpurun_out/r2_b7_c4_8gpu.err
tr 8 tools/trace_step.py > gpurun_out/r2_trace_c4_8gpu.txt 2>&1; grep -E "^rank" gpurun_out/r2_trace_c4_8gpu.txt
awk '/^rank 0/{f=1} f' gpurun_out/r2_trace_c4_8gpu.txt | grep -A24 "device time by activity" | head -26 | cut -c1-110
tr 4 bench.py --gpus 4 --steps 5 --warmup 3 --no-variants > gpurun_out/r2_b7_c4_4gpu.json 2> gpurun_out/r2_b7_c4_4gpu.err; show r2_b7_c4_4gpu
tr 8 bench.py --gpus 8 --workload c5 --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256 > gpurun_out/r2_b7_c5_8gpu.json 2> gpurun_out/r2_b7_c5_8gpu.err; show r2_b7_c5_8gpu; tail -2 gpurun_out/r2_b7_c5_8gpu.err
tr 8 bench.py --gpus 8 --workload c5dsl --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256 > gpurun_out/r2_b7_c5dsl_8gpu.json 2> gpurun_out/r2_b7_c5dsl_8gpu.err; show r2_b7_c5dsl_8gpu; tail -2 gpurun_out/r2_b7_c5dsl_8gpu.err
