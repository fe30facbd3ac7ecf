#!/bin/bash
# round 2, trip 10 (4 GPUs): the other BASELINE shapes at N = 2 and N = 4 (C5 with NICDM and with
# DisSimLocal, C3, C2), each with parity_check
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run() { # name, n, args
  tr $2 bench.py --gpus $2 $3 > gpurun_out/r2_b10_$1.json 2> gpurun_out/r2_b10_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_b10_$1.json')); r=d['roofline']
    print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'share', round(r['all_search_launches_share_of_step'],3), 'parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first')}, r['screen'])
    print('   launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:5]])
except Exception as e:
    print('$1 failed', e)
PY
  tail -2 gpurun_out/r2_b10_$1.err | cut -c1-300; }
run c5_4gpu 4 "--workload c5 --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256"
run c5dsl_4gpu 4 "--workload c5dsl --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256"
run c5_2gpu 2 "--workload c5 --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256"
run c5dsl_2gpu 2 "--workload c5dsl --steps 2 --warmup 1 --no-variants --no-e2e --parity-rows 256"
run c3_4gpu 4 "--workload c3 --steps 5 --warmup 2 --no-variants --e2e-steps 3"
run c3_2gpu 2 "--workload c3 --steps 5 --warmup 2 --no-variants --e2e-steps 3"
run c2_4gpu 4 "--workload c2 --steps 20 --warmup 3 --no-variants --e2e-steps 3"
run c2_2gpu 2 "--workload c2 --steps 20 --warmup 3 --no-variants --e2e-steps 3"
