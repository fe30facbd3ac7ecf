#!/bin/bash
# round 2, trip 13 (2 GPUs): the NCCL parity tests and the C4 line on 2 GPUs with the trip-12 kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q --timeout 500 > gpurun_out/r2_pytest13_dist.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest13_dist.log | tail -10
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-variants > gpurun_out/r2_b13_c4_2gpu.json 2> gpurun_out/r2_b13_c4_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b13_c4_2gpu.json')); r=d['roofline']; e=d.get('e2e') or {}
print('c4_2gpu', round(d['value']), 'ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],3), 'e2e', e and (round(e['value']), round(e['pinned']['value'])), 'parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch')}, 'clk', (d.get('clocks') or {}).get('sm_mhz'), [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:5]])
PY
tail -2 gpurun_out/r2_b13_c4_2gpu.err | cut -c1-300
