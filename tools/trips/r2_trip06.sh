#!/bin/bash
# round 2, trip 6 (2 GPUs, short): per-rank minimum segment + threaded staging copy in the
# row-sharded pass: dual-direction parity tests, C4 at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q --timeout 500 -k "dual_direction or sharded_upload" > gpurun_out/r2_pytest6.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest6.log | tail -10
timeout 300 python -m pytest tests/test_gpu_kiez.py tests/test_gpu_analysis.py tests/test_gpu_upload.py -m gpu -q --timeout 500 -k "not full_size" > gpurun_out/r2_pytest6b.log 2>&1; echo "pytest-b exit $?"; tail -2 gpurun_out/r2_pytest6b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_b6_c4_2gpu.json 2> gpurun_out/r2_b6_c4_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b6_c4_2gpu.json')); r=d['roofline']
print('2gpu q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'share', round(r['all_search_launches_share_of_step'],3))
print('   e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','ms_per_step','pinned','fraction_of_device_value')})
print('   parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first')}, r['screen'], r['dual_direction'])
print('   launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]])
print('   variants', {k:(round(v['value']), v['screen'], v['parity_check'] and v['parity_check']['mismatch']) for k,v in (d.get('data_variants') or {}).items()})
PY
tail -2 gpurun_out/r2_b6_c4_2gpu.err
