#!/bin/bash
# round 2, trip 5 (2 GPUs): column-sharded threshold sample + overlapped slice upload + strict
# probe in the row-sharded pass: 2-GPU parity tests, C4 at N=2 (both distributions), and on one
# GPU the chunked threshold search (e2e) and the KB2_DUAL_DB A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q --timeout 600 -k "dual_direction or dsl or sharded_upload" > gpurun_out/r2_pytest5.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_pytest5.log | tail -10
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_knn.py -m gpu -q --timeout 600 -k "upload or probe or fused" > gpurun_out/r2_pytest5b.log 2>&1; echo "pytest-b exit $?"; tail -3 gpurun_out/r2_pytest5b.log
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
show() { python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$1.json')); r=d['roofline']
    print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'share', round(r['all_search_launches_share_of_step'],3))
    print('   e2e', d['e2e'] and {k: d['e2e'][k] for k in ('value','ms_per_step','pinned','fraction_of_device_value')})
    print('   parity', d['parity_check'] and {k: d['parity_check'][k] for k in ('rows','columns','mismatch','first')}, r['screen'])
    print('   launches', [(x['kind'], x['nq'], x['ny'], round(x['avg_launch_ms'],2), round(x['algorithmic_tflops'],1)) for x in r['search_launches'][:6]])
    if d.get('data_variants'): print('   variants', {k:(round(v['value']), v['screen'], v['parity_check'] and v['parity_check']['mismatch']) for k,v in d['data_variants'].items()})
except Exception as e:
    print('$1 failed', e)
PY
}
tr bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_b5_c4_2gpu.json 2> gpurun_out/r2_b5_c4_2gpu.err; show r2_b5_c4_2gpu; tail -2 gpurun_out/r2_b5_c4_2gpu.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2_b5_c4.json 2> gpurun_out/r2_b5_c4.err; show r2_b5_c4
KB2_DUAL_DB=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --no-e2e > gpurun_out/r2_b5_c4_db.json 2> gpurun_out/r2_b5_c4_db.err; show r2_b5_c4_db
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-variants --no-e2e > gpurun_out/r2_b5_c4_nodb.json 2> gpurun_out/r2_b5_c4_nodb.err; show r2_b5_c4_nodb
