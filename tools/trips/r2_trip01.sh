#!/bin/bash
# round 2, trip 1: full GPU suite with the tighter proof bound / list boost / new evidence tests,
# smoke, the default bench line (parity_check, data_variants, pageable + pinned e2e),
# compute-sanitizer over every kernel family (small shapes)
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q --timeout 600 -x -s > gpurun_out/r2_pytest1.log 2>&1; echo "pytest exit $?"
grep -E "max \|screen|dual-direction:|probe fractions|c4 full size|passed|failed|error" gpurun_out/r2_pytest1.log | tail -40
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4_t1.json 2> gpurun_out/r2_bench_c4_t1.err; echo "bench exit $?"
tail -3 gpurun_out/r2_bench_c4_t1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_t1.json')); r=d['roofline']
print('q/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'frac', round(r['frac'],3), 'e2e', d['e2e'], 'parity', d['parity_check'])
print('variants', json.dumps(d['data_variants'])[:1500])
print('cpu', d['cpu_baseline']); print('screen', r['screen'], r['dual_direction'])
PY
KB2_SAN_TIMEOUT=300 bash tools/sanitize.sh gpurun_out/sanitizer
