#!/bin/bash
# round 2, trip 15 (1 GPU, short): the rescale stage against the reference's torch-eager route with
# the final kernels, ncu --set full of the thread-per-row rescale kernel (cp.async staging)
mkdir -p gpurun_out
timeout 120 python tools/bench_rescale_vs_reference.py --out gpurun_out/r2_rescale15_vs_reference_n1000000_c10.json > gpurun_out/r2_rescale15_c10.log 2>&1
timeout 100 python tools/bench_rescale_vs_reference.py --n 200000 --m 200000 --c 50 --out gpurun_out/r2_rescale15_vs_reference_n200000_c50.json > gpurun_out/r2_rescale15_c50.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/r2_rescale15_vs_reference_n1000000_c10.json', 'gpurun_out/r2_rescale15_vs_reference_n200000_c50.json'):
    try:
        for r in json.load(open(f))['results']:
            m = r['kiez_b200']; ref = r['reference_torch_eager_f64']
            print(r['method'], 'mine', round(m['ms'], 3), m['launches'], 'ref', round(ref['ms'], 3), ref['launches'], 'x', round(ref['speedup_of_kiez_b200'], 2), 'mismatch rows', ref.get('index_mismatch_rows'))
    except Exception as e:
        print(f, 'failed', e)
PY
timeout 120 ncu --set full --clock-control none --import-source on -k regex:rows_small_kernel -c 1 -o gpurun_out/r2_prof15_rows_small python tools/bench_kernels.py --iters 1 > gpurun_out/r2_prof15_rows_small.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_prof15_rows_small.ncu-rep > gpurun_out/r2_ncu15_rows_small_kernel.txt 2>&1; grep -E "kernel:|gpu__time_duration|dram__bytes|gpu__dram_throughput|sm__warps_active|registers" gpurun_out/r2_ncu15_rows_small_kernel.txt
rm -f gpurun_out/r2_prof15_rows_small.ncu-rep
