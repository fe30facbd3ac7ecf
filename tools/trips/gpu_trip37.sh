#!/bin/bash
# pinned D2H staging of numpy results: kiez-level parity tests + default bench line (e2e)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kiez.py tests/test_gpu_integration_stub.py tests/test_gpu_analysis.py -m gpu -q --timeout 600 > gpurun_out/pytest_kiez.log 2>&1; echo "pytest(kiez) exit $?"; tail -3 gpurun_out/pytest_kiez.log
timeout 900 python bench.py > gpurun_out/b_c4_default2.json 2> gpurun_out/b_c4_default2.err; echo "bench default exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b_c4_default2.json')); r=d['roofline']; print('default', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks'], d['e2e'])"; tail -2 gpurun_out/b_c4_default2.err
