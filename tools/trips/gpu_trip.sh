#!/bin/bash
# One GPU trip: diagnostics, then the gpu test-suite; everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== diag simt" > gpurun_out/diag.log
timeout 300 python tools/diag_knn.py simt >> gpurun_out/diag.log 2>&1; echo "exit $?" >> gpurun_out/diag.log
echo "== diag tc" >> gpurun_out/diag.log
timeout 300 python tools/diag_knn.py tc >> gpurun_out/diag.log 2>&1; echo "exit $?" >> gpurun_out/diag.log
tail -40 gpurun_out/diag.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -k "not tc" > gpurun_out/pytest_simt.log 2>&1; echo "pytest(simt) exit $?"
tail -15 gpurun_out/pytest_simt.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"
tail -30 gpurun_out/pytest_all.log
