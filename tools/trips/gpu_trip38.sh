#!/bin/bash
# the full-size C4 sampled-parity test (both directions) + kiez-level tests after the D2H revert
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kiez.py -m gpu -q --timeout 800 -k "full_size or api_behaviour or golden" --durations=3 > gpurun_out/pytest_full.log 2>&1; echo "pytest(full size) exit $?"; tail -5 gpurun_out/pytest_full.log; grep -n "slowest\|passed\|failed" gpurun_out/pytest_full.log | tail -3
