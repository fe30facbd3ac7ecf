#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_io.py -m gpu -q --timeout 150 2>&1 | tail -4
