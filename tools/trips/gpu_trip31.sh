#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/trace_step.py > gpurun_out/trace_c4.txt 2> gpurun_out/trace_c4.err; echo "trace exit $?"; cat gpurun_out/trace_c4.txt; tail -3 gpurun_out/trace_c4.err
