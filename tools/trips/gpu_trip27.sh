#!/bin/bash
# ncu --set full of one mid-size row segment of the dual-direction screen at true C4 shapes
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 3 -c 1 -o gpurun_out/prof_screen_dual_c4seg -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-hub-scores > gpurun_out/ncu_c4seg.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_c4seg.log | cut -c1-200
