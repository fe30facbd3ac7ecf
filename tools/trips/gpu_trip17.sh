#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_all.log 2>&1; echo "pytest(all) exit $?"; tail -5 gpurun_out/pytest_all.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_g2_fused.json 2> gpurun_out/bench_c4_g2_fused.err; echo "bench g2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_g2_fused.json')); r=d['roofline']; print('g2 fused', 'q/s', d['value'], 'ms/step', d['ms_per_step'], r['avg_launch_ms'], r['frac'], r['dual_direction'], d['e2e']['value'], d['clocks'])"; tail -3 gpurun_out/bench_c4_g2_fused.err
