#!/bin/bash
# 2 GPUs: NCCL parity tests + C4 bench over torchrun (dual-direction screen in row segments, column shards)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_dist.log 2>&1; echo "pytest(dist) exit $?"; tail -5 gpurun_out/pytest_dist.log
run2() { # name, args
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 $2 > gpurun_out/b_$1.json 2> gpurun_out/b_$1.err; echo "bench $1 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b_$1.json')); r=d['roofline']; print('$1', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks']['sm_mhz'], d['e2e'], (r['dual_direction'] or {}))"; tail -2 gpurun_out/b_$1.err | cut -c1-300; }
run2 c4_2gpu "--steps 3 --warmup 3 --no-cpu-baseline"
run2 c4_2gpu_twopass "--steps 2 --warmup 2 --no-cpu-baseline --no-e2e --fused off"
