#!/bin/bash
# N GPUs (N = $1): C4 bench over torchrun, default configuration
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_c4_${N}gpu.json 2> gpurun_out/b_c4_${N}gpu.err; echo "bench ${N}gpu exit $?"; python -c "
import json; d=json.load(open('gpurun_out/b_c4_${N}gpu.json')); r=d['roofline']; print('${N}gpu', 'q/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'top ms/step', round(r['avg_launch_ms']*r['launches']/d['steps'],1), 'frac', round(r['frac'],3), d['clocks'], d['e2e'], [round(s['avg_launch_ms'],1) for s in r['search_launches']])"; tail -2 gpurun_out/b_c4_${N}gpu.err | cut -c1-300
