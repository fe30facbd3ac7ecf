#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_g$n.json 2> gpurun_out/bench_c4_g$n.err; echo "bench g$n exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c4_g$n.json')); print('g$n', 'q/s', d['value'], 'ms/step', d['ms_per_step'], 'ms/launch', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/bench_c4_g$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_g8.json 2> gpurun_out/bench_ref_g8.err; echo "ref g8 exit $?"; cut -c1-300 gpurun_out/bench_ref_g8.json
