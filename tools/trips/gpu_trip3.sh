#!/bin/bash
# 2-GPU trip: NCCL path tests + a 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist exit $?"; tail -15 gpurun_out/pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c2 --steps 3 --warmup 3 > gpurun_out/bench_c2_g2.json 2> gpurun_out/bench_c2_g2.err; echo "bench g2 exit $?"; cat gpurun_out/bench_c2_g2.json; tail -5 gpurun_out/bench_c2_g2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_g2.json 2> gpurun_out/bench_c4_g2.err; echo "bench c4 g2 exit $?"; cat gpurun_out/bench_c4_g2.json; tail -5 gpurun_out/bench_c4_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --workload c2 --steps 1 --warmup 1 > gpurun_out/bench_ref_g2.json 2>gpurun_out/bench_ref_g2.err; echo "ref exit $?"; cat gpurun_out/bench_ref_g2.json
