#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -2 gpurun_out/bench_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_2gpu_ref.log 2>&1; echo "2gpu ref rc=$?"; tail -1 gpurun_out/bench_2gpu_ref.log
