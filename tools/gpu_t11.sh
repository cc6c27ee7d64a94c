#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1; echo "8gpu rc=$?"; tail -1 gpurun_out/bench_8gpu.log | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_4gpu.log 2>&1; echo "4gpu rc=$?"; tail -1 gpurun_out/bench_4gpu.log | cut -c1-300
