#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_c3_n2.log 2>&1; grep '^{' gpurun_out/r2k_bench_c3_n2.log | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --steps 1 --warmup 3 > gpurun_out/r2k_bench_c5_n2.log 2>&1; grep '^{' gpurun_out/r2k_bench_c5_n2.log | tail -1 | cut -c1-300
