#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_c3_n2.log 2>&1; grep '^{' gpurun_out/r2i_bench_c3_n2.log | tail -1 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/r2i_bench_ref_n2.log 2>&1; grep '^{' gpurun_out/r2i_bench_ref_n2.log | tail -1 | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c5 --steps 1 --warmup 3 > gpurun_out/r2i_bench_c5_n2.log 2>&1; grep '^{' gpurun_out/r2i_bench_c5_n2.log | tail -1 | cut -c1-300
tail -3 gpurun_out/r2i_bench_c5_n2.log | cut -c1-300
