#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n_bench_c3_n$N.log 2>&1; grep '^{' gpurun_out/r2n_bench_c3_n$N.log | tail -1 | cut -c1-300
