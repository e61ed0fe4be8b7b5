#!/bin/bash
# bench lines of the configurations the final check does not cover (c2, c4, c5) and the reference arm
mkdir -p gpurun_out
for c in c2 c4; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/fin_bench_$c.log 2>&1; grep '^{' gpurun_out/fin_bench_$c.log | tail -1 | cut -c1-120; done
timeout 400 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/fin_bench_c5.log 2>&1; grep '^{' gpurun_out/fin_bench_c5.log | tail -1 | cut -c1-120
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/fin_bench_ref.log 2>&1; grep '^{' gpurun_out/fin_bench_ref.log | tail -1 | cut -c1-160
