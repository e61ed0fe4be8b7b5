#!/bin/bash
TAG=${TAG:-r2b}
# round-2b ncu evidence (one B200, gpurun) for the wavefront aggregation build: the bench line first (no profiler), then
# the launch list of the same bench command, then --set full captures of the matcher kernels
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c3.log 2>&1 || { tail -5 gpurun_out/${TAG}_bench_c3.log; exit 1; }
grep '^{' gpurun_out/${TAG}_bench_c3.log | tail -1 | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_${TAG}_l.log 2>&1; tail -2 gpurun_out/ncu_${TAG}_l.log
python tools/pipe_one.py 14 14 > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sgbm_vwave" -c 2 -o gpurun_out/prof_${TAG}_vwave python tools/pipe_one.py 14 14 > gpurun_out/ncu_${TAG}_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sgbm_cost" -c 1 -o gpurun_out/prof_${TAG}_cost python tools/pipe_one.py 14 14 > gpurun_out/ncu_${TAG}_b.log 2>&1
ls -la gpurun_out/ | tail -8
