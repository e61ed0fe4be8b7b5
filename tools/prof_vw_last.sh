#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/pipe_one.py 14 14 > gpurun_out/plain_vw.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sgbm_vwave" -c 2 -o gpurun_out/prof_r2c_vwave python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2c.log 2>&1; tail -1 gpurun_out/ncu_r2c.log
