#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sgbm_vwave" -c 2 -o gpurun_out/prof_vwave python tools/pipe_one.py 14 14 > gpurun_out/ncu_vwave.log 2>&1; tail -3 gpurun_out/ncu_vwave.log
