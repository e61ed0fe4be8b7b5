#!/bin/bash
mkdir -p gpurun_out
L3D_PROBE_CFG=c4 timeout 300 python tools/pipe_one.py 14 14 > gpurun_out/plain_c4.log 2>&1 || exit 1
L3D_PROBE_CFG=c4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sgbm_vwave" -c 2 -o gpurun_out/prof_c4_vwave python tools/pipe_one.py 14 14 > gpurun_out/ncu_c4_vw.log 2>&1; tail -2 gpurun_out/ncu_c4_vw.log
