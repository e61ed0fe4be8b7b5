#!/bin/bash
# round-2 ncu evidence (one B200, gpurun): launch list of a pipeline run + --set full captures of the hot kernels
set -x
python tools/pipe_one.py 14 14 > gpurun_out/plain_r2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2.csv python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sgbm_cost|sgbm_scan_hpair" -c 2 -o gpurun_out/prof_r2_front python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sgbm_vgroup" -c 2 -o gpurun_out/prof_r2_vgroup python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2_b.log 2>&1
ncu --set full --clock-control none -k regex:"fgs_" -c 2 -o gpurun_out/prof_r2_fgs python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2_c.log 2>&1
# config 4: the 8-warp cluster kernel and the 32-column shared pixel-cost pass
L3D_PROBE_CFG=c4 ncu --set full --clock-control none -k regex:"sgbm_cost|sgbm_scan_hpair|sgbm_vgroup" -c 4 -o gpurun_out/prof_r2_c4 python tools/pipe_one.py 14 14 > gpurun_out/ncu_r2_d.log 2>&1
ls -la gpurun_out/ | tail -12
