set -x
python tools/pipe_one.py 14 14 > gpurun_out/plain_r1g.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1g.csv python tools/pipe_one.py 14 14 > gpurun_out/ncu_r1g_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sgbm_cost|sgbm_scan_hpair" -c 2 -o gpurun_out/prof_r1g_front python tools/pipe_one.py 14 14 > gpurun_out/ncu_r1g_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sgbm_vgroup" -c 2 -o gpurun_out/prof_r1g_vgroup python tools/pipe_one.py 14 14 > gpurun_out/ncu_r1g_b.log 2>&1
ncu --set full --clock-control none -k regex:"fgs_" -c 2 -o gpurun_out/prof_r1g_fgs python tools/pipe_one.py 14 14 > gpurun_out/ncu_r1g_c.log 2>&1
ls -la gpurun_out/
