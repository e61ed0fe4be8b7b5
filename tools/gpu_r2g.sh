#!/bin/bash
mkdir -p gpurun_out
{
for w in 1 2 3 4 6 8; do echo "== hpair waves $w"; L3D_HPAIR_WAVES=$w python tools/skip_probe.py 28 112; done
for s in 2 3; do echo "== cost split $s, hpair waves 3"; L3D_COST_SPLIT=$s L3D_HPAIR_WAVES=3 python tools/skip_probe.py 28 112; done
echo "== lanes 35/140 waves 3"; L3D_HPAIR_WAVES=3 python tools/skip_probe.py 35 140
echo "== lanes 42/168 waves 3"; L3D_HPAIR_WAVES=3 python tools/skip_probe.py 42 168
echo "== lanes 21/84 waves 3"; L3D_HPAIR_WAVES=3 python tools/skip_probe.py 21 84
echo "== c4 waves 3"; L3D_HPAIR_WAVES=3 L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
echo "== c4 waves 6"; L3D_HPAIR_WAVES=6 L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
echo "== c4 waves 6 cost split 4"; L3D_COST_SPLIT=4 L3D_HPAIR_WAVES=6 L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
} > gpurun_out/r2g_pipe.log 2>&1
cat gpurun_out/r2g_pipe.log
bash tools/prof_r2.sh > gpurun_out/r2g_prof.log 2>&1; tail -14 gpurun_out/r2g_prof.log
