#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
{
echo "== c3"; python tools/kernel_times.py 2
echo "== c4"; L3D_PROBE_CFG=c4 L3D_DEBUG_CLUSTERS=1 python tools/kernel_times.py 2 2>&1 | grep -v "max active" | sort | uniq -c | sort -rn | head -8
echo "== c1"; L3D_PROBE_CFG=c1 python tools/kernel_times.py 2
echo "== c4 pipeline 14 lanes 28 frames"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
echo "== c3 pipeline 28 lanes 112 frames"; python tools/skip_probe.py 28 112
echo "== c1 pipeline 28 lanes 448 frames HH"; L3D_PROBE_CFG=c1 python tools/skip_probe.py 28 448
echo "== latency"; python tools/latency_probe.py
} > gpurun_out/r2d_ktimes.log 2>&1
cat gpurun_out/r2d_ktimes.log
