#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
tools/ubench/sgmstep > gpurun_out/r2c_sgmstep.log 2>&1; cat gpurun_out/r2c_sgmstep.log
{
echo "== c3"; python tools/kernel_times.py 2
echo "== c4"; L3D_PROBE_CFG=c4 python tools/kernel_times.py 2
echo "== c4 pipeline 14 lanes 28 frames"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
echo "== c4 pipeline 28 lanes 56 frames"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 28 56
echo "== c4 pipeline no vgroup"; L3D_NO_VGROUP=1 L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
} > gpurun_out/r2c_ktimes.log 2>&1
cat gpurun_out/r2c_ktimes.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_bench.log 2>&1; tail -1 gpurun_out/r2c_bench.log | cut -c1-200
