#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -6 gpurun_out/r2f_tests.log
{
for m in 0 1; do for w in 1 2 3 4 6; do echo "== hpair mode $m waves $w"; L3D_HPAIR_MODE=$m L3D_HPAIR_WAVES=$w python tools/kernel_times.py 1; done; done
echo "== pipeline mode 0"; L3D_HPAIR_MODE=0 python tools/skip_probe.py 28 112
echo "== pipeline mode 1"; L3D_HPAIR_MODE=1 python tools/skip_probe.py 28 112
echo "== pipeline mode 1 waves 3"; L3D_HPAIR_MODE=1 L3D_HPAIR_WAVES=3 python tools/skip_probe.py 28 112
} > gpurun_out/r2f_ktimes.log 2>&1
cat gpurun_out/r2f_ktimes.log
