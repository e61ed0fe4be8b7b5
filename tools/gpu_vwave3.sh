#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/vw2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/vw2_tests.log
tail -4 gpurun_out/vw2_tests.log
timeout 200 python tools/kernel_times.py 2 2>&1 | tail -2
timeout 200 python tools/skip_probe.py 28 112 2>&1 | tail -1
for s in sgbm_vwave sgbm_cost fgs_ wls_,fgs_ steger,gauss,hess remap,prefilter; do
  L3D_DEBUG_SKIP=$s timeout 200 python tools/skip_probe.py 28 112 2>&1 | tail -1
done
for l in 21 35 42; do timeout 200 python tools/skip_probe.py $l 168 2>&1 | tail -1; done
L3D_GROUP=6 timeout 200 python tools/skip_probe.py 24 96 2>&1 | tail -1
L3D_GROUP=5 timeout 200 python tools/skip_probe.py 30 120 2>&1 | tail -1
