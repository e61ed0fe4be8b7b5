#!/bin/bash
# 9-CTA clusters x 8 columns per warp against 8 x 9 at config 3: kernel times alone, pipeline rate, parity of the grouped pipeline
mkdir -p gpurun_out
for cl in 0 9 ${MORE_CL}; do
  echo "== L3D_VWAVE_CLUSTER=$cl"
  L3D_VWAVE_CLUSTER=$cl timeout 120 python tools/kernel_times.py 2 2>&1 | tail -1
  L3D_VWAVE_CLUSTER=$cl timeout 120 python tools/skip_probe.py 28 112 2>&1 | tail -1
done
L3D_VWAVE_CLUSTER=9 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or pipeline" 2>&1 | tail -2
