#!/bin/bash
# wavefront kernel experiments: every step under its own timeout (a hung kernel must not hold the box)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped_pipeline_cluster_aggregation or ragged_geometry" > gpurun_out/vw_tests1.log 2>&1; echo "rc=$?" >> gpurun_out/vw_tests1.log
tail -5 gpurun_out/vw_tests1.log
for f in ${VW_FLAG_LIST:-0 1}; do
  echo "== L3D_VW_FLAGS=$f"
  L3D_VW_FLAGS=$f timeout 200 python tools/kernel_times.py 2 2>&1 | tail -2
  L3D_VW_FLAGS=$f timeout 200 python tools/skip_probe.py 28 112 2>&1 | tail -1
done
echo "== no vwave"
L3D_NO_VWAVE=1 timeout 200 python tools/skip_probe.py 28 112 2>&1 | tail -1
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
