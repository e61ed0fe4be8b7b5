#!/bin/bash
# A/B of wavefront kernel switches (L3D_VW_FLAGS); tight timeouts: a wrong switch may hang the kernel
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_geometry or grouped_pipeline_cluster" 2>&1 | tail -2
for f in ${VW_FLAG_LIST:-0 8}; do
  echo "== L3D_VW_FLAGS=$f"
  L3D_VW_FLAGS=$f timeout 60 python tools/kernel_times.py 2 2>&1 | tail -1
  L3D_VW_FLAGS=$f timeout 60 python tools/skip_probe.py 28 112 2>&1 | tail -1
done
