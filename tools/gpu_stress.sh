#!/bin/bash
timeout 300 python tools/stress_determinism.py 40
L3D_PROBE_CFG=c4 timeout 300 python tools/stress_determinism.py 12
L3D_VWAVE=1 L3D_PROBE_GEOM=437,200,64,5 timeout 200 python tools/stress_determinism.py 60
L3D_PROBE_GEOM=640,480,128,7 timeout 200 python tools/stress_determinism.py 60
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_geometry or without_wls" 2>&1 | tail -1; done
