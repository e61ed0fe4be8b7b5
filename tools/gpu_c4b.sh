#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4_full or (ragged_geometry and 256)" 2>&1 | tail -1
L3D_PROBE_CFG=c4 timeout 120 python tools/kernel_times.py 2 2>&1 | tail -1
L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py 21 42 2>&1 | tail -1
