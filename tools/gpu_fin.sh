#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped or c4_full" 2>&1 | tail -2
timeout 60 python tools/kernel_times.py 2 2>&1 | tail -1
timeout 60 python tools/skip_probe.py 28 112 2>&1 | tail -1
L3D_PROBE_CFG=c4 timeout 120 python tools/kernel_times.py 2 2>&1 | tail -1
L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py 21 42 2>&1 | tail -1
