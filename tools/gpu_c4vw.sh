#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged_geometry or grouped_pipeline_cluster or without_wls" > gpurun_out/c4vw_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c4vw_tests.log; tail -4 gpurun_out/c4vw_tests.log
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c4_full_pipeline" > gpurun_out/c4vw_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/c4vw_tests2.log; tail -3 gpurun_out/c4vw_tests2.log
L3D_PROBE_CFG=c4 timeout 120 python tools/kernel_times.py 2 2>&1 | tail -1
L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py 21 42 2>&1 | tail -1
L3D_VWAVE=0 L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py 21 42 2>&1 | tail -1
timeout 100 python tools/kernel_times.py 1 2>&1 | tail -1
