#!/bin/bash
# staged host copies: parity test of the three settings, then latency against the number of copy threads
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -x -q > gpurun_out/cp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/cp_tests.log; tail -3 gpurun_out/cp_tests.log
timeout 200 python tools/latency_copy_probe.py 0 1 2 4 8 2>&1 | tee gpurun_out/cp_latency.log | tail -6
timeout 200 python tools/latency_copy_probe.py 4 0 2>&1 | tee -a gpurun_out/cp_latency.log | tail -2
