#!/bin/bash
# tests of the matcher paths + kernel times + pipeline rate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${QK:-grouped or pipeline}" > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
tail -3 gpurun_out/q_tests.log
timeout 200 python tools/kernel_times.py 2 2>&1 | tail -2
timeout 200 python tools/skip_probe.py 28 112 2>&1 | tail -1
