#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/vw2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/vw2_tests.log
tail -4 gpurun_out/vw2_tests.log
VW_FLAG_LIST="0" bash tools/gpu_vwave.sh
[ -n "$VW_PROF" ] && bash tools/gpu_vwave_prof.sh
