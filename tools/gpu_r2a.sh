#!/bin/bash
# round-2 first GPU call: regression tests, SGM-step microbenchmark, kernel-group times with the cost-kernel experiment bits
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -3 gpurun_out/r2a_tests.log
tools/ubench/sgmstep > gpurun_out/r2a_sgmstep.log 2>&1; cat gpurun_out/r2a_sgmstep.log
for d in 0 1 2 4 6; do echo "L3D_COST_DBG=$d"; L3D_COST_NO_DUAL=1 L3D_COST_DBG=$d python tools/kernel_times.py 2; done > gpurun_out/r2a_ktimes.log 2>&1
cat gpurun_out/r2a_ktimes.log
python tools/vg_time.py > gpurun_out/r2a_vgtime.log 2>&1; cat gpurun_out/r2a_vgtime.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log | cut -c1-600
