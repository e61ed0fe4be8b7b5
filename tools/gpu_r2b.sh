#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_shared or sgbm_small or c3_full or c1_param" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
{
echo "== dual"; python tools/kernel_times.py 2
echo "== dual dbg=1 (no own copy-out)"; L3D_COST_DBG=1 python tools/kernel_times.py 1
echo "== dual dbg=8 (no sheared emission)"; L3D_COST_DBG=8 python tools/kernel_times.py 1
echo "== dual dbg=9"; L3D_COST_DBG=9 python tools/kernel_times.py 1
echo "== two staged singles"; L3D_COST_NO_DUAL=1 python tools/kernel_times.py 2
echo "== two round-1 singles"; L3D_COST_NO_DUAL=1 L3D_COST_NO_STAGED=1 python tools/kernel_times.py 2
} > gpurun_out/r2b_ktimes.log 2>&1
cat gpurun_out/r2b_ktimes.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench.log 2>&1; tail -1 gpurun_out/r2b_bench.log | cut -c1-200
