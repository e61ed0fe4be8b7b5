#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -6 gpurun_out/r2h_tests.log
{
echo "== c3"; python tools/kernel_times.py 2
echo "== c4"; L3D_PROBE_CFG=c4 python tools/kernel_times.py 1
echo "== c3 pipeline"; python tools/skip_probe.py 28 112
echo "== c4 pipeline"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
} > gpurun_out/r2h_ktimes.log 2>&1
cat gpurun_out/r2h_ktimes.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2h_bench_c3.log 2>&1; tail -1 gpurun_out/r2h_bench_c3.log | cut -c1-250
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_ref.log 2>&1; tail -1 gpurun_out/r2h_bench_ref.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -2 gpurun_out/r2h_smoke.log
