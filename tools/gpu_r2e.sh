#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -8 gpurun_out/r2e_tests.log
{
echo "== c3 parallel FGS"; python tools/kernel_times.py 2
echo "== c3 serial FGS"; L3D_FGS_SERIAL=1 python tools/kernel_times.py 1
echo "== c3 pipeline parallel"; python tools/skip_probe.py 28 112
echo "== c3 pipeline serial"; L3D_FGS_SERIAL=1 python tools/skip_probe.py 28 112
echo "== c4 pipeline 14/28"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
echo "== c4 pipeline 28/56"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 28 56
echo "== latency"; python tools/latency_probe.py 2>&1 | grep -v "初始化"
} > gpurun_out/r2e_ktimes.log 2>&1
cat gpurun_out/r2e_ktimes.log
for c in c3 c1 c2 c4; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2e_bench_$c.log 2>&1; tail -1 gpurun_out/r2e_bench_$c.log | cut -c1-300; done
python bench.py --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2e_bench_c5.log 2>&1; tail -1 gpurun_out/r2e_bench_c5.log | cut -c1-300
