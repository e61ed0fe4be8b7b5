#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
tail -6 gpurun_out/r2j_tests.log
{
echo "== c4 group 3 lanes 12"; L3D_GROUP=3 L3D_PROBE_CFG=c4 python tools/skip_probe.py 12 24
echo "== c4 group 7 lanes 21"; L3D_PROBE_CFG=c4 python tools/skip_probe.py 21 42
echo "== c4 prio variants"; L3D_PRIO=0,0,0 L3D_PROBE_CFG=c4 python tools/skip_probe.py 14 28
} > gpurun_out/r2j_pipe.log 2>&1
cat gpurun_out/r2j_pipe.log
for c in c3 c1 c2 c4; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2j_bench_$c.log 2>&1; grep '^{' gpurun_out/r2j_bench_$c.log | tail -1 | cut -c1-200; grep -vc '^{' gpurun_out/r2j_bench_$c.log; done
python bench.py --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c5.log 2>&1; grep '^{' gpurun_out/r2j_bench_c5.log | tail -1 | cut -c1-200
