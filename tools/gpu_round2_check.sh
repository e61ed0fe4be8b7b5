#!/bin/bash
# end-of-round check on one B200: all GPU tests, smoke, the bench lines of the five configurations, the reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
tail -4 gpurun_out/r2l_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; tail -1 gpurun_out/r2l_smoke.log
for c in ${BENCH_CFGS:-c3 c1 c2 c4}; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2l_bench_$c.log 2>&1; grep '^{' gpurun_out/r2l_bench_$c.log | tail -1 | cut -c1-160; grep -vc '^{' gpurun_out/r2l_bench_$c.log; done
timeout 600 python bench.py --config c5 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2l_bench_c5.log 2>&1; grep '^{' gpurun_out/r2l_bench_c5.log | tail -1 | cut -c1-160
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_ref.log 2>&1; grep '^{' gpurun_out/r2l_bench_ref.log | tail -1 | cut -c1-200
