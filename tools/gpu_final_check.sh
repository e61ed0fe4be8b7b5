#!/bin/bash
# final-build check on one B200: all GPU tests (with durations), smoke, bench lines of c3 and c4
mkdir -p gpurun_out
S=$(date +%s)
timeout 1000 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/fin_tests.log 2>&1; echo "tests rc=$? after $(( $(date +%s) - S )) s" >> gpurun_out/fin_tests.log
tail -3 gpurun_out/fin_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/fin_smoke.log
for c in ${BENCH_CFGS:-c3 c4}; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/fin_bench_$c.log 2>&1; grep '^{' gpurun_out/fin_bench_$c.log | tail -1 | cut -c1-120; done
echo "total $(( $(date +%s) - S )) s"
