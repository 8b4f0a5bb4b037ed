#!/bin/bash
# one GPU call: full GPU test suite (every failure is worth seeing; per-test and overall time limits), prefetch experiments, bench
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu -p pytest_timeout --timeout 420 -o faulthandler_timeout=400 > gpurun_out/r02_t3.log 2>&1
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
rm -f gpurun_out/r02_exp2.log
for dbg in 0 8 16 24 32 40 56; do
  echo "== SIMMR_DBG=$dbg" >> gpurun_out/r02_exp2.log
  SIMMR_DBG=$dbg timeout 120 python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp2.log
done
for w in 14 15; do
  echo "== SIMMR_WARPS=$w" >> gpurun_out/r02_exp2.log
  SIMMR_WARPS=$w timeout 120 python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp2.log
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
SIMMR_DBG=56 timeout 300 python bench.py --steps 5 --warmup 3 --no-strong --no-to-file --no-cpu-baseline > gpurun_out/r02_bench3_pf.json 2> gpurun_out/r02_bench3_pf.err
tail -40 gpurun_out/r02_t3.log; cat gpurun_out/r02_exp2.log; tail -c 800 gpurun_out/r02_bench3.err
