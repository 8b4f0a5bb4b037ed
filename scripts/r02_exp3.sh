#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x > gpurun_out/r02_t5.log 2>&1
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
rm -f gpurun_out/r02_exp3.log
for p in 0 1; do
  echo "== SIMMR_PAIRS=$p" >> gpurun_out/r02_exp3.log
  SIMMR_PAIRS=$p timeout 120 python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp3.log
done
for w in 12 14 15; do
  echo "== pairs SIMMR_WARPS=$w" >> gpurun_out/r02_exp3.log
  SIMMR_WARPS=$w timeout 120 python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp3.log
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-strong --no-to-file --no-cpu-baseline > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
SIMMR_PAIRS=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-strong --no-to-file --no-cpu-baseline > gpurun_out/r02_bench4_lanes.json 2> gpurun_out/r02_bench4_lanes.err
tail -25 gpurun_out/r02_t5.log; cat gpurun_out/r02_exp3.log; tail -c 400 gpurun_out/r02_bench4.err
