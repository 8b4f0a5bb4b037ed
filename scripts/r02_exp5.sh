#!/bin/bash
cd "$(dirname "$0")/.."
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
rm -f gpurun_out/r02_exp5.log
for p in 0 1; do
  echo "== SIMMR_PAIRS=$p" >> gpurun_out/r02_exp5.log
  SIMMR_PAIRS=$p timeout 120 python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp5.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "minimal or edge or novaseq or batching" > gpurun_out/r02_t6.log 2>&1
tail -5 gpurun_out/r02_t6.log; cat gpurun_out/r02_exp5.log
