#!/bin/bash
# one GPU call: full GPU test suite, then the new format kernel under experiment switches, then one ncu capture
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_t2.log
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
for dbg in 0 1 2 4 5; do
  echo "== SIMMR_DBG=$dbg" >> gpurun_out/r02_exp1.log
  SIMMR_DBG=$dbg python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp1.log
done
for w in 14 15; do
  echo "== SIMMR_WARPS=$w" >> gpurun_out/r02_exp1.log
  SIMMR_WARPS=$w python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp1.log
done
QP_ITERS=2 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_sorted python scripts/quick_perf.py 20 40e6 > gpurun_out/ncu_r02_sorted.log 2>&1
cat gpurun_out/r02_t2.log gpurun_out/r02_exp1.log
