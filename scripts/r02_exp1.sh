#!/bin/bash
# one GPU call: full GPU test suite, then the format kernel under experiment switches, then one ncu capture
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r02_t2.log
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
rm -f gpurun_out/r02_exp1.log
echo "== SIMMR_TILES=linear" >> gpurun_out/r02_exp1.log
SIMMR_TILES=linear python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp1.log
for dbg in 0 1 2 4 5; do
  echo "== sorted SIMMR_DBG=$dbg" >> gpurun_out/r02_exp1.log
  SIMMR_DBG=$dbg python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp1.log
done
for w in 14 15; do
  echo "== sorted SIMMR_WARPS=$w" >> gpurun_out/r02_exp1.log
  SIMMR_WARPS=$w python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp1.log
done
QP_ITERS=2 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_sorted python scripts/quick_perf.py 20 40e6 > gpurun_out/ncu_r02_sorted.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
cat gpurun_out/r02_t2.log gpurun_out/r02_exp1.log; tail -c 1500 gpurun_out/r02_bench2.err
