#!/bin/bash
# one GPU call: full GPU test suite (no -x: every failure is worth seeing), prefetch experiments, bench
cd "$(dirname "$0")/.."
python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/r02_t3.log
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
rm -f gpurun_out/r02_exp2.log
for dbg in 0 8 16 24 32 40 56; do
  echo "== SIMMR_DBG=$dbg" >> gpurun_out/r02_exp2.log
  SIMMR_DBG=$dbg python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp2.log
done
for w in 14 15; do
  echo "== SIMMR_WARPS=$w" >> gpurun_out/r02_exp2.log
  SIMMR_WARPS=$w python scripts/quick_perf.py 20 40e6 2>&1 | grep Minimal >> gpurun_out/r02_exp2.log
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
SIMMR_DBG=56 python bench.py --steps 5 --warmup 3 --no-strong --no-to-file --no-cpu-baseline > gpurun_out/r02_bench3_pf.json 2> gpurun_out/r02_bench3_pf.err
cat gpurun_out/r02_t3.log gpurun_out/r02_exp2.log; tail -c 800 gpurun_out/r02_bench3.err
