#!/bin/bash
# one GPU call of the kernel work: parity tests (all but the slow ONT config), then the short-read probe
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 -k "not c4_long" 2>&1 | tail -8 > gpurun_out/r02_iter_tests.log
cat gpurun_out/r02_iter_tests.log
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=3
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/r02_iter_perf.log
  env $v timeout 200 python scripts/quick_perf.py 20 40e6 2>&1 | grep -E "Minimal|rror" | tee -a gpurun_out/r02_iter_perf.log
done
