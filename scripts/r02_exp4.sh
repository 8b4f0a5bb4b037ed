#!/bin/bash
cd "$(dirname "$0")/.."
export QP_PROF=minimal QP_NOSINK=1 QP_ITERS=2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_pairs --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_pairs python scripts/quick_perf.py 20 40e6 > gpurun_out/ncu_r02_pairs.log 2>&1
tail -3 gpurun_out/ncu_r02_pairs.log
