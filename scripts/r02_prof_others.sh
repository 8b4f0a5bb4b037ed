#!/bin/bash
# one GPU call: timings, then one `ncu --set full` capture each, of the long-read kernel (C4a shape, minimal-long) and of the
# per-position-table short kernel (C3 shape, NovaSeq-style blob); the captures follow their own commands' clean exits
cd "$(dirname "$0")/.."
timeout 300 python scripts/long_perf.py 400000 > gpurun_out/r02_long_plain.log 2>&1 || exit 1
timeout 300 python scripts/novaseq_perf.py > gpurun_out/r02_novaseq_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_long --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_long python scripts/long_perf.py 400000 > gpurun_out/ncu_r02_long.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_qmpos python scripts/novaseq_perf.py > gpurun_out/ncu_r02_qmpos.log 2>&1
cat gpurun_out/r02_long_plain.log gpurun_out/r02_novaseq_plain.log; tail -2 gpurun_out/ncu_r02_long.log gpurun_out/ncu_r02_qmpos.log
