#!/bin/bash
# long-read kernel work: parity tests that touch long reads, the long-read and NovaSeq probes, one ncu capture of the per-position short kernel
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x --timeout 400 -k "long or c4 or kmer or cli" 2>&1 | tail -5
timeout 300 python scripts/long_perf.py 400000 2>&1 | tail -2 | tee -a gpurun_out/r02_long_iter.log
timeout 300 python scripts/novaseq_perf.py 2>&1 | tail -1 | tee -a gpurun_out/r02_long_iter.log
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_qmpos python scripts/novaseq_perf.py > gpurun_out/ncu_r02_qmpos.log 2>&1
tail -2 gpurun_out/ncu_r02_qmpos.log
fi
