#!/bin/bash
# one `ncu --set full` capture of the short-read format kernel inside the C2 bench job (after the job's own clean run)
cd "$(dirname "$0")/.."
tag=${1:-c2}
ARGS="--steps 1 --warmup 1 --e2e-steps 0 --no-strong --no-to-file --no-cpu-baseline"
timeout 300 python bench.py $ARGS > gpurun_out/r02_ncu_${tag}_plain.json 2> gpurun_out/r02_ncu_${tag}_plain.err || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 5 -c 1 -f -o gpurun_out/prof_r02_${tag} python bench.py $ARGS > gpurun_out/ncu_r02_${tag}.log 2>&1
tail -3 gpurun_out/ncu_r02_${tag}.log
