#!/bin/bash
# kernel work: the C2 bench job (device-resident part only) under environment variants; usage: r02_bench_iter.sh VAR=x VAR=y ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 3 --warmup 2 --e2e-steps 0 --no-strong --no-to-file --no-cpu-baseline 2> gpurun_out/r02_bi.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$v', 'value %.1f ms/step %.2f frac %.4f launch_ms %.3f alone %.4f (%.3f ms) share %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['launch_ms'], r['alone']['frac'], r['alone']['launch_ms'], r['kernel_share_of_step']))
" | tee -a gpurun_out/r02_bench_iter.log
  tail -3 gpurun_out/r02_bi.err
done
