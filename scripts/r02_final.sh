#!/bin/bash
# the round's evidence in one GPU call: full GPU test suite, the bench line (both arms), the ncu launch list of the bench
# command and one `ncu --set full` capture each of the three format kernels (after their commands' own clean runs)
cd "$(dirname "$0")/.."
if [ "$1" != "notests" ]; then
timeout 1700 python -m pytest tests -q -m gpu --timeout 500 -o faulthandler_timeout=450 --durations=8 > gpurun_out/r02_final_tests.log 2>&1
tail -14 gpurun_out/r02_final_tests.log
fi
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || { tail -5 gpurun_out/r02_bench_final.err; }
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_ref.json 2> gpurun_out/r02_bench_final_ref.err
ARGS="--steps 2 --warmup 1 --e2e-steps 0 --no-strong --no-to-file --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 5 -c 1 -f -o gpurun_out/prof_r02_c2final python bench.py $ARGS > gpurun_out/ncu_r02_c2final.log 2>&1
timeout 300 python scripts/long_perf.py 400000 > gpurun_out/r02_long_final.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_long --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_long_final python scripts/long_perf.py 400000 > gpurun_out/ncu_r02_long_final.log 2>&1
timeout 300 python scripts/novaseq_perf.py > gpurun_out/r02_novaseq_final.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_format_lanes --launch-skip 1 -c 1 -f -o gpurun_out/prof_r02_qmpos_final python scripts/novaseq_perf.py > gpurun_out/ncu_r02_qmpos_final.log 2>&1
cat gpurun_out/r02_long_final.log gpurun_out/r02_novaseq_final.log | tail -3
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_final.json", "gpurun_out/r02_bench_final_ref.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); r = d.get("roofline") or {}
            print(f, "value", d["value"], "e2e", d["e2e"]["value"], "frac", r.get("frac"), "alone", (r.get("alone") or {}).get("frac"), "strong", (d.get("strong") or {}).get("value"), "to_file", json.dumps(d.get("to_file"))[:300])
PY
