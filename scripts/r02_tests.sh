#!/bin/bash
# full GPU test suite with per-test and overall time limits; output streamed to a file
cd "$(dirname "$0")/.."
timeout 1700 python -m pytest tests -q -m gpu --timeout 500 -o faulthandler_timeout=450 --durations=15 > gpurun_out/r02_t4.log 2>&1
tail -60 gpurun_out/r02_t4.log
