#!/bin/bash
# kernel work: quick parity subset, then the C2 bench job under environment variants
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 2>&1 | tail -5
bash scripts/r02_bench_iter.sh "$@"
