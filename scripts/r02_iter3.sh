#!/bin/bash
# kernel work: parity subset with the in-tree library, then an A/B of library builds on the C2 job
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -q -m gpu -x --timeout 300 -k "not c4_long" 2>&1 | tail -5
bash scripts/r02_ab.sh "$@"
