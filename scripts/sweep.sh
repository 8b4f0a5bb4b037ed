#!/bin/bash
# sweep of (table copies, warps per CTA) for the format kernel on the quick probe
export QP_NOSINK=1 QP_ITERS=3 QP_PROF=${QP_PROF:-minimal}
for cfg in "1 8" "1 24" "4 24" "8 24" "16 20" "16 16" "8 16"; do
  set -- $cfg
  echo "rep=$1 warps=$2: $(SIMMR_DEBUG_SYNC=1 SIMMR_REP=$1 SIMMR_WARPS=$2 python scripts/quick_perf.py 20 16e6 2>&1 | grep -E 'occ=|Gbases' | tail -2 | sed 's/.*threads/threads/; s/MinimalShortErrorProfile: 16000000 rec 2391904392 bases//' | tr '\n' ' ')"
done
