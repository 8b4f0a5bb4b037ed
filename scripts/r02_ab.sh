#!/bin/bash
# A/B of library builds on the C2 job inside one GPU call: r02_ab.sh <variant name> ... (simmr_b200/variants/lib_<name>.so)
cd "$(dirname "$0")/.."
for v in "$@"; do SIMMR_B200_LIB=simmr_b200/variants/lib_$v.so timeout 300 python scripts/c2_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02_ab.log; done
