#!/bin/bash
# builds the working tree's library as simmr_b200/variants/lib_<name>.so (for A/B runs in one GPU call: SIMMR_B200_LIB=...)
set -e
cd "$(dirname "$0")/../simmr_b200/csrc"
mkdir -p ../variants
# usage: build_variant.sh <name> [extra nvcc flags, e.g. -DSIMMR_KO=1]; the in-tree library is rebuilt without them afterwards
if [ -n "$2" ]; then touch engine.cu; make -s EXTRA="$2"; cp ../libsimmr_b200.so ../variants/lib_$1.so; touch engine.cu; make -s
else make -s; cp ../libsimmr_b200.so ../variants/lib_$1.so; fi
echo "built variants/lib_$1.so"
