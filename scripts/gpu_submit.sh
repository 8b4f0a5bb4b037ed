#!/bin/bash
# usage: scripts/gpu_submit.sh <timeout_s> '<command>' : re-submits a gpurun call until the pod takes it (a busy pod
# answers "transient" without charging anything)
T=$1; shift
for try in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gpu_submit: gave up after 40 tries"
