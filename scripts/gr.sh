#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient"): usage: scripts/gr.sh <timeout_s> '<command>'
t=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 60; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gr.sh: gave up after 30 transient answers"; exit 3
