#!/bin/bash
# gpu_retry.sh TIMEOUT 'command' — gpurun with retries while the pod answers busy/transient (nothing is charged for those).
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
