#!/bin/bash
# usage: gpu_retry.sh LOGFILE TIMEOUT [--gpus N] -- COMMAND   : re-submits to gpurun while the pod answers "busy" (nothing charged)
log=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "gpu_retry finished rc=$rc after $i attempt(s)" >> "$log"
