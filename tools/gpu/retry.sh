#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit code 3): usage tools/gpu/retry.sh LOG TIMEOUT 'command'
LOG=$1; TO=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
