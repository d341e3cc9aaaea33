#!/bin/bash
log=$1; to=$2; g=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $g --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "done rc=$rc try=$i" >> $log; exit $rc; fi
  sleep 90
done
echo "gave up" >> $log; exit 3
