#!/bin/bash
# usage: retry.sh <logfile> <timeout> <command string>
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "done rc=$rc try=$i" >> $log; exit $rc; fi
  sleep 90
done
echo "gave up" >> $log; exit 3
