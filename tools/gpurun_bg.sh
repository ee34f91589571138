#!/bin/bash
# usage: tools/gpurun_bg.sh <tag> <timeout_s> <command...>  — retries while the pod answers busy (exit 3); log in /tmp/gpurun_<tag>.log
tag=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > /tmp/gpurun_$tag.log 2>&1; rc=$?
  echo "rc=$rc attempt=$i" >> /tmp/gpurun_$tag.log
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo DONE >> /tmp/gpurun_$tag.log
