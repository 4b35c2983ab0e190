#!/bin/bash
# retry gpurun while the pod answers "busy" (rc 3, nothing charged).  usage: scripts/gpurun_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry $i] busy, sleeping 90 s"; sleep 90
done
exit 3
