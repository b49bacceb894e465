#!/bin/bash
# gpurun with retry on "busy" (exit 3): [GPUS=n] scripts/gpu.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T ${GPUS:+--gpus $GPUS} -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
