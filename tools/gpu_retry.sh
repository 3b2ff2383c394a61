#!/bin/bash
# tools/gpu_retry.sh <timeout> '<command>' [gpurun flags]: like tools/gpu.sh, retrying while the pod answers busy (exit 3)
cd /root/repo
for i in $(seq 1 40); do
  tools/gpu.sh "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
