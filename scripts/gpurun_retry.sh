#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "no box/slot free", rc 3)
for try in $(seq 1 30); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry] attempt $try answered busy; sleeping 90 s"
    sleep 90
done
exit 3
