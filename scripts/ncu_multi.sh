#!/bin/bash
# usage: scripts/ncu_multi.sh <what> <kernel-regex> <tag> <count>   — full ncu capture of <count> launches matching the regex
PROF_SINGLE=1 python scripts/prof_one.py $1 > gpurun_out/plain_$3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${5:-8} -c $4 -f -o gpurun_out/prof_$3 env PROF_SINGLE=1 python scripts/prof_one.py $1 > gpurun_out/ncu_$3.log 2>&1
tail -2 gpurun_out/plain_$3.log
