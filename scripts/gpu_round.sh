#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list, ncu full capture of the top kernel.
# usage: scripts/gpu_round.sh <tag> [kernel-regex]
TAG=${1:-r1}
KREGEX=${2:-fedvr_apply_pipe_kernel}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/tests_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1
echo round $TAG done
