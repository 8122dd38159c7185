#!/bin/bash
# Round-2 evidence run: launch list of the bench command, full ncu captures of the headline apply kernel (traffic),
# the streaming assembly kernel and the CSC scatter kernel.  usage: scripts/gpu_round2.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fedvr_apply_pipe_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_fedvr_apply \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1
python scripts/asm_bench.py > gpurun_out/plain_asm_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bspline_assemble_stream -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_stream \
    python scripts/asm_bench.py > gpurun_out/ncu_asm_$TAG.log 2>&1
python scripts/csc_bench.py > gpurun_out/plain_csc_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csc_scatter -s 1 -c 1 -f -o gpurun_out/prof_${TAG}_csc \
    python scripts/csc_bench.py > gpurun_out/ncu_csc_$TAG.log 2>&1
echo round $TAG done
