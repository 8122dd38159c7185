#!/bin/bash
# Evidence run after the FE-DVR assembly rewrite: launch list of the bench command, full ncu captures of the headline
# apply kernel (traffic), the assembly rows kernel and the density right-hand-side kernel.  usage: scripts/gpu_round4.sh <tag>
TAG=${1:-r4}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fedvr_apply_pipe_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_fedvr_apply \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fedvr_assemble_rows_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_fedvr_asm \
    python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_full2_$TAG.log 2>&1
PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/plain_dens_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:density_rhs_kernel -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_density_rhs \
    env PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/ncu_dens_$TAG.log 2>&1
echo round $TAG done
