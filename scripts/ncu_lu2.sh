python scripts/lu_bench.py > gpurun_out/plain_lu2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lp_bwd_local_kernel -s 2 -c 1 -f -o gpurun_out/prof_lu2 python scripts/lu_bench.py > gpurun_out/ncu_lu2.log 2>&1
tail -1 gpurun_out/plain_lu2.log
