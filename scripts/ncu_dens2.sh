PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/plain_k10c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"density_bwd_local_kernel|density_bwd_fix_kernel" -s 4 -c 2 -f -o gpurun_out/prof_k10c env PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/ncu_k10c.log 2>&1
tail -2 gpurun_out/plain_k10c.log
