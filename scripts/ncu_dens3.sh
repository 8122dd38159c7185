PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/plain_k10d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:density -s 7 -c 7 --csv --log-file gpurun_out/k10d_list.csv env PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/ncu_k10d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:density_bwd_local2_kernel -s 1 -c 1 -f -o gpurun_out/prof_k10d env PROF_SINGLE=1 python scripts/prof_one.py density > gpurun_out/ncu_k10d2.log 2>&1
