python scripts/lu_bench.py > gpurun_out/plain_lu.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"lp_" -c 40 --csv --log-file gpurun_out/lu_list.csv python scripts/lu_bench.py > gpurun_out/ncu_lu.log 2>&1
tail -3 gpurun_out/plain_lu.log
