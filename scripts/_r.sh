timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
ncu --set full --clock-control none -k regex:pb_adam_w2 -c 1 -f -o /tmp/aw2 python scripts/sg_perf.py adam_cfg4 "" 1 > gpurun_out/ncu_aw2.log 2>&1
python scripts/ncu_summary.py /tmp/aw2.ncu-rep gpurun_out/aw2_summary.csv > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pb_adam_w2 -s 3 -c 1 --csv --log-file gpurun_out/aw2_traffic.csv python bench.py --workload adam_cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
tail -3 gpurun_out/aw2_traffic.csv
