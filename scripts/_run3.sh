python scripts/sg_perf.py sgld_cfg1x "" 1 > gpurun_out/ncu_w1_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/w1_launches.csv python scripts/sg_perf.py sgld_cfg1x "" 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pb_sgw -c 1 -f -o /tmp/w1_sgld python scripts/sg_perf.py sgld_cfg1x "" 1 > gpurun_out/ncu_w1.log 2>&1
ncu -i /tmp/w1_sgld.ncu-rep --page raw --csv > gpurun_out/w1_sgld_raw.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/w1_sgld.ncu-rep gpurun_out/w1_sgld_summary.csv > /dev/null 2>&1
python scripts/sg_perf.py psgld_cfg3,sghmc_cfg3,adam_cfg4 "" 2 > gpurun_out/sg_perf2.log 2>&1
cat gpurun_out/sg_perf2.log
