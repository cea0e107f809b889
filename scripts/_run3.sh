ncu --set full --clock-control none --import-source on -k regex:pb_sgw -c 1 -f -o /tmp/w1_sgld python scripts/sg_perf.py sgld_cfg1x "" 1 > gpurun_out/ncu_w1.log 2>&1
ls -la /tmp/w1_sgld.ncu-rep >> gpurun_out/ncu_w1.log
ncu -i /tmp/w1_sgld.ncu-rep --page raw --csv > gpurun_out/w1_sgld_raw.csv 2>/dev/null
ncu -i /tmp/w1_sgld.ncu-rep --page source --csv > gpurun_out/w1_sgld_source.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/w1_sgld.ncu-rep gpurun_out/w1_sgld_summary.csv > /dev/null 2>&1
ls -la gpurun_out/
