ncu --set full --clock-control none --import-source on -k regex:pb_hmc_tc -c 1 -f -o /tmp/hmc python bench.py --iters 20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hmc.log 2>&1
ncu -i /tmp/hmc.ncu-rep --page source --csv > gpurun_out/hmc_source.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/hmc.ncu-rep gpurun_out/hmc_summary.csv > /dev/null 2>&1
ls -la gpurun_out/hmc_source.csv
