for C in 148 296 592 600 1024 1184 2048 2368; do
python scripts/sg_perf.py sgld_cfg1x "w1_min_chains=1,w1_min_chains=1+w1_no_spread=1" 2 C=$C,T=1000 >> gpurun_out/w1_sweep.log 2>&1
done
cat gpurun_out/w1_sweep.log
