for C in 32 64 148 296 444; do
python scripts/sg_perf.py sgld_cfg1x "no_w1=1,w1_min_chains=1" 2 C=$C,T=2000 >> gpurun_out/w1_sweep2.log 2>&1
done
cat gpurun_out/w1_sweep2.log
