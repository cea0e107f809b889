timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -n 6 gpurun_out/gpu_tests.log
timeout 600 python scripts/sg_perf.py sgld_cfg3,psgld_cfg3,sghmc_cfg3,sgld_cfg5,sgld_cfg1x "" 2 > gpurun_out/sg_perf.log 2>&1
cat gpurun_out/sg_perf.log
