timeout 900 python -m pytest tests/test_parity.py tests/test_api.py tests/test_warp_engine.py -m gpu -x -q -k "cuda_warp or warp_engine" > gpurun_out/w1_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w1_tests.log
python scripts/sg_perf.py sgld_cfg1x "no_w1=1," 3 > gpurun_out/w1_perf.log 2>&1
python scripts/sg_perf.py sgld_cfg1x "" 2 C=4096,T=1000 >> gpurun_out/w1_perf.log 2>&1
tail -n 5 gpurun_out/w1_tests.log; cat gpurun_out/w1_perf.log
