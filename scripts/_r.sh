timeout 900 python -m pytest tests -m gpu -x -q -k "adam or ensemble or train or map or engines" > gpurun_out/adam_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/adam_tests.log
tail -n 3 gpurun_out/adam_tests.log
timeout 300 python scripts/sg_perf.py adam_cfg4 "no_w2=1," 3 > gpurun_out/adam_perf.log 2>&1; cat gpurun_out/adam_perf.log
