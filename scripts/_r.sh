timeout 900 python -m pytest tests/test_warp_engine.py -m gpu -x -q > gpurun_out/w2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w2_tests.log
tail -n 12 gpurun_out/w2_tests.log
