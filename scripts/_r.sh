timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -n 3 gpurun_out/gpu_tests.log
timeout 900 python bench.py > gpurun_out/bench_n1.jsonl 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.jsonl 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
