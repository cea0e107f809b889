timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.jsonl 2> gpurun_out/bench_n8.err; echo "rc=$?"
tail -c 200 gpurun_out/bench_n8.err
