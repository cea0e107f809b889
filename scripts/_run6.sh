timeout 300 python scripts/pred50_check.py > gpurun_out/pred50.log 2>&1; tail -n 12 gpurun_out/pred50.log
timeout 900 python bench.py > gpurun_out/bench_n1.jsonl 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.jsonl').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['roofline']['frac'])
for k,v in d['workloads'].items(): print(k, round(v['value']/1e6,3), v['unit'], v['roofline']['engine'], round(v['roofline']['frac'],4), v['predictive']['ms'])
PY
