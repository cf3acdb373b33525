timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo bench-exit $?; tail -3 gpurun_out/bench_v8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(d['infer']); print(d.get('cpu_baseline'))
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-600
