timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench-exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v5.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
for k in d['kernels']: print(k['kernel'], round(k['ms_per_launch'],3), round(k.get('share_of_step',0),3))
PY
