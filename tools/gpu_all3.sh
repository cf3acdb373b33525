timeout 900 python -m pytest tests/test_gpu_pass.py -m gpu -x -q --timeout 300 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo bench-exit $?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
for k in d['kernels']: print(k['kernel'], round(k['ms_per_launch'],3), round(k.get('share_of_step',0),3))
PY
