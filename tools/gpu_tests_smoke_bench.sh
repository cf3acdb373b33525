timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; echo bench-exit $?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v9.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['infer']); print(d['roofline']['kernel'], d['roofline']['frac'], d.get('clocks'))
for k in d['kernels']: print(k['kernel'], round(k['ms_per_launch'],3), round(k.get('share_of_step',0),3))
PY
