# GPU regression call: parity tests, smoke(), the bench line and the reference arm.  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/gpu_tests_smoke_bench.sh v12'
V=${1:-vX}
timeout 1700 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_$V.log 2>&1; echo pytest-exit $?; tail -3 gpurun_out/pytest_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$V.json 2> gpurun_out/bench_$V.err; echo bench-exit $?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$V.json 2> gpurun_out/bench_ref_$V.err; echo ref-exit $?; tail -c 600 gpurun_out/bench_ref_$V.json
python - $V <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print(d['infer']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('latency_bound'), d.get('clocks'))
for k in d['kernels']: print(k['kernel'], round(k['ms_per_launch'],3), round(k.get('share_of_step',0),3), round(k.get('latency_bound',{}).get('frac',0),3))
PY
