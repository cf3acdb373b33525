timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench-exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v4.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['cpu_baseline'])
for k in d['kernels']: print(k['kernel'], round(k['ms_per_launch'],3), round(k.get('share_of_step',0),3))
PY
export MSA_REPS=1
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv python profiles/run_pass.py 1 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
