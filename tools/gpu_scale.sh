# Scaling run on ONE box: bash tools/gpu_scale.sh v13 2 4 8   (under gpurun --gpus 8)
V=$1; shift
for N in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${V}_n$N.json 2> gpurun_out/bench_${V}_n$N.err; echo N=$N exit $?
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${V}_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['infer'].get('value'), d['infer'].get('fp32_accurate'))"
done
