N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_v7_n$N.json 2> gpurun_out/bench_v7_n$N.err; echo exit $?
tail -c 1500 gpurun_out/bench_v7_n$N.json | head -c 600; tail -3 gpurun_out/bench_v7_n$N.err
