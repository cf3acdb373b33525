timeout 600 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 300 2>&1 | tail -5
for v in 0 1 2 3; do echo "MSA_IR_CFG=$v"; MSA_IR_CFG=$v MSA_REPS=3 timeout 300 python profiles/run_infer.py 1000 2>&1 | tail -1; done
MSA_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_infer_v3.csv python profiles/run_infer.py 20 > gpurun_out/ncu_infer_v3.log 2>&1
tail -1 gpurun_out/ncu_infer_v3.log
