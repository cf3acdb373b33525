set -x
timeout 600 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 300 2>&1 | tail -15
MSA_REPS=3 timeout 300 python profiles/run_infer.py 1000 2>&1 | tail -4
MSA_INFER_GRAPH=1 MSA_REPS=2 timeout 300 python profiles/run_infer.py 1000 2>&1 | tail -3
MSA_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_infer_v2.csv python profiles/run_infer.py 20 > gpurun_out/ncu_infer_v2.log 2>&1
tail -2 gpurun_out/ncu_infer_v2.log
