timeout 300 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 120 2>&1 | tail -2
python profiles/run_infer.py 1000
MSA_REPS=1 python profiles/run_infer.py 20 > gpurun_out/plain.log 2>&1 && MSA_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infer.csv python profiles/run_infer.py 20 > gpurun_out/ncu_infer.log 2>&1
