timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6
MSA_REPS=3 timeout 120 python profiles/run_infer.py 1000 2>&1 | tail -1
MSA_INFER_GRAPH=1 MSA_REPS=3 timeout 120 python profiles/run_infer.py 1000 2>&1 | tail -1
