timeout 600 python -m pytest tests/test_gpu_infer.py -m gpu -x -q --timeout 300 2>&1 | tail -5
for v in 0 1 2 3; do echo "MSA_IR_CFG=$v"; MSA_IR_CFG=$v MSA_REPS=3 timeout 300 python profiles/run_infer.py 1000 2>&1 | tail -1; done
MSA_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__cycles_active.avg --clock-control none --cache-control none -k regex:ker_infer -s 60 -c 12 --csv --log-file gpurun_out/launches_infer_v4w.csv python profiles/run_infer.py 40 > gpurun_out/ncu_infer_v4w.log 2>&1
grep -E "gpu__time_duration" gpurun_out/launches_infer_v4w.csv | awk -F'","' '{print substr($5,1,50), $(NF)}'
