MSA_REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:ker_infer -s 60 -c 12 --csv --log-file gpurun_out/launches_infer_v7w.csv python profiles/run_infer.py 40 > gpurun_out/ncu_infer_v7w.log 2>&1
grep -E "gpu__time_duration" gpurun_out/launches_infer_v7w.csv | awk -F'","' '{print substr($5,1,50), $(NF)}'
MSA_REPS=4 timeout 120 python profiles/run_infer.py 1000 2>&1 | tail -3
