export MSA_REPS=1
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v6.csv python profiles/run_pass.py 1 > gpurun_out/ncu_v6.log 2>&1
tail -1 gpurun_out/ncu_v6.log
