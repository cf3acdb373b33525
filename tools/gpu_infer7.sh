MSA_REPS=1 timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k ker_infer_rows -s 52 -c 1 -o gpurun_out/prof_infer_lstm_v5 -f python profiles/run_infer.py 40 > gpurun_out/ncu_infer_v5f.log 2>&1
tail -2 gpurun_out/ncu_infer_v5f.log
