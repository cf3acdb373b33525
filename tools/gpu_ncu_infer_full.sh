export MSA_REPS=1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"ker_infer_lstm_tma|ker_infer_attn|ker_infer_prenet|ker_infer_proj" -s 40 -c 5 -o gpurun_out/prof_infer_v9 -f python profiles/run_infer.py 40 > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
