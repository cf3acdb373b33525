timeout 900 python profiles/bench_configs.py > gpurun_out/bench_configs_v7.txt 2> gpurun_out/bench_configs_v7.err; echo cfg-exit $?; cat gpurun_out/bench_configs_v7.txt
export MSA_REPS=1
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k ker_infer_rows -s 52 -c 2 -o gpurun_out/prof_infer_rows_v6 -f python profiles/run_infer.py 40 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k ker_infer_attn -s 10 -c 1 -o gpurun_out/prof_infer_attn_v6 -f python profiles/run_infer.py 40 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v7.csv python profiles/run_pass.py 1 > gpurun_out/ncu_v7.log 2>&1
tail -1 gpurun_out/ncu_v7.log
