export MSA_REPS=1
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_attn_chain' -s 0 -c 2 -f -o gpurun_out/prof_attn_v5 python profiles/run_pass.py 1 > gpurun_out/ncu_attn_v5.log 2>&1
tail -2 gpurun_out/ncu_attn_v5.log
