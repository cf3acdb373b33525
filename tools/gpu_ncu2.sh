export MSA_REPS=1
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_attn_chain_fwd' -s 0 -c 1 -f -o gpurun_out/prof_attnfwd_v4 python profiles/run_pass.py 1 > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/ncu_fwd.log
