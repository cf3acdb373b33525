set -x
export MSA_REPS=1
python profiles/run_pass.py 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_attn_chain_fwd|k_lstm_rec_fwd' -s 1 -c 2 -f -o gpurun_out/prof_fwd_v2 python profiles/run_pass.py 1 > gpurun_out/ncu_fwd.log 2>&1
tail -5 gpurun_out/ncu_fwd.log
ls -la gpurun_out/
