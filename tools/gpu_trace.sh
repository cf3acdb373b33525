for k in attn_chain_fwd attn_chain_bwd dec_lstm_fwd dec_lstm_bwd; do MSA_REC_FLAGS=0 timeout 200 python profiles/trace_profile.py $k 100 > gpurun_out/trace_${k}_f0.txt 2>&1; done
head -3 gpurun_out/trace_attn_chain_bwd_f0.txt
