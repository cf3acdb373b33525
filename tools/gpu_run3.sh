for f in 0; do MSA_REC_FLAGS=-1 timeout 200 python profiles/trace_profile.py attn_chain_bwd 100 > gpurun_out/trace_attn_chain_bwd_f$f.txt 2>&1; head -1 gpurun_out/trace_attn_chain_bwd_f$f.txt; done
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
