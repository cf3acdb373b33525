timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_v4.log 2>&1; echo "pytest-exit $?" >> gpurun_out/pytest_v4.log
tail -3 gpurun_out/pytest_v4.log
for f in 0 1 2 3; do MSA_REC_FLAGS=$f timeout 200 python profiles/trace_profile.py attn_chain_fwd 100 > gpurun_out/trace_attn_chain_fwd_f$f.txt 2>&1; head -1 gpurun_out/trace_attn_chain_fwd_f$f.txt; done
MSA_REC_FLAGS=0 timeout 200 python profiles/trace_profile.py attn_chain_bwd 100 > gpurun_out/trace_attn_chain_bwd_f0.txt 2>&1
