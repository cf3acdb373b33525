for f in 0 1 2 3; do echo "flags $f"; MSA_REC_FLAGS=$f timeout 300 python profiles/chain_vs_batch.py 4 2>&1 | grep "^B="; done
