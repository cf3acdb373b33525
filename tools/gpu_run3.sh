for f in 1 9; do echo flags $f; MSA_REC_FLAGS=$f timeout 200 python profiles/phase_profile.py 2>&1 | grep "lstm_fwd" | cut -c1-60; done
