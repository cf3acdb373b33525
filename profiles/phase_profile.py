"""Per-phase cycle breakdown of the persistent recurrent kernels at the bench dimensions (thread 0 of every CTA,
clock64 deltas summed over the T steps of one launch).  Not a benchmark: explains where a step's time goes.
    python profiles/phase_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from msa_tts_b200.maml import MAML

bench.N_TASKS = 1
tr = MAML(**bench.trainer_params(1))
items = bench.make_tasks(tr.model_params, pinned=False)
items = {s: {k: tuple(x.to(tr.device) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in items.items()}
for _ in range(2):
    tr._metatrain_step(items)
eng = tr.engine
eng.profile(True, inkernel=True)
tr._metatrain_step(items)
torch.cuda.synchronize()
eng.check_abort()
ev = eng.profile_read()
steps = {"enc_lstm_fwd": bench.L, "enc_lstm_bwd": bench.L, "attn_chain_fwd": bench.T, "attn_chain_bwd": bench.T,
         "dec_lstm_fwd": bench.T, "dec_lstm_bwd": bench.T}
for name, (ms, cnt) in ev.items():
    ph = eng.profile_phases(name)
    act = [r for r in ph if sum(r) > 0]
    n = steps[name]
    tot = [sum(r) for r in act]
    print(f"{name}: {ms / max(cnt, 1) * 1e3:.1f} us/launch (events), {len(act)} CTAs reporting, "
          f"{sum(tot) / len(tot) / n:.0f} cycles/step (mean over CTAs)")
    for j in range(8):
        col = [r[j] / n for r in act]
        if max(col) > 0:
            print(f"   phase {j}: mean {sum(col) / len(col):8.0f}  min {min(col):8.0f}  max {max(col):8.0f}  cycles/step")
eng.profile(False)
