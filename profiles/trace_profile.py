"""Per-warp timeline of a few steps of a persistent recurrent kernel (development tool, see rec_common.cuh ChainProf).
    MSA_REC_FLAGS=<0..3> python profiles/trace_profile.py [kernel] [t0]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from msa_tts_b200 import _lib
from msa_tts_b200.maml import MAML

kernel = sys.argv[1] if len(sys.argv) > 1 else "attn_chain_fwd"
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bench.N_TASKS = 1
tr = MAML(**bench.trainer_params(1))
items = bench.make_tasks(tr.model_params, pinned=False)
items = {s: {k: tuple(x.to(tr.device) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in items.items()}
for _ in range(2):
    tr._metatrain_step(items)
eng = tr.engine
eng.profile(True, inkernel=True)
_lib.check(eng.lib.msa_profile_trace_step(eng.h, t0))
tr._metatrain_step(items)
torch.cuda.synchronize()
eng.check_abort()
ev = eng.profile_read()
print("flags", os.environ.get("MSA_REC_FLAGS", "0"), {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in ev.items()}, "us/launch")
tr_ = eng.profile_trace(kernel)          # [cta][warp][step][tag][2]
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"trace_{kernel}_f{os.environ.get('MSA_REC_FLAGS', '0')}.npy"), tr_)
clk, gt = tr_[..., 0], tr_[..., 1]
ntag = clk.shape[3]
used = [j for j in range(ntag) if clk[:, :, :, j].max() > 0]
print(kernel, "tags used", used)
# per-step duration from globaltimer of the last tag, CTA 0 warp 0
last = used[-1]
for st in range(1, 4):
    d = gt[:, 0, st, last] - gt[:, 0, st - 1, last]
    print(f"step {t0 + st}: step time (globaltimer ns, over CTAs) min {d.min()} median {int(np.median(d))} max {d.max()}")
# timeline of step t0+1 relative to the end of step t0 (per CTA clock), warps min..max
st = 1
for cta in (0, 1, 73, 147):
    base = clk[cta, :, st - 1, last].max()
    print(f"CTA {cta}: cycles since the last warp finished step {t0}")
    for j in used:
        v = clk[cta, :, st, j] - base
        v = v[clk[cta, :, st, j] > 0]
        if len(v):
            print(f"   tag {j:2d}: first warp {v.min():7d}  last warp {v.max():7d}   (warp0 {clk[cta, 0, st, j] - base:7d})")
# cross-CTA skew at each tag (globaltimer ns)
print("cross-CTA skew of warp 0 at each tag of step", t0 + 1, "(ns, relative to the earliest CTA)")
for j in used:
    g = gt[:, 0, st, j]
    print(f"   tag {j:2d}: spread {g.max() - g.min():6d} ns   mean offset from tag {used[0]}: {np.mean(g - gt[:, 0, st, used[0]]):8.0f} ns")
eng.profile(False)
