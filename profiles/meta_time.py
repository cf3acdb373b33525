"""Wall time (CUDA events) of the bench meta-step (8 tasks, device-resident batches) without the extra legs of bench.py.
    python profiles/meta_time.py [reps]      (env MSA_CONV_TC / MSA_PT_GROUP / MSA_GEMM_PDL ... select variants)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from msa_tts_b200.maml import MAML

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tr = MAML(**bench.trainer_params(1))
items = bench.make_tasks(tr.model_params, pinned=False)
items = {s: {k: tuple(x.to(tr.device) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in items.items()}
for _ in range(3):
    tr._metatrain_step(items)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    tr._metatrain_step(items)
e1.record()
torch.cuda.synchronize()
tr.engine.check_abort()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("MSA_"))
print(f"[{tag}] {e0.elapsed_time(e1) / reps:.2f} ms per meta-step")
