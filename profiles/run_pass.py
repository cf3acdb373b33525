"""One FOMAML task (train pass + inner SGD step + test pass) at the bench dimensions; used under ncu.
    python profiles/run_pass.py [n_tasks]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from msa_tts_b200.maml import MAML

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bench.N_TASKS = n
tr = MAML(**bench.trainer_params({"fp32": 0, "tf32x": 1, "tf32": 2}[os.environ.get("MSA_GEMM", "tf32x")]))
items = bench.make_tasks(tr.model_params, pinned=False)
items = {s: {k: tuple(x.to(tr.device) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in items.items()}
for _ in range(int(os.environ.get("MSA_REPS", "2"))):
    tr._metatrain_step(items)
torch.cuda.synchronize()
print("launches", tr.engine.kernel_launches())
