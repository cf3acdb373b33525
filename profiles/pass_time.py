"""Wall time (CUDA events) of one forward+backward pass at the bench dimensions under the bench GEMM policy.
    python profiles/pass_time.py [reps]      (env MSA_CONV_TC=0 / MSA_GEMM_TC=... select variants)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine, batch_to_device

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = pkg.default_params()
dev = torch.device("cuda:0")
B, T, L = 4, 200, 64
eng = Engine(cfg, dev, gemm_tf32=1)
flat = eng.flat_from_dict(synth.init_params(cfg, 0))
bd = batch_to_device(synth.make_batch(cfg, B, T, L, 100), dev)
masks = eng.generate_masks(B, T, L, 7)
bn, g = eng.new_bn_stats(), eng.new_flat()
for _ in range(3):
    eng.forward(flat, bn, bd, masks, outputs=False)
    eng.backward(flat, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.forward(flat, bn, bd, masks, outputs=False)
    eng.backward(flat, g)
e1.record()
torch.cuda.synchronize()
eng.check_abort()
print(f"MSA_CONV_TC={os.environ.get('MSA_CONV_TC', '1')} MSA_GEMM_TC={os.environ.get('MSA_GEMM_TC', '2')}: {e0.elapsed_time(e1) / reps:.3f} ms per forward+backward pass")
