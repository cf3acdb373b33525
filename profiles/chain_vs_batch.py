"""Event timings of the six persistent recurrent kernels for one forward+backward pass at several batch sizes (default
dims, T=200, L=64): how much a chain launch costs when more rows share the resident weights.
    python profiles/chain_vs_batch.py 4 8"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine, batch_to_device

cfg = pkg.default_params()
dev = torch.device("cuda:0")
T, L = 200, 64
for B in [int(a) for a in sys.argv[1:]] or [4, 8]:
    eng = Engine(cfg, dev, gemm_tf32=1)
    flat, g, bn = eng.flat_from_dict(synth.init_params(cfg, 0)), eng.new_flat(), eng.new_bn_stats()
    bd = batch_to_device(synth.make_batch(cfg, B, T, L, 100), dev)
    masks = eng.generate_masks(B, T, L, 7)
    for _ in range(2):
        eng.forward(flat, bn, bd, masks, outputs=False)
        eng.backward(flat, g)
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.forward(flat, bn, bd, masks, outputs=False)
        eng.backward(flat, g)
    e1.record()
    torch.cuda.synchronize()
    eng.check_abort()
    ev = eng.profile_read()
    eng.profile(False)
    print(f"B={B}: pass {e0.elapsed_time(e1) / 5:.2f} ms; " + "  ".join(f"{k} {ms / max(c, 1) * 1e3:.0f}us" for k, (ms, c) in ev.items()))
    del eng
