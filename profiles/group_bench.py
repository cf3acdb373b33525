"""Grouped train passes (G tasks from the same weights in one pass) against G passes one by one: wall time per task and the CUDA-event
time of the persistent recurrences (default dims, B=4, T=200, L=64, bench GEMM policy).
    python profiles/group_bench.py 1 2 4 8"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine, batch_to_device

cfg = pkg.default_params()
dev = torch.device("cuda:0")
B, T, L = 4, 200, 64
eng = Engine(cfg, dev, gemm_tf32=1)
flat = eng.flat_from_dict(synth.init_params(cfg, 0))
for G in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    bds = [batch_to_device(synth.make_batch(cfg, B, T, L, 100 + g), dev) for g in range(G)]
    masks = [eng.generate_masks(B, T, L, 7 + g) for g in range(G)]
    bn = [eng.new_bn_stats() for _ in range(G)]
    grads = [eng.new_flat() for _ in range(G)]

    def run():
        eng.forward_group(flat, bn, bds, masks)
        eng.backward_group(flat, grads)
    for _ in range(2):
        run()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    eng.check_abort()
    ev = eng.profile_read()
    eng.profile(False)
    ms = e0.elapsed_time(e1) / 3
    print(f"G={G}: grouped pass {ms:.2f} ms = {ms / G:.2f} ms/task; " + "  ".join(f"{k} {m / 3 * 1e3:.0f}us" for k, (m, c) in ev.items()), flush=True)
