"""Config-5 shaped inference (B=32, L=64, default dims) for a given number of decoder steps; used under ncu.
    python profiles/run_infer.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = pkg.default_params()
cfg["max_decoder_steps"] = steps
cfg["decoder_no_early_stopping"] = True
eng = Engine(cfg, torch.device("cuda:0"), gemm_tf32=int(os.environ.get("MSA_POLICY", "0")))     # 0 fp32-accurate, 2 TF32
B = 32
lens = torch.arange(64, 64 - B, -1)
inp = torch.randint(1, 123, (B, 64))
for b in range(B):
    inp[b, lens[b]:] = 0
spk = torch.randn(B, cfg["speaker_embedding_dim"])
pm = synth.make_infer_masks(cfg, B, steps, 5)
flat, bn = eng.flat_from_dict(synth.init_params(cfg, 0)), eng.new_bn_stats()
for _ in range(int(os.environ.get("MSA_REPS", "2"))):
    t0 = time.perf_counter()
    out = eng.infer(flat, bn, inp, lens, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{steps} steps: {dt * 1e3:.1f} ms, {dt / steps * 1e6:.1f} us/step, {B * steps / dt:.0f} mel-frames/s")
