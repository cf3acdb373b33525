""""PyTorch-eager-on-B200 running the same modules" -- the library-kernel bar on the same box.  The reference tree
cannot travel to the GPU box, so the oracle restatement (plain torch ops, oracle/model.py, pinned to the reference by the golden
fixtures) is run with CUDA tensors: cuBLAS / cuDNN / ATen kernels launched op by op from Python, exactly how the reference
trainers would run on a GPU.  A reported baseline, never part of the product path: it lives under tests/ because only tests/,
smoke() and bench.py's CPU legs may execute the oracle.
    python tests/baselines/eager_gpu_baseline.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from oracle import model as OM

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = pkg.default_params()
crit = dict(reduction="none", pos_weight=10.0)
B, T, L = 4, 200, 64
P = {k: v.to(dev) for k, v in synth.init_params(cfg, 0).items()}
batch = synth.make_batch(cfg, B, T, L, 100)
_, inp, inp_len, mels, mel_len, _, spk, stop = [x.to(dev) if hasattr(x, "to") else x for x in batch]
masks = synth.make_masks(cfg, B, T, L, 200)
md = {k: ([x.to(dev) for x in v] if isinstance(v, list) else v.to(dev)) for k, v in masks.items()}
names = list(P.keys())


def one_pass():
    Pl = {k: v.detach().requires_grad_(True) for k, v in P.items()}
    stats = {k: v.to(dev) for k, v in OM.fresh_bn_stats(Pl, cfg).items()}
    out = OM.forward(Pl, cfg, inp, inp_len.cpu(), mels, mel_len.cpu(), spk, md, stats, True, None)
    loss = OM.loss_fn(out, (mels, stop), mel_len.cpu(), **crit)
    torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    return float(loss)


for _ in range(2):
    one_pass()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    one_pass()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / reps * 1e3
print(json.dumps({"baseline": "torch eager (oracle restatement) on B200, fp32", "config": "1: fwd+bwd pass B4 T200 L64",
                  "ms_per_pass": ms, "mel_frames_per_s": B * T / ms * 1e3}))

# inference, B=32, L=64, 100 free-running steps
Bi, steps = 32, 100
cfg5 = pkg.default_params()
cfg5["max_decoder_steps"] = steps
cfg5["decoder_no_early_stopping"] = True
lens = torch.arange(64, 64 - Bi, -1)
tok = torch.randint(1, 123, (Bi, 64))
for b in range(Bi):
    tok[b, lens[b]:] = 0
spk5 = torch.randn(Bi, cfg5["speaker_embedding_dim"]).to(dev)
pm = synth.make_infer_masks(cfg5, Bi, steps, 5)
pm = pm.to(dev) if hasattr(pm, "to") else [x.to(dev) for x in pm]
stats = {k: v.to(dev) for k, v in OM.fresh_bn_stats(P, cfg5).items()}
with torch.no_grad():
    OM.infer(P, cfg5, tok.to(dev), lens, spk5, pm, stats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    OM.infer(P, cfg5, tok.to(dev), lens, spk5, pm, stats)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(json.dumps({"baseline": "torch eager (oracle restatement) on B200, fp32", "config": f"5: infer B32 L64 {steps} steps",
                  "us_per_step": dt / steps * 1e6, "mel_frames_per_s": Bi * steps / dt}))
