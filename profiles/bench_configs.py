"""The other BASELINE.json configs on one B200 (records for profiles/, not the driver's bench line):
   config 1: one forward+backward pass, B=4, T=200, L=64           (s/pass, mel-frames/s)
   config 3: batched Reptile meta-step, 16 tasks x 5 inner steps    (meta-steps/s)
   config 4: one EWC training step with the fused penalty+update     (s/step)
   config 5: free-running inference, B=32, L=64, 1000 decoder steps  (mel-frames/s)
    python profiles/bench_configs.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine, batch_to_device
from msa_tts_b200.reptile import Reptile

dev = torch.device("cuda:0")
cfg = pkg.default_params()
B, T, L = 4, 200, 64


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- config 1 ----
eng = Engine(cfg, dev, gemm_tf32=1)
P = synth.init_params(cfg, 0)
flat, g, bn = eng.flat_from_dict(P), eng.new_flat(), eng.new_bn_stats()
bd = batch_to_device(synth.make_batch(cfg, B, T, L, 100), dev)
masks = eng.generate_masks(B, T, L, 7)


def one_pass():
    eng.forward(flat, bn, bd, masks, outputs=False)
    eng.backward(flat, g)


ms = timed(one_pass, 10)
eng.check_abort()
print(json.dumps({"config": "1: fwd+bwd pass B4 T200 L64", "ms_per_pass": ms, "mel_frames_per_s": B * T / ms * 1e3}))

# ---- config 4: EWC step (forward, backward, fused penalty + SGD) ----
mu, fisher = flat.clone(), torch.rand_like(flat) * 1e-3


def ewc_step():
    eng.forward(flat, bn, bd, masks, outputs=False)
    eng.backward(flat, g)
    eng.ewc_sgd_step(flat, g, mu, fisher, 1e-4, 100.0)


ms = timed(ewc_step, 10)
print(json.dumps({"config": "4: EWC step (fwd+bwd + fused penalty/update)", "ms_per_step": ms}))
ms_u = timed(lambda: eng.ewc_sgd_step(flat, g, mu, fisher, 1e-9, 100.0), 20)
n = eng.layout.total
print(json.dumps({"config": "4: fused EWC penalty+update kernel alone", "ms": ms_u, "GBps": 20.0 * n / ms_u / 1e6,
                  "frac_of_measured_hbm_peak": 20.0 * n / ms_u / 1e6 / bench.measured_peak()[0]}))

# ---- config 5: inference B=32, 1000 steps, no early stopping ----
cfg5 = pkg.default_params()
cfg5["max_decoder_steps"] = 1000
cfg5["decoder_no_early_stopping"] = True
eng5 = Engine(cfg5, dev)
B5 = 32
lens = torch.arange(64, 64 - B5, -1)
inp = torch.randint(1, 123, (B5, 64))
for b in range(B5):
    inp[b, lens[b]:] = 0
spk = torch.randn(B5, cfg5["speaker_embedding_dim"])
pm = synth.make_infer_masks(cfg5, B5, 1000, 5)
flat5, bn5 = eng5.flat_from_dict(P), eng5.new_bn_stats()
t0 = time.perf_counter()
out = eng5.infer(flat5, bn5, inp, lens, spk, pm, max_steps=1000)
torch.cuda.synchronize()
t1 = time.perf_counter()
out = eng5.infer(flat5, bn5, inp, lens, spk, pm, max_steps=1000)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(json.dumps({"config": "5: infer B32 L64 1000 steps", "s_first_call": t1 - t0, "s": t2 - t1, "steps": int(out[0].shape[2]),
                  "mel_frames_per_s": B5 * out[0].shape[2] / (t2 - t1)}))
del eng5

# ---- config 3: batched Reptile, 16 tasks x 5 inner steps ----
bench.N_TASKS = 16
params = bench.trainer_params(1)
params["n_inner_train"] = 5
params["reptile_sequential"] = False
tr = Reptile(**params)
items = bench.make_tasks(cfg, pinned=False)
items = {s: {k: tuple(x.to(dev) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in items.items()}
ms = timed(lambda: tr._metatrain_step(items), 2, warm=1)
print(json.dumps({"config": "3: Reptile meta-step 16 tasks x 5 inner steps (+ test forward)", "ms_per_meta_step": ms,
                  "meta_steps_per_s": 1e3 / ms}))
