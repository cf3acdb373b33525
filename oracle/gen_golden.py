"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden.py            # needs /root/reference

For every case the reference's own ``Tacotron2NV`` / ``Tacotron2Loss``
(msa_tts/models/tacotron2nv.py, .../tacotron2nv_loss.py) are imported, loaded
with the synthetic weights, and run with ``torch.nn.functional.dropout`` replaced
by a queue of injected keep-masks (call order: SURVEY.md 8c).  The fixture holds
the reference's outputs; this script also asserts that the oracle restatement
(oracle/model.py) reproduces them, which is what pins the oracle.
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import msa_tts_b200 as pkg                                    # noqa: E402
from msa_tts_b200 import synth                                # noqa: E402
from oracle import model as OM                                # noqa: E402
from oracle.gen_cases import CASES, INFER_CASES, infer_stats, speaker_input  # noqa: E402

from msa_tts.models.tacotron2nv import Tacotron2NV            # noqa: E402  (reference)
from msa_tts.models.modules_tacotron2nv.tacotron2nv_loss import Tacotron2Loss  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class DropoutQueue:
    """Replaces F.dropout: pops the next injected keep-mask, checks its shape and p."""

    def __init__(self, seq):
        self.seq = list(seq)
        self.i = 0

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training:
            return x
        keep, p_exp = self.seq[self.i]
        self.i += 1
        assert tuple(keep.shape) == tuple(x.shape), (self.i, keep.shape, x.shape)
        assert abs(p - p_exp) < 1e-12, (self.i, p, p_exp)
        return x * keep * (1.0 / (1.0 - p))


def train_mask_sequence(cfg, masks, T):
    seq = [(m, 0.5) for m in masks["enc"]] + [(m, 0.5) for m in masks["prenet"]]
    for t in range(T):
        seq += [(masks["attn_h"][t], cfg["p_attention_dropout"]), (masks["dec_h"][t], cfg["p_decoder_dropout"])]
    seq += [(m, 0.5) for m in masks["post"]]
    return seq


def build_ref(cfg, P, stats=None):
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m = Tacotron2NV(copy.deepcopy(cfg))
    sd = m.state_dict()
    for k, v in P.items():
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        sd[k] = v.clone()
    if stats is not None:
        for k, v in stats.items():
            sd[k] = v.clone()
    m.load_state_dict(sd)
    assert [n for n, _ in m.named_parameters()] == OM.param_names(cfg) == list(P.keys())
    return m


def run_ref_train(cfg, P, batch, masks, crit):
    _, inp, inp_len, mels, mel_len, _, _, stop = batch
    spk = speaker_input(cfg, batch)
    m = build_ref(cfg, P)
    m.train()
    orig = torch.nn.functional.dropout
    q = DropoutQueue(train_mask_sequence(cfg, masks, mels.shape[2]))
    torch.nn.functional.dropout = q
    try:
        out = m(inputs=inp, input_lengths=inp_len, melspecs=mels, melspec_lengths=mel_len, speaker_vecs=spk)
    finally:
        torch.nn.functional.dropout = orig
    assert q.i == len(q.seq)
    loss = Tacotron2Loss(cfg["n_frames_per_step"], crit["reduction"], crit["pos_weight"], "cpu")(out, (mels, stop), mel_len)
    loss.backward()
    grads = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for n, p in m.named_parameters()}
    stats = {k: v.detach().clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
    return [o.detach() for o in out], loss.detach(), grads, stats


def run_oracle_train(cfg, P, batch, masks, crit, dtype=torch.float32):
    _, inp, inp_len, mels, mel_len, _, _, stop = batch
    spk = speaker_input(cfg, batch)
    spk = spk.to(dtype) if spk.dtype.is_floating_point else spk
    Pl = {k: v.to(dtype).clone().requires_grad_(True) for k, v in P.items()}
    stats = OM.fresh_bn_stats(Pl, cfg)
    masks_d = {k: ([x.to(dtype) for x in v] if isinstance(v, list) else v.to(dtype)) for k, v in masks.items()}
    out = OM.forward(Pl, cfg, inp, inp_len, mels.to(dtype), mel_len, spk, masks_d, stats, True)
    loss = OM.loss_fn(out, (mels.to(dtype), stop.to(dtype)), mel_len, **crit)
    names = list(P.keys())
    g = torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    grads = {n: (torch.zeros_like(Pl[n]) if x is None else x) for n, x in zip(names, g)}
    return [o.detach() for o in out], loss.detach(), grads, stats


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def check_train(tag, ref, ora, tol=2e-5):
    (ro, rl, rg, rs), (oo, ol, og, os_) = ref, ora
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in rg.values())))
    worst = 0.0
    for n, a, b in zip(("mel", "mel_post", "gate", "align"), oo, ro):
        worst = max(worst, rel(a, b))
    assert abs(float(ol) - float(rl)) <= tol * abs(float(rl)), (tag, float(ol), float(rl))
    gworst = max(float((og[n].double() - rg[n].double()).norm()) / gn for n in rg)
    sworst = max(rel(os_[k].float(), rs[k].float()) for k in rs if "running" in k)
    print(f"[{tag}] oracle vs reference: outputs {worst:.2e}  loss {abs(float(ol)-float(rl))/abs(float(rl)):.2e}  "
          f"grads(global-rel) {gworst:.2e}  bn-stats {sworst:.2e}")
    assert worst < tol and gworst < tol and sworst < tol, tag


def save_train_case(name, full=True):
    cfg, seed, (B, T, L), crit = CASES[name]()
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    ref = run_ref_train(cfg, P, batch, masks, crit)
    ora = run_oracle_train(cfg, P, batch, masks, crit)
    check_train(name, ref, ora)
    out, loss, grads, stats = ref
    d = {"meta_seed": np.int64(seed), "meta_BTL": np.array([B, T, L]), "loss": loss.numpy(),
         "mel": out[0].numpy(), "mel_post": out[1].numpy(), "gate": out[2].numpy()}
    if full:
        d["align"] = out[3].numpy()
        for k, v in grads.items():
            d["grad/" + k] = v.numpy()
        for k, v in stats.items():
            d["stat/" + k] = v.numpy()
    else:
        # compact: per-tensor grad norms + a strided sample of every gradient
        d["align_sample"] = out[3][:, ::8].numpy()
        d["grad_norms"] = np.array([float(g.double().norm()) for g in grads.values()])
        for k, v in grads.items():
            flat = v.flatten()
            d["gsample/" + k] = flat[:: max(1, flat.numel() // 64)][:64].numpy()
        d["param_checksum"] = np.array([float(v.double().sum()) for v in P.values()])
        for k, v in stats.items():
            if "running" in k:
                d["stat/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)


def run_ref_infer(cfg, P, stats, inp, inp_len, spk, pm):
    import io, contextlib
    m = build_ref(cfg, P, stats)
    m.eval()
    seq = []
    for s in range(pm.shape[0]):
        seq += [(pm[s, 0], 0.5), (pm[s, 1], 0.5)]
    q = DropoutQueue(seq)
    orig = torch.nn.functional.dropout
    torch.nn.functional.dropout = q
    try:
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            return m.infer(inp, inp_len, spk)
    finally:
        torch.nn.functional.dropout = orig


def save_infer_case(name):
    cfg, seed, (B, L), steps = INFER_CASES[name]()
    P = synth.init_params(cfg, seed)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, seed + 100)
    stats = infer_stats(P, cfg, seed)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    mel_post, mel_lengths, align = run_ref_infer(cfg, P, stats, inp, inp_len, spk, pm)
    o_post, o_len, o_al = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert o_post.shape == mel_post.shape, (o_post.shape, mel_post.shape)
    assert torch.equal(o_len, mel_lengths), (o_len, mel_lengths)
    e = max(rel(o_post, mel_post), rel(o_al, align))
    print(f"[{name}] oracle vs reference infer: steps {mel_post.shape[2]}  lengths {mel_lengths.tolist()}  err {e:.2e}")
    assert e < 2e-5
    d = {"meta_seed": np.int64(seed), "meta_BL": np.array([B, L]), "steps": np.int64(mel_post.shape[2]),
         "mel_post": mel_post.numpy(), "mel_lengths": mel_lengths.numpy(), "align": align.numpy()}
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]          # optional: names of the cases to (re)generate
    for name in CASES:
        if not only or name in only:
            save_train_case(name, full=(name != "default_train_b4_t200"))
    for name in INFER_CASES:
        if not only or name in only:
            save_infer_case(name)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
