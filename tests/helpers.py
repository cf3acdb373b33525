"""Shared helpers for the parity tests: run the oracle and the CUDA path on the same seeded inputs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from msa_tts_b200 import synth  # noqa: E402
from oracle import model as OM  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_pass(cfg, P, batch, masks, crit, dtype=torch.float32):
    """Oracle forward + loss + autograd grads; returns (outputs, loss, grads dict, bn stats, intermediates)."""
    _, inp, inp_len, mels, mel_len, spk_ids, spk, stop = batch
    if cfg["speaker_emb_type"] == "learnable_lookup":
        spk = spk_ids
    Pl = {k: v.to(dtype).clone().requires_grad_(True) for k, v in P.items()}
    stats = OM.fresh_bn_stats(Pl, cfg)
    md = {k: ([x.to(dtype) for x in v] if isinstance(v, list) else v.to(dtype)) for k, v in masks.items()}
    inter = {}
    spk_in = spk.to(dtype) if spk.dtype.is_floating_point else spk
    out = OM.forward(Pl, cfg, inp, inp_len, mels.to(dtype), mel_len, spk_in, md, stats, True, inter)
    loss = OM.loss_fn(out, (mels.to(dtype), stop.to(dtype)), mel_len, **crit)
    names = list(P.keys())
    g = torch.autograd.grad(loss, [Pl[n] for n in names], allow_unused=True)
    grads = {n: (torch.zeros_like(Pl[n]) if x is None else x.detach()) for n, x in zip(names, g)}
    return [o.detach() for o in out], loss.detach(), grads, stats, {k: v.detach() for k, v in inter.items()}


def cuda_pass(eng, cfg, P, batch, masks, backward=True):
    """The CUDA path through the C ABI; returns (outputs, loss, grads dict, bn dict)."""
    from msa_tts_b200.engine import batch_to_device
    B, L = batch[1].shape
    T = batch[3].shape[2]
    flat = eng.flat_from_dict(P)
    bn = eng.new_bn_stats()
    bd = batch_to_device(batch, eng.device, cfg["speaker_emb_type"])
    mflat = eng.pack_masks(masks, B, T, L)
    out, loss = eng.forward(flat, bn, bd, mflat)
    grads = None
    if backward:
        gflat = eng.new_flat(float("nan"))
        # padding floats between tensors are never written by the library: zero them so that NaN means "not written"
        gflat.zero_()
        for n in eng.layout.names():
            o = eng.layout.offsets[n]
            gflat[o:o + eng.layout.numel(n)] = float("nan")
        eng.backward(flat, gflat, accumulate=False, scale=1.0)
        grads = {k: v.clone() for k, v in eng.dict_from_flat(gflat).items()}
    torch.cuda.synchronize()
    eng.check_abort()      # a persistent kernel that timed out while polling invalidates everything
    return out, loss, grads, eng.bn_dict(bn)


def intermediates_report(eng, cfg, inter, B, T, L):
    """Per-stage relative errors of the CUDA intermediates against the oracle's (forward order)."""
    C = cfg["encoder_embedding_dim"]
    rows = []
    n_enc = cfg["encoder_n_convolutions"]
    ex = eng.get_buffer("enc_x").view(n_enc + 1, B, L, C)
    for i in range(n_enc):
        rows.append((f"enc_conv{i}", rel(ex[i + 1], inter[f"enc_conv{i}"].permute(0, 2, 1))))
    Hh = C // 2
    eh = eng.get_buffer("enc_h").view(2, L, B, Hh)
    enc = torch.cat([eh[0], eh[1]], dim=-1).permute(1, 0, 2)
    rows.append(("enc_out", rel(enc, inter["enc_out"])))
    E = inter["memory"].shape[2]
    rows.append(("memory", rel(eng.get_buffer("memory").view(B, L, E), inter["memory"])))
    Pd = cfg["prenet_dim"]
    rows.append(("prenet_out", rel(eng.get_buffer("xpre").view(T + 1, B, Pd), inter["prenet_out"])))
    A = cfg["attention_params"]["attention_dim"]
    rows.append(("pm", rel(eng.get_buffer("pm").view(B, L, A), inter["pm"])))
    Ha, Hd = inter["ha"].shape[2], inter["hd"].shape[2]
    rows.append(("ha", rel(eng.get_buffer("ha").view(T, B, Ha), inter["ha"])))
    rows.append(("ctx", rel(eng.get_buffer("ctx").view(T, B, E), inter["ctx"])))
    rows.append(("hd", rel(eng.get_buffer("hd").view(T, B, Hd), inter["hd"])))
    n_post = cfg["postnet_n_convolutions"]
    M, Cp = cfg["n_mel_channels"], cfg["postnet_embedding_dim"]
    Cmax = max(M, Cp)
    px = eng.get_buffer("post_x").view(n_post + 1, B * T * Cmax)
    for i in range(n_post):
        co = M if i == n_post - 1 else Cp
        rows.append((f"post_conv{i}", rel(px[i + 1][:B * T * co].view(B, T, co), inter[f"post_conv{i}"].permute(0, 2, 1))))
    return rows
