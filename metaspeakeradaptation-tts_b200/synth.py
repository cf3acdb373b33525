"""Seeded synthetic speaker-task batches, dropout keep-masks and random-init weights.

There is no corpus and no checkpoint in the reference repo,
so every test and benchmark runs on the synthetic inputs specified in SURVEY.md
section 8(d).  The batch tuple has the layout ``MetaCollator.__call__`` produces
(msa_tts/dataloaders/dataloader_meta.py:133-179) and ``_unpack_batch`` consumes
(msa_tts/metatrainer.py:95-117).  Everything here is CPU torch; nothing touches
the GPU or the oracle.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from .config import memory_dim, rnn_dims, speaker_dim
from .layout import param_shapes


def make_batch(cfg: dict, B: int, T: int, L: int, seed: int, spk_vec: Optional[torch.Tensor] = None,
               ragged: bool = True) -> tuple:
    """One (item_ids, transcripts, trans_lengths, melspecs, melspec_lengths, speaker_ids, spk_embs, stop_targets)."""
    g = torch.Generator().manual_seed(seed)
    n_mel = cfg["n_mel_channels"]
    if ragged and B > 1:
        # lengths sorted descending, lengths[0] == L (dataloader_meta.py:145-147)
        tl = [max(1, L - (i * L) // (2 * B)) for i in range(B)]
        ml = [max(2, T - (i * T) // (2 * B)) for i in range(B)]
    else:
        tl, ml = [L] * B, [T] * B
    trans_lengths = torch.tensor(tl, dtype=torch.int64)
    mel_lengths = torch.tensor(ml, dtype=torch.int64)
    transcripts = torch.randint(1, cfg["n_symbols"], (B, L), generator=g, dtype=torch.int64)
    transcripts[torch.arange(L)[None, :] >= trans_lengths[:, None]] = 0       # pad id 0
    mels = (torch.randn(B, n_mel, T, generator=g) * 2.0 - 4.0).clamp_(-10.0, 2.0)
    tmask = torch.arange(T)[None, :] >= mel_lengths[:, None]
    mels.masked_fill_(tmask[:, None, :], 0.0)                                  # dataloader_meta.py:217-221
    stop = torch.zeros(B, T)
    stop[torch.arange(B), mel_lengths - 1] = 1.0
    stop.masked_fill_(tmask, 1.0)                                              # dataloader_meta.py:161-162, 239-243
    if spk_vec is None:
        spk_vec = torch.randn(cfg["speaker_embedding_dim"], generator=g)
    spk_embs = spk_vec[None, :].expand(B, -1).contiguous()                     # dataloader_meta.py:108
    ns = max(cfg.get("num_speakers", 1), 1)
    if cfg["speaker_emb_type"] == "learnable_lookup":
        # rows of different speakers (with a repeat), so that the embedding gather / scatter-add is exercised
        speaker_ids = (seed + torch.arange(B, dtype=torch.int64) // 2 * 3) % ns
    else:
        speaker_ids = torch.full((B,), seed % ns, dtype=torch.int64)
    item_ids = [f"synth_{seed}_{i}" for i in range(B)]
    return (item_ids, transcripts, trans_lengths, mels, mel_lengths, speaker_ids, spk_embs, stop)


def make_task(cfg: dict, B: int, T: int, L: int, seed: int) -> Dict[str, tuple]:
    """{"train": batch, "test": batch} for one speaker (dataloader_meta.py:70-111)."""
    g = torch.Generator().manual_seed(seed)
    spk = torch.randn(cfg["speaker_embedding_dim"], generator=g)
    return {"train": make_batch(cfg, B, T, L, seed * 2 + 1, spk),
            "test": make_batch(cfg, B, T, L, seed * 2 + 2, spk)}


def mask_shapes(cfg: dict, B: int, T: int, L: int) -> Dict[str, list]:
    """Reference-layout shapes of every F.dropout call of one training forward."""
    C = cfg["encoder_embedding_dim"]
    Ha, Hd = rnn_dims(cfg)
    n = cfg["postnet_n_convolutions"]
    return {
        "enc": [(B, C, L)] * cfg["encoder_n_convolutions"],
        "prenet": [(T + 1, B, cfg["prenet_dim"])] * 2,
        "attn_h": [(T, B, Ha)],
        "dec_h": [(T, B, Hd)],
        "post": [(B, cfg["n_mel_channels"] if i == n - 1 else cfg["postnet_embedding_dim"], T) for i in range(n)],
    }


def make_masks(cfg: dict, B: int, T: int, L: int, seed: int) -> dict:
    """Bernoulli(1-p) keep-masks (float 0/1) in the reference's layouts."""
    g = torch.Generator().manual_seed(seed)
    shp = mask_shapes(cfg, B, T, L)
    pa, pd = cfg["p_attention_dropout"], cfg["p_decoder_dropout"]

    def bern(s, p):
        return (torch.rand(*s, generator=g) >= p).float()

    return {
        "enc": [bern(s, 0.5) for s in shp["enc"]],
        "prenet": [bern(s, 0.5) for s in shp["prenet"]],
        "attn_h": bern(shp["attn_h"][0], pa),
        "dec_h": bern(shp["dec_h"][0], pd),
        "post": [bern(s, 0.5) for s in shp["post"]],
    }


def make_infer_masks(cfg: dict, B: int, steps: int, seed: int) -> torch.Tensor:
    """[steps, 2, B, prenet_dim] keep-masks for the always-on prenet dropout (decoder.py:19,366)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(steps, 2, B, cfg["prenet_dim"], generator=g) >= 0.5).float()


def _xavier(shape, gain, g):
    # torch.nn.init.xavier_uniform_: fan_in/out include the receptive field
    rf = 1
    for d in shape[2:]:
        rf *= d
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    a = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def init_params(cfg: dict, seed: int = 0) -> "Dict[str, torch.Tensor]":
    """Random-init weights with the reference's distributions (same families and
    scales as tacotron2nv.py:19-22, modules.py:8-37, torch LSTM/LSTMCell/BatchNorm
    defaults); NOT the reference's RNG stream -- both sides load this dict."""
    g = torch.Generator().manual_seed(seed)
    shapes = param_shapes(cfg)
    P: Dict[str, torch.Tensor] = {}
    gains = {"relu": math.sqrt(2.0), "tanh": 5.0 / 3.0, "linear": 1.0, "sigmoid": 1.0}
    Ha, Hd = rnn_dims(cfg)
    n_post = cfg["postnet_n_convolutions"]
    for name, shp in shapes.items():
        if name == "embedding.weight":
            std = math.sqrt(2.0 / (cfg["n_symbols"] + cfg["symbols_embedding_dim"]))
            val = math.sqrt(3.0) * std
            t = (torch.rand(*shp, generator=g) * 2 - 1) * val
        elif name.endswith(".1.weight"):                       # BatchNorm gamma
            t = torch.ones(*shp) + 0.1 * (torch.rand(*shp, generator=g) - 0.5)
        elif name.endswith(".1.bias"):                         # BatchNorm beta
            t = 0.1 * (torch.rand(*shp, generator=g) - 0.5)
        elif ".conv.weight" in name:
            if name.startswith("encoder"):
                gain = gains["relu"]
            else:
                i = int(name.split(".")[2])
                gain = gains["linear"] if i == n_post - 1 else gains["tanh"]
            t = _xavier(shp, gain, g)
        elif ".conv.bias" in name:
            fan_in = shapes[name.replace("bias", "weight")][1] * shapes[name.replace("bias", "weight")][2]
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(*shp, generator=g) * 2 - 1) * b
        elif name.startswith("encoder.lstm"):
            b = 1.0 / math.sqrt(cfg["encoder_embedding_dim"] // 2)
            t = (torch.rand(*shp, generator=g) * 2 - 1) * b
        elif name.startswith("decoder.attention_rnn"):
            t = (torch.rand(*shp, generator=g) * 2 - 1) * (1.0 / math.sqrt(Ha))
        elif name.startswith("decoder.decoder_rnn"):
            t = (torch.rand(*shp, generator=g) * 2 - 1) * (1.0 / math.sqrt(Hd))
        elif "location_conv1d" in name:
            fan_in = shp[1] * shp[2]
            t = (torch.rand(*shp, generator=g) * 2 - 1) * (1.0 / math.sqrt(fan_in))
        elif name.endswith("linear_layer.weight") or name.endswith("ta.weight") or name == "speaker_lin.weight":
            gain = gains["tanh"] if ("query_layer" in name or "inputs_layer" in name or "location_dense" in name) else 1.0
            t = _xavier(shp, gain, g)
        elif name == "speaker_embedder.weight":
            t = torch.randn(*shp, generator=g)
        else:                                                  # remaining biases
            t = (torch.rand(*shp, generator=g) * 2 - 1) * 0.05
        P[name] = t.float().contiguous()
    return P
