"""Flat parameter-buffer layout.

All parameter tensors live in ONE contiguous fp32 buffer in
``model.parameters()`` order so that the inner SGD step,
the meta-gradient accumulation, the Reptile delta, clipping and the outer update
are single fused launches (msa_tts/maml.py:94-105, reptile.py:73-89,
utils/grad_utils.py:8-31 operate tensor by tensor).  Each tensor starts on a
128-byte boundary; the padding floats are zero in every buffer that shares the
layout (params, grads, optimizer state, Fisher, means) and therefore do not
change norms or sums.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Tuple

from .config import memory_dim, rnn_dims, speaker_dim

ALIGN = 32  # floats


def param_shapes(cfg: dict) -> "OrderedDict[str, Tuple[int, ...]]":
    C = cfg["encoder_embedding_dim"]
    assert cfg["symbols_embedding_dim"] == C, "the reference feeds the embedding straight into the encoder convs"
    K = cfg["encoder_kernel_size"]
    E = memory_dim(cfg)
    Ha, Hd = rnn_dims(cfg)
    Pd = cfg["prenet_dim"]
    M = cfg["n_mel_channels"] * cfg["n_frames_per_step"]
    ap = cfg["attention_params"]
    A, F_, Kl = ap["attention_dim"], ap["attention_location_n_filters"], ap["attention_location_kernel_size"]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["embedding.weight"] = (cfg["n_symbols"], C)
    for i in range(cfg["encoder_n_convolutions"]):
        p = f"encoder.convolutions.{i}"
        s[p + ".0.conv.weight"] = (C, C, K)
        s[p + ".0.conv.bias"] = (C,)
        s[p + ".1.weight"] = (C,)
        s[p + ".1.bias"] = (C,)
    Hh = C // 2
    for sfx in ("", "_reverse"):
        s[f"encoder.lstm.weight_ih_l0{sfx}"] = (4 * Hh, C)
        s[f"encoder.lstm.weight_hh_l0{sfx}"] = (4 * Hh, Hh)
        s[f"encoder.lstm.bias_ih_l0{sfx}"] = (4 * Hh,)
        s[f"encoder.lstm.bias_hh_l0{sfx}"] = (4 * Hh,)
    if cfg["speaker_emb_type"] == "learnable_lookup":
        s["speaker_embedder.weight"] = (cfg["num_speakers"], cfg["speaker_embedding_dim"])
    elif cfg["speaker_emb_type"] == "static+linear":
        s["speaker_lin.weight"] = (cfg["speaker_embedding_dim_lin"], cfg["speaker_embedding_dim"])
        s["speaker_lin.bias"] = (cfg["speaker_embedding_dim_lin"],)
    s["decoder.prenet.layers.0.linear_layer.weight"] = (Pd, M)
    s["decoder.prenet.layers.1.linear_layer.weight"] = (Pd, Pd)
    s["decoder.attention_rnn.weight_ih"] = (4 * Ha, Pd + E)
    s["decoder.attention_rnn.weight_hh"] = (4 * Ha, Ha)
    s["decoder.attention_rnn.bias_ih"] = (4 * Ha,)
    s["decoder.attention_rnn.bias_hh"] = (4 * Ha,)
    a = "decoder.attention_layer."
    s[a + "query_layer.linear_layer.weight"] = (A, Ha)
    s[a + "inputs_layer.linear_layer.weight"] = (A, E)
    s[a + "v.linear_layer.weight"] = (1, A)
    s[a + "v.linear_layer.bias"] = (1,)
    if ap["trans_agent"]:
        s[a + "ta.weight"] = (1, Ha + E)
        s[a + "ta.bias"] = (1,)
    s[a + "location_layer.location_conv1d.weight"] = (F_, 2, Kl)
    s[a + "location_layer.location_dense.linear_layer.weight"] = (A, F_)
    s["decoder.decoder_rnn.weight_ih"] = (4 * Hd, Ha + E)
    s["decoder.decoder_rnn.weight_hh"] = (4 * Hd, Hd)
    s["decoder.decoder_rnn.bias_ih"] = (4 * Hd,)
    s["decoder.decoder_rnn.bias_hh"] = (4 * Hd,)
    s["decoder.linear_projection.linear_layer.weight"] = (M, Hd + E)
    s["decoder.linear_projection.linear_layer.bias"] = (M,)
    s["decoder.gate_layer.linear_layer.weight"] = (1, Hd + E)
    s["decoder.gate_layer.linear_layer.bias"] = (1,)
    Cp, Kp, n = cfg["postnet_embedding_dim"], cfg["postnet_kernel_size"], cfg["postnet_n_convolutions"]
    for i in range(n):
        p = f"postnet.convolutions.{i}"
        cin = cfg["n_mel_channels"] if i == 0 else Cp
        cout = cfg["n_mel_channels"] if i == n - 1 else Cp
        s[p + ".0.conv.weight"] = (cout, cin, Kp)
        s[p + ".0.conv.bias"] = (cout,)
        s[p + ".1.weight"] = (cout,)
        s[p + ".1.bias"] = (cout,)
    return s


def bn_layer_names(cfg: dict) -> List[str]:
    return [f"encoder.convolutions.{i}.1" for i in range(cfg["encoder_n_convolutions"])] + \
           [f"postnet.convolutions.{i}.1" for i in range(cfg["postnet_n_convolutions"])]


def bn_channels(cfg: dict) -> List[int]:
    n = cfg["postnet_n_convolutions"]
    return [cfg["encoder_embedding_dim"]] * cfg["encoder_n_convolutions"] + \
           [cfg["n_mel_channels"] if i == n - 1 else cfg["postnet_embedding_dim"] for i in range(n)]


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= d
    return n


class FlatLayout:
    """name -> (offset, shape) in floats, plus the BN running-stat layout."""

    def __init__(self, cfg: dict):
        self.cfg = cfg
        self.shapes = param_shapes(cfg)
        self.offsets: "OrderedDict[str, int]" = OrderedDict()
        off = 0
        for name, shp in self.shapes.items():
            self.offsets[name] = off
            off += (_numel(shp) + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self.n_params = sum(_numel(s) for s in self.shapes.values())
        # BN running stats: [mean(C), var(C)] per layer, one flat fp32 buffer (private per task)
        self.bn_names = bn_layer_names(cfg)
        self.bn_ch = bn_channels(cfg)
        self.bn_offsets = []
        o = 0
        for c in self.bn_ch:
            self.bn_offsets.append(o)
            o += 2 * ((c + ALIGN - 1) // ALIGN * ALIGN)
        self.bn_total = o

    def names(self) -> List[str]:
        return list(self.shapes.keys())

    def numel(self, name: str) -> int:
        return _numel(self.shapes[name])

    def offset_table(self) -> List[int]:
        """Offsets in the fixed order the C ABI expects (include/msa_b200.h, MSA_P_*)."""
        return [self.offsets[n] for n in self.names()]
