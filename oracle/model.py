"""Functional CPU restatement of ``Tacotron2NV`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
/root/reference/msa_tts/).  Parameters are passed as a dict keyed by the
reference's ``state_dict`` names; dropout is replaced by
explicit keep-masks in the reference's own tensor
layouts and call order.

Mask dict (values are 0/1 float tensors, 1 = keep):
    enc    : list[n_enc_convs] of [B, C_enc, L]      (encoder.py:36-37)
    prenet : list[2] of [T+1, B, prenet_dim]         (decoder.py:19, 293)
    attn_h : [T, B, H_attn]                          (decoder.py:256)
    dec_h  : [T, B, H_dec]                           (decoder.py:265)
    post   : list[n_post] of [B, C, T]               (decoder.py:63-72)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _drop(x, keep, p):
    """F.dropout(x, p, training=True) with an injected keep-mask."""
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


def batch_norm(x, weight, bias, stats: Optional[dict], key: str, train: bool):
    """nn.BatchNorm1d on [B, C, N] (encoder.py:26, decoder.py:39,50,60).

    train: normalise with biased batch statistics over (B, N) -- padded
    positions included (Q6) -- and update running stats with the unbiased
    variance, momentum 0.1.  eval: running statistics.
    """
    if train:
        n = x.shape[0] * x.shape[2]
        mean = x.mean(dim=(0, 2))
        var = x.var(dim=(0, 2), unbiased=False)
        if stats is not None:
            with torch.no_grad():
                rm = stats[key + ".running_mean"]
                rv = stats[key + ".running_var"]
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(rm.dtype))
                unb = var.detach() * (n / max(n - 1, 1))
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * unb.to(rv.dtype))
                stats[key + ".num_batches_tracked"] += 1
    else:
        mean = stats[key + ".running_mean"].to(x.dtype)
        var = stats[key + ".running_var"].to(x.dtype)
    xh = (x - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
    return xh * weight[None, :, None] + bias[None, :, None]


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell: gate order i, f, g, o (decoder.py:107-108, 135-137)."""
    z = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    i, f, g, o = z.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def encoder(P, cfg, inputs, input_lengths, masks, stats, train, inter=None):
    """tacotron2nv.py:88 + encoder.py:35-52 (forward) / 55-71 (infer)."""
    emb = P["embedding.weight"][inputs]                       # [B, L, C]
    x = emb.transpose(1, 2)                                   # [B, C, L]
    pad = (cfg["encoder_kernel_size"] - 1) // 2
    for i in range(cfg["encoder_n_convolutions"]):
        pre = f"encoder.convolutions.{i}"
        y = F.conv1d(x, P[pre + ".0.conv.weight"], P[pre + ".0.conv.bias"], padding=pad)
        y = batch_norm(y, P[pre + ".1.weight"], P[pre + ".1.bias"], stats, pre + ".1", train)
        keep = masks["enc"][i] if (train and masks is not None) else None
        x = _drop(F.relu(y), keep, 0.5) if train else F.relu(y)
        if inter is not None:
            inter[f"enc_conv{i}"] = x
    x = x.transpose(1, 2)                                     # [B, L, C]
    B, L, C = x.shape
    Hh = C // 2
    outs = []
    for sfx, rev in (("", False), ("_reverse", True)):
        w_ih = P[f"encoder.lstm.weight_ih_l0{sfx}"]
        w_hh = P[f"encoder.lstm.weight_hh_l0{sfx}"]
        b_ih = P[f"encoder.lstm.bias_ih_l0{sfx}"]
        b_hh = P[f"encoder.lstm.bias_hh_l0{sfx}"]
        h = x.new_zeros(B, Hh)
        c = x.new_zeros(B, Hh)
        out = [None] * L
        order = range(L - 1, -1, -1) if rev else range(L)
        for t in order:
            # packed-sequence semantics: rows with t >= len are not stepped and output 0
            act = (t < input_lengths.to(x.device)).to(x.dtype).unsqueeze(1)
            h2, c2 = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            h = act * h2 + (1 - act) * h
            c = act * c2 + (1 - act) * c
            out[t] = act * h2
        outs.append(torch.stack(out, dim=1))                  # [B, L, Hh]
    enc = torch.cat(outs, dim=-1)
    if inter is not None:
        inter["enc_out"] = enc
    return enc


def speaker_concat(P, cfg, enc, speaker_vecs):
    """tacotron2nv.py:104-111."""
    t = cfg["speaker_emb_type"]
    if t == "learnable_lookup":
        v = P["speaker_embedder.weight"][speaker_vecs]
    elif t == "static":
        v = speaker_vecs
    elif t == "static+linear":
        v = speaker_vecs @ P["speaker_lin.weight"].t() + P["speaker_lin.bias"]
    else:
        raise NotImplementedError(t)
    v = v.unsqueeze(1).expand(enc.shape[0], enc.shape[1], -1)
    return torch.cat([enc, v.to(enc.dtype)], dim=-1)


def prenet(P, x, keeps):
    """decoder.py:9-20 -- dropout p=0.5 always on (Q4)."""
    for i in range(2):
        w = P[f"decoder.prenet.layers.{i}.linear_layer.weight"]
        x = _drop(F.relu(x @ w.t()), None if keeps is None else keeps[i], 0.5)
    return x


class AttnState:
    """Module-attribute state of ForwardAttention (forward_attn.py:90-113)."""

    def __init__(self, P, cfg, memory):
        ap = cfg["attention_params"]
        B, L, _ = memory.shape
        self.prev = memory.new_zeros(B, L)
        self.cum = memory.new_zeros(B, L)
        self.pm = memory @ P["decoder.attention_layer.inputs_layer.linear_layer.weight"].t()
        if ap["forward_attn"]:
            self.alpha = torch.cat([memory.new_ones(B, 1), memory.new_zeros(B, L - 1) + 1e-7], dim=1)
            self.u = 0.5 * memory.new_ones(B, 1)
        if ap["windowing"]:
            self.win_idx, self.win_back, self.win_front = -1, 2, 6


def attention_step(P, cfg, st: AttnState, query, memory, train):
    """forward_attn.py:121-131 (energies) and 178-225 (forward)."""
    ap = cfg["attention_params"]
    pre = "decoder.attention_layer."
    ksz = ap["attention_location_kernel_size"]
    cat = torch.stack([st.prev, st.cum], dim=1)                               # [B, 2, L]
    q = (query @ P[pre + "query_layer.linear_layer.weight"].t()).unsqueeze(1)  # [B, 1, A]
    loc = F.conv1d(cat, P[pre + "location_layer.location_conv1d.weight"], padding=(ksz - 1) // 2)
    loc = loc.transpose(1, 2) @ P[pre + "location_layer.location_dense.linear_layer.weight"].t()
    e = torch.tanh(q + loc + st.pm) @ P[pre + "v.linear_layer.weight"].t() + P[pre + "v.linear_layer.bias"]
    e = e.squeeze(-1)                                                          # [B, L]
    # NOTE: the padding mask is never applied (forward_attn.py:192-194, Q2)
    if (not train) and ap["windowing"]:
        # forward_attn.py:139-152 -- argmax of batch row 0 drives the whole batch (Q16)
        e = e.clone()
        back, front = st.win_idx - st.win_back, st.win_idx + st.win_front
        if back > 0:
            e[:, :back] = -float("inf")
        if front < memory.shape[1]:
            e[:, front:] = -float("inf")
        if st.win_idx == -1:
            e[:, 0] = e.max()
        st.win_idx = int(torch.argmax(e, 1)[0].item())
    if ap["norm"] == "softmax":
        a = torch.softmax(e, dim=-1)
    elif ap["norm"] == "sigmoid":
        s = torch.sigmoid(e)
        a = s / s.sum(dim=1, keepdim=True)
    else:
        raise ValueError("Unknown value for attention norm type")
    st.cum = st.cum + a                                                        # pre-forward-attention a
    if ap["forward_attn"]:
        # forward_attn.py:154-176
        shifted = F.pad(st.alpha[:, :-1], (1, 0, 0, 0))
        alpha = ((1 - st.u) * st.alpha + st.u * shifted + 1e-8) * a
        if (not train) and ap["forward_attn_mask"]:
            alpha = alpha.clone()
            _, n = shifted.max(1)
            val, _ = alpha.max(1)
            for b in range(a.shape[0]):
                nb = int(n[b])
                alpha[b, nb + 3:] = 0
                alpha[b, :(nb - 1)] = 0      # Python slice semantics incl. negative stop (Q16)
                alpha[b, (nb - 2)] = 0.01 * val[b]
        a = alpha / alpha.sum(dim=1, keepdim=True)
        st.alpha = a
    ctx = torch.bmm(a.unsqueeze(1), memory).squeeze(1)
    st.prev = a
    if ap["forward_attn"] and ap["trans_agent"]:
        ta_in = torch.cat([ctx, query], dim=-1)
        st.u = torch.sigmoid(ta_in @ P[pre + "ta.weight"].t() + P[pre + "ta.bias"])
    return ctx, a


def decode_step(P, cfg, st, x_t, state, memory, keep_a, keep_d, train):
    """decoder.py:234-274."""
    ha, ca, hd, cd, ctx = state
    pa, pd = cfg["p_attention_dropout"], cfg["p_decoder_dropout"]
    ha, ca = lstm_cell(torch.cat([x_t, ctx], -1), ha, ca,
                       P["decoder.attention_rnn.weight_ih"], P["decoder.attention_rnn.weight_hh"],
                       P["decoder.attention_rnn.bias_ih"], P["decoder.attention_rnn.bias_hh"])
    if train:
        ha = _drop(ha, keep_a, pa)                     # dropped value IS the recurrent state (Q5)
    ctx, a = attention_step(P, cfg, st, ha, memory, train)
    hd, cd = lstm_cell(torch.cat([ha, ctx], -1), hd, cd,
                       P["decoder.decoder_rnn.weight_ih"], P["decoder.decoder_rnn.weight_hh"],
                       P["decoder.decoder_rnn.bias_ih"], P["decoder.decoder_rnn.bias_hh"])
    if train:
        hd = _drop(hd, keep_d, pd)
    hc = torch.cat([hd, ctx], dim=1)
    mel = hc @ P["decoder.linear_projection.linear_layer.weight"].t() + P["decoder.linear_projection.linear_layer.bias"]
    gate = hc @ P["decoder.gate_layer.linear_layer.weight"].t() + P["decoder.gate_layer.linear_layer.bias"]
    return mel, gate, a, (ha, ca, hd, cd, ctx)


def _zero_state(cfg, memory):
    B = memory.shape[0]
    # Q8: Tacotron2NV passes decoder_rnn_dim into the attention_rnn_dim slot and vice versa.
    Ha, Hd = cfg["decoder_rnn_dim"], cfg["attention_rnn_dim"]
    z = memory.new_zeros
    return (z(B, Ha), z(B, Ha), z(B, Hd), z(B, Hd), z(B, memory.shape[2]))


def decoder_forward(P, cfg, memory, mels, masks, train, inter=None):
    """decoder.py:277-331 (teacher forced; n_frames_per_step == 1, Q14)."""
    B, n_mel, T = mels.shape
    frames = torch.cat([mels.new_zeros(1, B, n_mel), mels.permute(2, 0, 1)], dim=0)   # [T+1, B, n_mel]
    X = prenet(P, frames, None if masks is None else masks["prenet"])
    st = AttnState(P, cfg, memory)
    state = _zero_state(cfg, memory)
    mel_o, gate_o, al_o = [], [], []
    ha_o, hd_o, ctx_o = [], [], []
    for t in range(T):
        ka = masks["attn_h"][t] if (masks is not None and train) else None
        kd = masks["dec_h"][t] if (masks is not None and train) else None
        mel, gate, a, state = decode_step(P, cfg, st, X[t], state, memory, ka, kd, train)
        mel_o.append(mel); gate_o.append(gate.squeeze(1)); al_o.append(a)
        if inter is not None:
            ha_o.append(state[0]); hd_o.append(state[2]); ctx_o.append(state[4])
    if inter is not None:
        inter["prenet_out"] = X
        inter["pm"] = st.pm
        inter["ha"] = torch.stack(ha_o); inter["hd"] = torch.stack(hd_o); inter["ctx"] = torch.stack(ctx_o)
    mel_o = torch.stack(mel_o).permute(1, 2, 0)               # [B, n_mel, T]
    gate_o = torch.stack(gate_o).transpose(0, 1)              # [B, T]
    al_o = torch.stack(al_o).transpose(0, 1)                  # [B, T, L]
    return mel_o, gate_o, al_o


def postnet(P, cfg, x, masks, stats, train, inter=None):
    """decoder.py:63-72."""
    n = cfg["postnet_n_convolutions"]
    pad = (cfg["postnet_kernel_size"] - 1) // 2
    for i in range(n):
        pre = f"postnet.convolutions.{i}"
        y = F.conv1d(x, P[pre + ".0.conv.weight"], P[pre + ".0.conv.bias"], padding=pad)
        y = batch_norm(y, P[pre + ".1.weight"], P[pre + ".1.bias"], stats, pre + ".1", train)
        if i < n - 1:
            y = torch.tanh(y)
        keep = masks["post"][i] if (train and masks is not None) else None
        x = _drop(y, keep, 0.5) if train else y
        if inter is not None:
            inter[f"post_conv{i}"] = x
    return x


def forward(P: Dict[str, torch.Tensor], cfg: dict, inputs, input_lengths, melspecs, melspec_lengths,
            speaker_vecs, masks: Optional[dict], stats: Optional[dict] = None, train: bool = True,
            inter: Optional[dict] = None) -> List[torch.Tensor]:
    """Tacotron2NV.forward, tacotron2nv.py:81-127 -> [mel, mel_post, gate, align]."""
    assert not cfg.get("mask_padding", False), "mask_padding=True breaks backward in the reference (Q3)"
    # tacotron2nv.py:88-101: freeze_charemb detaches the embedded characters (both their use by the encoder and by the residual
    # connection), freeze_encoder the encoder output (incl. the residual term), use_residual_encoder adds the embedded characters
    Pe = dict(P)
    if cfg.get("freeze_charemb", False):
        Pe["embedding.weight"] = P["embedding.weight"].detach()
    enc = encoder(Pe, cfg, inputs, input_lengths, masks, stats, train, inter)
    if cfg.get("use_residual_encoder", False):
        enc = enc + Pe["embedding.weight"][inputs]
    if cfg.get("freeze_encoder", False):
        enc = enc.detach()
    memory = speaker_concat(P, cfg, enc, speaker_vecs)
    if inter is not None:
        inter["memory"] = memory
    mel, gate, align = decoder_forward(P, cfg, memory, melspecs, masks, train, inter)
    if cfg.get("freeze_decoder", False):          # tacotron2nv.py:118-121: only the postnet learns
        mel, gate, align = mel.detach(), gate.detach(), align.detach()
    post = postnet(P, cfg, mel, masks, stats, train, inter)
    return [mel, mel + post, gate, align]


def infer(P, cfg, inputs, input_lengths, speaker_vecs, prenet_masks, stats, max_steps=None, return_gates=False):
    """Tacotron2NV.infer / Decoder.infer, tacotron2nv.py:130-162, decoder.py:334-411.

    eval mode: BN uses running stats, only the prenet dropout stays on (Q4);
    prenet_masks: [steps, 2, B, prenet_dim].  Returns (mel_post [B,n_mel,T'],
    mel_lengths int32 [B], align [B,T',L]).
    """
    with torch.no_grad():
        enc = encoder(P, cfg, inputs, input_lengths, None, stats, False)
        if cfg.get("use_residual_encoder", False):
            enc = enc + P["embedding.weight"][inputs]
        memory = speaker_concat(P, cfg, enc, speaker_vecs)
        B = memory.shape[0]
        st = AttnState(P, cfg, memory)
        state = _zero_state(cfg, memory)
        frame = memory.new_zeros(B, cfg["n_mel_channels"])
        mel_lengths = torch.zeros(B, dtype=torch.int32, device=memory.device)
        not_finished = torch.ones(B, dtype=torch.int32, device=memory.device)
        max_steps = max_steps or cfg["max_decoder_steps"]
        early = not cfg["decoder_no_early_stopping"]
        mels, aligns, gates = [], [], []
        while True:
            x = prenet(P, frame, prenet_masks[len(mels)])
            mel, gate, a, state = decode_step(P, cfg, st, x, state, memory, None, None, False)
            mels.append(mel); aligns.append(a); gates.append(gate.squeeze(1))
            dec = torch.le(torch.sigmoid(gate), cfg["gate_threshold"]).to(torch.int32).squeeze(1)
            not_finished = not_finished * dec
            mel_lengths += not_finished
            if early and int(not_finished.sum()) == 0:
                break
            if len(mels) == max_steps:
                break
            frame = mel
        mel = torch.stack(mels).permute(1, 2, 0)
        post = postnet(P, cfg, mel, None, stats, False)
        # tacotron2nv.py:160: unfold/transpose of the cat'ed [T'*B, L] alignments == [B, T', L]
        align = torch.stack(aligns).transpose(0, 1)
        if return_gates:
            return mel + post, mel_lengths, align, torch.stack(gates, dim=1)
        return mel + post, mel_lengths, align


def loss_fn(outputs, targets, mel_len, reduction="none", pos_weight=10.0, n_frames_per_step=1):
    """Tacotron2Loss.__call__, tacotron2nv_loss.py:17-52 (+ _pad_mask 55-61)."""
    pre, post, gate, _ = outputs
    mel, stop = targets
    B, n_mel, T = mel.shape
    pw = torch.tensor(pos_weight, dtype=gate.dtype, device=gate.device)
    l1 = (post - mel).abs() + (pre - mel).abs()
    mse = (post - mel) ** 2 + (pre - mel) ** 2
    bce = F.binary_cross_entropy_with_logits(gate, stop.to(gate.dtype), pos_weight=pw, reduction="none")
    if reduction == "none":
        max_len = int(mel_len.max())
        r = n_frames_per_step
        rem = max_len % r
        pad_len = max_len + (r - rem) if rem > 0 else max_len
        assert pad_len == T, "Tacotron2Loss(reduction='none') needs T == padded max(mel_len)"
        m = (torch.arange(T, device=mel.device)[None, :] < mel_len.to(mel.device)[:, None]).to(mel.dtype)        # [B, T]
        w = m / m.sum(dim=1, keepdim=True)
        ow = (w / (B * n_mel)).unsqueeze(1)                                    # [B, 1, T]
        lw = w / B
        return (l1 * ow).sum() + (mse * ow).sum() + (bce * lw).sum()
    elif reduction == "mean":
        return _mean_split(post, pre, mel) + bce.mean()
    raise ValueError(reduction)


def _mean_split(post, pre, mel):
    # l1_criterion(post)+l1_criterion(pre)+mse_criterion(post)+mse_criterion(pre), each a plain mean
    return ((post - mel).abs().mean() + (pre - mel).abs().mean()
            + ((post - mel) ** 2).mean() + ((pre - mel) ** 2).mean())


def param_names(cfg) -> List[str]:
    """model.parameters() order."""
    n = ["embedding.weight"]
    for i in range(cfg["encoder_n_convolutions"]):
        p = f"encoder.convolutions.{i}"
        n += [p + ".0.conv.weight", p + ".0.conv.bias", p + ".1.weight", p + ".1.bias"]
    for sfx in ("", "_reverse"):
        n += [f"encoder.lstm.{k}_l0{sfx}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    if cfg["speaker_emb_type"] == "learnable_lookup":
        n += ["speaker_embedder.weight"]
    elif cfg["speaker_emb_type"] == "static+linear":
        n += ["speaker_lin.weight", "speaker_lin.bias"]
    n += [f"decoder.prenet.layers.{i}.linear_layer.weight" for i in range(2)]
    n += [f"decoder.attention_rnn.{k}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    a = "decoder.attention_layer."
    n += [a + "query_layer.linear_layer.weight", a + "inputs_layer.linear_layer.weight",
          a + "v.linear_layer.weight", a + "v.linear_layer.bias"]
    if cfg["attention_params"]["trans_agent"]:
        n += [a + "ta.weight", a + "ta.bias"]
    n += [a + "location_layer.location_conv1d.weight", a + "location_layer.location_dense.linear_layer.weight"]
    n += [f"decoder.decoder_rnn.{k}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    n += ["decoder.linear_projection.linear_layer.weight", "decoder.linear_projection.linear_layer.bias",
          "decoder.gate_layer.linear_layer.weight", "decoder.gate_layer.linear_layer.bias"]
    for i in range(cfg["postnet_n_convolutions"]):
        p = f"postnet.convolutions.{i}"
        n += [p + ".0.conv.weight", p + ".0.conv.bias", p + ".1.weight", p + ".1.bias"]
    return n


def bn_layers(cfg) -> List[str]:
    return [f"encoder.convolutions.{i}.1" for i in range(cfg["encoder_n_convolutions"])] + \
           [f"postnet.convolutions.{i}.1" for i in range(cfg["postnet_n_convolutions"])]


def fresh_bn_stats(P, cfg) -> dict:
    s = {}
    for k in bn_layers(cfg):
        c = P[k + ".weight"].shape[0]
        s[k + ".running_mean"] = torch.zeros(c, dtype=P[k + ".weight"].dtype)
        s[k + ".running_var"] = torch.ones(c, dtype=P[k + ".weight"].dtype)
        s[k + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return s
