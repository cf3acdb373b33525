"""Host driver of the CUDA hot path: flat buffers, workspace, one teacher-forced pass.

This is the thin layer between the reference-shaped Python API (model.py, innerloop.py, maml.py ...)
and the C ABI (include/msa_b200.h).  PyTorch is used only for device memory, streams and (in
parallel.py) torch.distributed; every computation is a call into libmsa_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .config import rnn_dims, speaker_dim
from .layout import FlatLayout

_SPK_MODES = {"static": 0, "static+linear": 1, "learnable_lookup": 2}


def make_c_config(cfg: dict, reduction: str = "none", pos_weight: float = 10.0, gemm_tf32: int = 0) -> "_lib.MsaConfig":
    """params["model"] (+ criterion settings, metatrainer.py:83-86) -> msa_config."""
    ap = cfg["attention_params"]
    if ap["attention_type"] != "ForwardAttention":
        raise NotImplementedError('attention_type "LSA" is broken in the reference itself (SURVEY.md Q1)')
    if cfg["n_frames_per_step"] != 1:
        raise NotImplementedError("n_frames_per_step > 1 cannot train in the reference (SURVEY.md Q14)")
    if cfg.get("mask_padding", False):
        raise NotImplementedError("mask_padding=True breaks backward in the reference (SURVEY.md Q3)")
    if cfg["symbols_embedding_dim"] != cfg["encoder_embedding_dim"]:
        raise ValueError("symbols_embedding_dim must equal encoder_embedding_dim (tacotron2nv.py:88)")
    ha, hd = rnn_dims(cfg)
    c = _lib.MsaConfig()
    c.n_symbols = cfg["n_symbols"]
    c.enc_dim = cfg["encoder_embedding_dim"]
    c.enc_kernel = cfg["encoder_kernel_size"]
    c.enc_n_convs = cfg["encoder_n_convolutions"]
    c.spk_mode = _SPK_MODES[cfg["speaker_emb_type"]]
    c.spk_in_dim = cfg["speaker_embedding_dim"]
    c.spk_dim = speaker_dim(cfg)
    c.num_speakers = cfg.get("num_speakers", 1)
    c.n_mel = cfg["n_mel_channels"]
    c.prenet_dim = cfg["prenet_dim"]
    c.attn_rnn_dim = ha
    c.dec_rnn_dim = hd
    c.attn_dim = ap["attention_dim"]
    c.loc_filters = ap["attention_location_n_filters"]
    c.loc_kernel = ap["attention_location_kernel_size"]
    c.post_dim = cfg["postnet_embedding_dim"]
    c.post_kernel = cfg["postnet_kernel_size"]
    c.post_n_convs = cfg["postnet_n_convolutions"]
    c.attn_norm = {"softmax": 0, "sigmoid": 1}[ap["norm"]]
    c.forward_attn = int(ap["forward_attn"])
    c.trans_agent = int(ap["trans_agent"])
    c.windowing = int(ap["windowing"])
    c.forward_attn_mask = int(ap["forward_attn_mask"])
    c.max_decoder_steps = cfg["max_decoder_steps"]
    c.early_stopping = int(not cfg["decoder_no_early_stopping"])
    c.loss_reduction = {"none": 0, "mean": 1}[reduction]
    c.gemm_tf32 = int(gemm_tf32)
    c.p_attn_dropout = cfg["p_attention_dropout"]
    c.p_dec_dropout = cfg["p_decoder_dropout"]
    c.gate_threshold = cfg["gate_threshold"]
    c.loss_pos_weight = pos_weight
    c.freeze_charemb = int(bool(cfg.get("freeze_charemb", False)))          # tacotron2nv.py:88-121
    c.freeze_encoder = int(bool(cfg.get("freeze_encoder", False)))
    c.freeze_decoder = int(bool(cfg.get("freeze_decoder", False)))
    c.residual_encoder = int(bool(cfg.get("use_residual_encoder", False)))
    return c


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One msa_handle + its workspace on one GPU."""

    def __init__(self, cfg: dict, device: Optional[torch.device] = None, reduction: str = "none",
                 pos_weight: float = 10.0, gemm_tf32: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("msa_tts_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.layout = FlatLayout(cfg)
        self._ccfg = make_c_config(cfg, reduction, pos_weight, gemm_tf32)
        h = C.c_void_p()
        _lib.check(self.lib.msa_create(C.byref(self._ccfg), self.device.index or 0, C.byref(h)), "msa_create")
        self.h = h
        self._check_layout()
        self._gap_idx = self._make_gap_index()
        self._ws: Optional[torch.Tensor] = None
        self._enc_prefix: Optional[int] = None
        self._partials = torch.empty(self.lib.msa_flat_partials(), dtype=torch.float32, device=self.device)
        self._keep: tuple = ()
        self.launches = 0   # kernels of this library enqueued by the calls below (bench.py's gpu_launches)
        self._last_out = None
        self._abort_host: Optional[torch.Tensor] = None
        self._abort_evt = [None, None]
        self._abort_k = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.msa_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- layout -----------------------------------------------------------------------------------
    def _check_layout(self) -> None:
        n = self.lib.msa_param_count(self.h)
        names = self.layout.names()
        if n != len(names) or self.lib.msa_param_total(self.h) != self.layout.total:
            raise RuntimeError("flat layout mismatch between layout.py and libmsa_b200")
        for i, name in enumerate(names):
            cn, off, ne = C.c_char_p(), C.c_int64(), C.c_int64()
            _lib.check(self.lib.msa_param_info(self.h, i, C.byref(cn), C.byref(off), C.byref(ne)), "msa_param_info")
            if cn.value.decode() != name or off.value != self.layout.offsets[name] or ne.value != self.layout.numel(name):
                raise RuntimeError(f"flat layout mismatch at {name}")
        if self.lib.msa_bn_total(self.h) != self.layout.bn_total:
            raise RuntimeError("BN layout mismatch")

    def new_flat(self, fill: Optional[float] = 0.0) -> torch.Tensor:
        """A flat buffer in the parameter layout.  ``fill=None``: the tensor regions are left uninitialised (the caller overwrites
        them), but the alignment gaps between tensors are always zeroed: the fused kernels stream over the WHOLE buffer, so garbage
        in a gap would end up in norms, Fisher sums and penalties."""
        t = torch.empty(self.layout.total, dtype=torch.float32, device=self.device)
        if fill is not None:
            t.fill_(fill)
        elif self._gap_idx.numel():
            t.index_fill_(0, self._gap_idx, 0.0)
        return t

    def _make_gap_index(self) -> torch.Tensor:
        idx, pos = [], 0
        for name in self.layout.names():
            off = self.layout.offsets[name]
            idx.extend(range(pos, off))
            pos = off + self.layout.numel(name)
        idx.extend(range(pos, self.layout.total))
        return torch.tensor(idx, dtype=torch.int64, device=self.device)

    def flat_from_dict(self, P: Dict[str, torch.Tensor]) -> torch.Tensor:
        flat = torch.zeros(self.layout.total, dtype=torch.float32)
        for name in self.layout.names():
            o, n = self.layout.offsets[name], self.layout.numel(name)
            flat[o:o + n] = P[name].detach().reshape(-1).float().cpu()
        return flat.to(self.device)

    def dict_from_flat(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return {n: flat[self.layout.offsets[n]:self.layout.offsets[n] + self.layout.numel(n)].view(self.layout.shapes[n])
                for n in self.layout.names()}

    def new_bn_stats(self) -> torch.Tensor:
        """Fresh BatchNorm running stats (mean 0, var 1), private per task."""
        t = torch.zeros(self.layout.bn_total, dtype=torch.float32)
        for o, c in zip(self.layout.bn_offsets, self.layout.bn_ch):
            cp = (c + 31) // 32 * 32
            t[o + cp:o + cp + c] = 1.0
        return t.to(self.device)

    def bn_from_dict(self, stats: Dict[str, torch.Tensor]) -> torch.Tensor:
        """{"<bn layer>.running_mean" / ".running_var": tensor} (reference state_dict keys) -> flat BN buffer."""
        t = torch.zeros(self.layout.bn_total, dtype=torch.float32)
        for name, o, c in zip(self.layout.bn_names, self.layout.bn_offsets, self.layout.bn_ch):
            cp = (c + 31) // 32 * 32
            t[o:o + c] = stats[name + ".running_mean"].detach().float().cpu()
            t[o + cp:o + cp + c] = stats[name + ".running_var"].detach().float().cpu()
        return t.to(self.device)

    def bn_dict(self, stats: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for name, o, c in zip(self.layout.bn_names, self.layout.bn_offsets, self.layout.bn_ch):
            cp = (c + 31) // 32 * 32
            out[name + ".running_mean"] = stats[o:o + c]
            out[name + ".running_var"] = stats[o + cp:o + cp + c]
        return out

    # ---- masks ------------------------------------------------------------------------------------
    def mask_bytes(self, B: int, T: int, L: int) -> int:
        return int(self.lib.msa_mask_total(self.h, B, T, L))

    def mask_sections(self, B: int, T: int, L: int):
        out = []
        for i in range(self.lib.msa_mask_count(self.h)):
            nm, off, ne, p = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_float()
            _lib.check(self.lib.msa_mask_info(self.h, i, B, T, L, C.byref(nm), C.byref(off), C.byref(ne), C.byref(p)), "msa_mask_info")
            out.append((nm.value.decode(), off.value, ne.value, p.value))
        return out

    def pack_masks(self, masks: dict, B: int, T: int, L: int) -> torch.Tensor:
        """Reference-layout keep-masks (synth.make_masks / oracle) -> the library's uint8 buffer."""
        secs = self.mask_sections(B, T, L)
        parts = [m.permute(0, 2, 1) for m in masks["enc"]] + list(masks["prenet"]) + [masks["attn_h"], masks["dec_h"]] + \
                [m.permute(0, 2, 1) for m in masks["post"]]
        assert len(parts) == len(secs)
        buf = torch.zeros(self.mask_bytes(B, T, L), dtype=torch.uint8)
        for (_, off, ne, _), m in zip(secs, parts):
            assert m.numel() == ne, (m.shape, ne)
            buf[off:off + ne] = m.contiguous().reshape(-1).to(torch.uint8)
        return buf.to(self.device)

    def generate_masks(self, B: int, T: int, L: int, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(self.mask_bytes(B, T, L), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.msa_masks_generate(self.h, _ptr(out), B, T, L, C.c_uint64(seed & (2 ** 64 - 1)), _stream()),
                   "msa_masks_generate")
        self.launches += 1
        return out

    # ---- one pass -----------------------------------------------------------------------------------
    def _workspace(self, B: int, T: int, L: int) -> torch.Tensor:
        need = int(self.lib.msa_workspace_bytes(self.h, B, T, L))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        return self._ws

    def _ws_ptr(self):
        a = self._ws.data_ptr()
        return C.c_void_p((a + 255) // 256 * 256)

    def forward(self, params: torch.Tensor, bn_stats: Optional[torch.Tensor], batch_dev: dict, masks: torch.Tensor,
                outputs: bool = True):
        """batch_dev: dict of device tensors (inputs, input_lengths, melspecs, melspec_lengths, speaker_vecs, stop)."""
        inp, mel = batch_dev["inputs"], batch_dev["melspecs"]
        B, L = inp.shape
        T = mel.shape[2]
        self._workspace(B, T, L)
        dev = self.device
        out = None
        if outputs:
            out = [torch.empty(B, self.cfg["n_mel_channels"], T, device=dev), torch.empty(B, self.cfg["n_mel_channels"], T, device=dev),
                   torch.empty(B, T, device=dev), torch.empty(B, T, L, device=dev)]
        loss = torch.empty(1, device=dev)
        spk = batch_dev["speaker_vecs"]
        spk_f = spk if spk.dtype.is_floating_point else None
        spk_i = None if spk.dtype.is_floating_point else spk
        self._keep = (params, bn_stats, batch_dev, masks)          # the library keeps raw pointers until backward
        rc = self.lib.msa_train_forward(self.h, self._ws_ptr(), C.c_size_t(self._ws.numel() - 256), _ptr(params), _ptr(bn_stats),
                                        _ptr(inp), _ptr(batch_dev["input_lengths"]), _ptr(mel), _ptr(batch_dev["melspec_lengths"]),
                                        _ptr(spk_f), _ptr(spk_i), _ptr(batch_dev.get("stop")), _ptr(masks), B, T, L,
                                        _ptr(out[0]) if out else None, _ptr(out[1]) if out else None,
                                        _ptr(out[2]) if out else None, _ptr(out[3]) if out else None, _ptr(loss), _stream())
        _lib.check(rc, "msa_train_forward")
        self.launches += 1
        return out, loss

    def encoder_prefix(self) -> int:
        """Floats of the encoder part of a flat buffer (embedding.weight, encoder.*: the prefix of the layout), whose gradients are the
        LAST ones a backward pass produces."""
        if self._enc_prefix is None:
            n = C.c_int64()
            _lib.check(self.lib.msa_backward_mark_event(self.h, None, C.byref(n)), "msa_backward_mark_event")
            self._enc_prefix = int(n.value)
        return self._enc_prefix

    def backward(self, params: torch.Tensor, grads: torch.Tensor, accumulate: bool = False, scale: float = 1.0,
                 d_outputs: Optional[Sequence[torch.Tensor]] = None, decoder_done: Optional["torch.cuda.Event"] = None) -> None:
        """``decoder_done``: a CUDA event the pass records as soon as ``grads[encoder_prefix():]`` is final (msa_backward_mark_event)."""
        if decoder_done is not None:
            _lib.check(self.lib.msa_backward_mark_event(self.h, C.c_void_p(decoder_done.cuda_event), None), "msa_backward_mark_event")
        d = [None, None, None] if d_outputs is None else [x.contiguous() for x in d_outputs]
        rc = self.lib.msa_train_backward(self.h, self._ws_ptr(), C.c_size_t(self._ws.numel() - 256), _ptr(params), _ptr(d[0]),
                                         _ptr(d[1]), _ptr(d[2]), _ptr(grads), int(accumulate), C.c_float(scale), _stream())
        _lib.check(rc, "msa_train_backward")
        self.launches += 1

    # ---- task groups: the theta_0 train-split passes of several tasks as one pass ------------------------------
    GROUP_MAX = 8

    def group_size(self, n_tasks: int, B: int) -> int:
        """Largest number of tasks one grouped pass may carry: at most 8 tasks and 32 rows per shared recurrence launch."""
        return max(1, min(n_tasks, self.GROUP_MAX, 32 // max(B, 1)))

    def _ptr_array(self, tensors):
        return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

    def forward_group(self, params, bn_stats: Sequence[Optional[torch.Tensor]], batches_dev: Sequence[dict],
                      masks: Sequence[torch.Tensor]) -> torch.Tensor:
        """One grouped forward of ``len(batches_dev)`` tasks (include/msa_b200.h "task groups").  ``params``: ONE flat weight buffer
        shared by all tasks (the theta_0 train passes: the recurrences share one launch) or a list with one buffer per task (the
        test passes with the adapted weights: recurrences task by task, the stages between them overlapped across tasks).
        Returns the per-task losses [G] (device).  All tasks must share B, T, L."""
        G = len(batches_dev)
        plist = list(params) if isinstance(params, (list, tuple)) else [params] * G
        if len(plist) != G:
            raise ValueError("forward_group: one weight buffer per task (or a single shared one)")
        B, L = batches_dev[0]["inputs"].shape
        T = batches_dev[0]["melspecs"].shape[2]
        for bd in batches_dev:
            if tuple(bd["inputs"].shape) != (B, L) or bd["melspecs"].shape[2] != T:
                raise ValueError("forward_group: all tasks of a group must have the same B, T, L")
        need = int(self.lib.msa_group_workspace_bytes(self.h, G, B, T, L))
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        self._group = (G, B, T, L, int(self.lib.msa_group_workspace_bytes(self.h, 1, B, T, L)))
        loss = torch.empty(G, device=self.device)
        spk = [bd["speaker_vecs"] for bd in batches_dev]
        fl = spk[0].dtype.is_floating_point
        self._keep = (plist, list(bn_stats), list(batches_dev), list(masks))
        A = self._ptr_array
        rc = self.lib.msa_train_forward_group(
            self.h, G, self._ws_ptr(), C.c_size_t(self._ws.numel() - 256), A(plist), A(list(bn_stats)),
            A([bd["inputs"] for bd in batches_dev]), A([bd["input_lengths"] for bd in batches_dev]),
            A([bd["melspecs"] for bd in batches_dev]), A([bd["melspec_lengths"] for bd in batches_dev]),
            A(spk) if fl else None, None if fl else A(spk), A([bd["stop"] for bd in batches_dev]), A(list(masks)), B, T, L,
            _ptr(loss), _stream())
        _lib.check(rc, "msa_train_forward_group")
        self.launches += 1
        return loss

    def backward_group(self, params, grads: Sequence[torch.Tensor], accumulate: bool = False, scale: float = 1.0) -> None:
        """grads[g] <- (or +=) scale * d loss_g / d params[g] for every task of the last ``forward_group`` (same ``params``)."""
        plist = list(params) if isinstance(params, (list, tuple)) else [params] * len(grads)
        rc = self.lib.msa_train_backward_group(self.h, self._ws_ptr(), C.c_size_t(self._ws.numel() - 256), self._ptr_array(plist),
                                               self._ptr_array(list(grads)), int(accumulate), C.c_float(scale), _stream())
        _lib.check(rc, "msa_train_backward_group")
        self.launches += 1

    def group_ws_ptr(self, g: int):
        a = (self._ws.data_ptr() + 255) // 256 * 256
        return C.c_void_p(a + g * self._group[4])

    def mcd_group(self, g: int, mel_lengths: torch.Tensor, which: int = 0) -> torch.Tensor:
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_train_mcd(self.h, self.group_ws_ptr(g), _ptr(mel_lengths), int(which), _ptr(out), _stream()), "msa_train_mcd")
        self.launches += 3
        return out

    def loss(self, stop: torch.Tensor, mel_lengths: torch.Tensor, reduction: str = "none", pos_weight: float = 10.0) -> torch.Tensor:
        """Tacotron2Loss on the outputs of the last forward (tacotron2nv_loss.py:17-52)."""
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_train_loss(self.h, self._ws_ptr(), _ptr(stop), _ptr(mel_lengths), {"none": 0, "mean": 1}[reduction],
                                           C.c_float(pos_weight), _ptr(out), _stream()), "msa_train_loss")
        self.launches += 1
        return out

    def mcd(self, mel_lengths: torch.Tensor, which: int = 0) -> torch.Tensor:
        """utils/metrics.py:15-22 ``mcd_batch`` of the last forward on the device (1-element tensor, no sync); which = 0: the first
        model output, the one the reference's logs use (maml.py:78-82), 1: mel_post."""
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_train_mcd(self.h, self._ws_ptr(), _ptr(mel_lengths), int(which), _ptr(out), _stream()), "msa_train_mcd")
        self.launches += 3
        return out

    def loss_grads(self, B: int, T: int):
        M = self.cfg["n_mel_channels"]
        d = [torch.empty(B, M, T, device=self.device), torch.empty(B, M, T, device=self.device), torch.empty(B, T, device=self.device)]
        _lib.check(self.lib.msa_loss_grads(self.h, self._ws_ptr(), _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _stream()), "msa_loss_grads")
        return d

    INFER_ROWS = 32      # rows of one msa_infer call (one batch tile of the fused decoder-step kernels)

    def _infer_chunked(self, params, bn_stats, inputs, input_lengths, speaker_vecs, prenet_masks, max_steps):
        """Batches beyond one 32-row tile: inference has no cross-row coupling in eval mode (BatchNorm uses running statistics), so the
        rows are decoded tile by tile.  ``mel_lengths`` and every frame below a row's length are those of one big call; the output
        is as long as the longest tile ran, and a tile that stopped earlier is zero-padded there (the reference keeps decoding
        finished rows until the whole batch has stopped -- frames nobody reads, infer.py uses ``mel_lengths``)."""
        B = inputs.shape[0]
        outs = []
        pm = prenet_masks.to(device=self.device, dtype=torch.uint8)
        for b0 in range(0, B, self.INFER_ROWS):
            sl = slice(b0, min(B, b0 + self.INFER_ROWS))
            outs.append(self.infer(params, bn_stats, inputs[sl], input_lengths[sl], speaker_vecs[sl], pm[:, :, sl].contiguous(), max_steps))
        T = max(o[0].shape[2] for o in outs)
        post = torch.zeros(B, outs[0][0].shape[1], T, device=self.device)
        align = torch.zeros(B, T, inputs.shape[1], device=self.device)
        b0 = 0
        for mp, _, al in outs:
            n, t = mp.shape[0], mp.shape[2]
            post[b0:b0 + n, :, :t] = mp
            align[b0:b0 + n, :t] = al
            b0 += n
        return post, torch.cat([o[1] for o in outs]), align

    def infer(self, params: torch.Tensor, bn_stats: torch.Tensor, inputs: torch.Tensor, input_lengths: torch.Tensor,
              speaker_vecs: torch.Tensor, prenet_masks: torch.Tensor, max_steps: Optional[int] = None):
        """Tacotron2NV.infer (tacotron2nv.py:130-162): -> (mel_post [B, n_mel, T'], mel_lengths int32 [B], align [B, T', L]).

        prenet_masks: uint8 keep-masks [max_steps, 2, B, prenet_dim] (the prenet dropout stays on in eval)."""
        B, L = inputs.shape
        max_steps = int(max_steps or self.cfg["max_decoder_steps"])
        M = self.cfg["n_mel_channels"]
        if B > self.INFER_ROWS:
            return self._infer_chunked(params, bn_stats, inputs, input_lengths, speaker_vecs, prenet_masks, max_steps)
        need = int(self.lib.msa_infer_workspace_bytes(self.h, B, L, max_steps))
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        dev = self.device
        mel_post = torch.empty(B, M, max_steps, device=dev)
        mel_len = torch.empty(B, dtype=torch.int32, device=dev)
        align = torch.empty(B, max_steps, L, device=dev)
        n_steps = torch.zeros(1, dtype=torch.int32, device=dev)
        pm = prenet_masks.to(device=dev, dtype=torch.uint8).contiguous()
        if pm.numel() < max_steps * 2 * B * self.cfg["prenet_dim"]:
            raise ValueError("prenet_masks must cover max_steps decoder steps")
        spk = speaker_vecs.to(dev).contiguous()
        spk_f = spk if spk.dtype.is_floating_point else None
        spk_i = None if spk.dtype.is_floating_point else spk
        inp = inputs.to(dev).contiguous()
        inl = input_lengths.to(dev).contiguous()
        rc = self.lib.msa_infer(self.h, self._ws_ptr(), C.c_size_t(self._ws.numel() - 256), _ptr(params), _ptr(bn_stats), _ptr(inp),
                                _ptr(inl), _ptr(spk_f), _ptr(spk_i), _ptr(pm), B, L, max_steps, _ptr(mel_post), _ptr(mel_len),
                                _ptr(align), _ptr(n_steps), _stream())
        _lib.check(rc, "msa_infer")
        self.launches += 1
        T = int(n_steps.item())
        return mel_post[:, :, :T], mel_len, align[:, :T, :]

    def get_buffer(self, name: str) -> torch.Tensor:
        """Copy of a named intermediate of the last pass (tests only)."""
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.msa_get_buffer(self.h, self._ws_ptr(), name.encode(), C.byref(p), C.byref(n)), "msa_get_buffer")
        off = (p.value - self._ws.data_ptr())
        return self._ws[off:off + 4 * n.value].view(torch.float32).clone()

    # ---- instrumentation ---------------------------------------------------------------------------------
    def kernel_launches(self) -> int:
        """Kernels of libmsa_b200 enqueued so far by this process (cuBLAS GEMMs excluded)."""
        return int(self.lib.msa_launch_count())

    def profile(self, enable, inkernel: bool = False) -> None:
        """CUDA-event timing of the persistent kernels; inkernel=True also selects their instrumented variants."""
        _lib.check(self.lib.msa_profile_enable(self.h, (2 if inkernel else 1) if enable else 0), "msa_profile_enable")

    def check_abort(self) -> None:
        """Raise if a persistent kernel of this engine timed out while polling since the last ``abort_clear`` (synchronises)."""
        _lib.check(self.lib.msa_check_abort(self.h, None, _stream()), "msa_check_abort")

    def abort_guard(self, sumsq: torch.Tensor) -> None:
        """Device side, no sync: NaN into the gradient-norm scalar if a persistent kernel gave up; the clip + optimizer kernels skip
        their update on a non-finite norm, so garbage gradients never reach the weights or the optimizer state."""
        _lib.check(self.lib.msa_abort_guard(self.h, _ptr(sumsq), _stream()), "msa_abort_guard")
        self.launches += 1

    def abort_poll(self) -> None:
        """Deferred host-side check without a stall: enqueue a 4-byte device->host copy of the abort word into pinned memory and
        raise if the PREVIOUS copy (complete by now in any loop that reads a loss per step) saw it set."""
        if self._abort_host is None:
            self._abort_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._abort_evt = [None, None]
        k = self._abort_k
        prev = 1 - k
        if self._abort_evt[prev] is not None:
            self._abort_evt[prev].synchronize()
            if int(self._abort_host[prev]) != 0:
                self._abort_raise()
        _lib.check(self.lib.msa_abort_read_async(self.h, C.c_void_p(self._abort_host.data_ptr() + 4 * k), _stream()), "msa_abort_read_async")
        ev = torch.cuda.Event()
        ev.record()
        self._abort_evt[k] = ev
        self._abort_k = prev

    def abort_flush(self) -> None:
        """Wait for the outstanding ``abort_poll`` copies and raise if one saw the word (end of an epoch / of a bench loop)."""
        if self._abort_host is None:
            return
        for k in (0, 1):
            if self._abort_evt[k] is not None:
                self._abort_evt[k].synchronize()
                self._abort_evt[k] = None
                if int(self._abort_host[k]) != 0:
                    self._abort_raise()

    def _abort_raise(self):
        self._abort_host.zero_()
        self._abort_evt = [None, None]
        _lib.check(self.lib.msa_abort_clear(self.h, _stream()), "msa_abort_clear")
        raise RuntimeError("libmsa_b200: a persistent kernel gave up waiting for data of another CTA (polling time-out); the "
                           "gradients of this meta-step were discarded on the device (the outer update was skipped)")

    def abort_clear(self) -> None:
        _lib.check(self.lib.msa_abort_clear(self.h, _stream()), "msa_abort_clear")

    def debug_raise_abort(self) -> None:
        """Test hook: raise the abort word exactly as a polling thread that timed out would."""
        _lib.check(self.lib.msa_debug_raise_abort(self.h, _stream()), "msa_debug_raise_abort")

    def profile_phases(self, name: str):
        """[ncta][8] cycles per phase of persistent kernel `name` in the last profiled pass (profiles only)."""
        n = self.lib.msa_profile_kernels()
        ids = {self.lib.msa_profile_name(i).decode(): i for i in range(n)}
        ncta = self.lib.msa_sm_count(self.h)
        out = (C.c_int64 * (ncta * 8))()
        _lib.check(self.lib.msa_profile_phases(self.h, self._ws_ptr(), ids[name], out, ncta), "msa_profile_phases")
        return [[out[i * 8 + j] for j in range(8)] for i in range(ncta)]

    def profile_trace(self, name: str):
        """numpy [ncta][16][4][12][2] int64 {SM clock, globaltimer ns} of the traced steps (profiles only)."""
        import numpy as np
        n = self.lib.msa_profile_kernels()
        ids = {self.lib.msa_profile_name(i).decode(): i for i in range(n)}
        ncta = self.lib.msa_sm_count(self.h)
        out = (C.c_int64 * (ncta * 16 * 4 * 12 * 2))()
        _lib.check(self.lib.msa_profile_trace(self.h, self._ws_ptr(), ids[name], out, ncta), "msa_profile_trace")
        return np.ctypeslib.as_array(out).reshape(ncta, 16, 4, 12, 2).copy()

    def profile_read(self) -> dict:
        n = self.lib.msa_profile_kernels()
        ms, cnt = (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(self.lib.msa_profile_read(self.h, ms, cnt), "msa_profile_read")
        return {self.lib.msa_profile_name(i).decode(): (ms[i], cnt[i]) for i in range(n)}

    # ---- flat-buffer kernels ----------------------------------------------------------------------------
    def sgd_step(self, p, g, p_out=None, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, buf=None,
                 first_step=True):
        p_out = p if p_out is None else p_out
        _lib.check(self.lib.msa_flat_sgd_step(_ptr(p), _ptr(g), _ptr(p_out), _ptr(buf), p.numel(), lr, momentum, dampening,
                                              weight_decay, int(nesterov), int(first_step), _stream()), "msa_flat_sgd_step")
        self.launches += 1
        return p_out

    def axpy(self, acc, g, w: float, init: bool):
        _lib.check(self.lib.msa_flat_axpy(_ptr(acc), _ptr(g), acc.numel(), w, int(init), _stream()), "msa_flat_axpy")
        self.launches += 1

    def reptile_delta(self, acc, p_T, p_0, w: float, init: bool):
        _lib.check(self.lib.msa_flat_reptile_delta(_ptr(acc), _ptr(p_T), _ptr(p_0), acc.numel(), w, int(init), _stream()),
                   "msa_flat_reptile_delta")
        self.launches += 1

    def sumsq(self, g, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = torch.empty(1, device=self.device) if out is None else out
        _lib.check(self.lib.msa_flat_sumsq(_ptr(g), g.numel(), _ptr(self._partials), _ptr(out), _stream()), "msa_flat_sumsq")
        self.launches += 2
        return out

    def clip_sgd(self, p, g, sumsq, lr, max_norm=0.0, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, buf=None,
                 first_step=True):
        _lib.check(self.lib.msa_flat_clip_sgd(_ptr(p), _ptr(g), _ptr(buf), _ptr(sumsq), p.numel(), lr, max_norm, momentum, dampening,
                                              weight_decay, int(nesterov), int(first_step), _stream()), "msa_flat_clip_sgd")
        self.launches += 1

    def clip_adam(self, p, g, m, v, sumsq, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0):
        _lib.check(self.lib.msa_flat_clip_adam(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(sumsq), p.numel(), lr, betas[0], betas[1], eps,
                                               weight_decay, step, max_norm, _stream()), "msa_flat_clip_adam")
        self.launches += 1

    def adam_step(self, p, g, p_out, m, v, lr, step, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """Functional (out-of-place) Adam step of the inner loop; m, v are updated in place."""
        _lib.check(self.lib.msa_flat_adam_step(_ptr(p), _ptr(g), _ptr(p_out), _ptr(m), _ptr(v), p.numel(), lr, betas[0], betas[1], eps,
                                               weight_decay, step, _stream()), "msa_flat_adam_step")
        self.launches += 1

    def ewc_fisher_accum(self, fisher, g, inv_n: float, init: bool):
        _lib.check(self.lib.msa_ewc_fisher_accum(_ptr(fisher), _ptr(g), g.numel(), inv_n, int(init), _stream()), "msa_ewc_fisher_accum")
        self.launches += 1

    def ewc_penalty(self, p, mu, fisher) -> torch.Tensor:
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_ewc_penalty(_ptr(p), _ptr(mu), _ptr(fisher), p.numel(), _ptr(self._partials), _ptr(out), _stream()),
                   "msa_ewc_penalty")
        self.launches += 2
        return out

    def ewc_penalty_grad(self, p, g, mu, fisher, lam: float) -> torch.Tensor:
        """g += 2*lam*F*(p - mu) in place; returns the penalty sum F (p - mu)^2 at p (one pass over the four buffers)."""
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_ewc_penalty_grad(_ptr(p), _ptr(g), _ptr(mu), _ptr(fisher), p.numel(), lam, _ptr(self._partials),
                                                 _ptr(out), _stream()), "msa_ewc_penalty_grad")
        self.launches += 2
        return out

    def ewc_sgd_step(self, p, g, mu, fisher, lr: float, lam: float) -> torch.Tensor:
        out = torch.empty(1, device=self.device)
        _lib.check(self.lib.msa_ewc_sgd_step(_ptr(p), _ptr(g), _ptr(mu), _ptr(fisher), p.numel(), lr, lam, _ptr(self._partials),
                                             _ptr(out), _stream()), "msa_ewc_sgd_step")
        self.launches += 2
        return out


def batch_to_device(batch: tuple, device, speaker_emb_type: str = "static", non_blocking: bool = False) -> dict:
    """MetaTrainer._unpack_batch (metatrainer.py:95-117): batch tuple -> kwargs dict on the device."""
    _, inp, inp_len, mels, mel_len, spk_ids, spk_embs, stop = batch
    spk = spk_ids if speaker_emb_type == "learnable_lookup" else spk_embs
    mv = lambda t: t.to(device, non_blocking=non_blocking).contiguous()
    return {"inputs": mv(inp), "input_lengths": mv(inp_len), "melspecs": mv(mels), "melspec_lengths": mv(mel_len),
            "speaker_vecs": mv(spk), "stop": mv(stop)}
