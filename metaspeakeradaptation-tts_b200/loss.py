"""Drop-in for ``msa_tts.models.modules_tacotron2nv.tacotron2nv_loss.Tacotron2Loss`` (tacotron2nv_loss.py:17-52).

``criterion((mel, mel_post, gate, align), (mel_target, stop_target), mel_lengths)`` -> 0-d tensor, differentiable w.r.t. the
three outputs; computed by one fused kernel (msa_tacotron2_loss) that also produces the output gradients."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mel, post, gate, target, stop, mel_len, reduction, pos_weight):
        lib = _lib.load()
        B, M, T = mel.shape
        dev = mel.device
        scratch = torch.empty(int(lib.msa_loss_scratch_floats(B, T, M)), device=dev)
        loss = torch.empty(1, device=dev)
        d = [torch.empty_like(mel), torch.empty_like(post), torch.empty_like(gate)]
        P = lambda t: C.c_void_p(t.data_ptr())
        rc = lib.msa_tacotron2_loss(P(mel.contiguous()), P(post.contiguous()), P(gate.contiguous()), P(target.contiguous()),
                                    P(stop.contiguous()), P(mel_len.contiguous()), B, T, M, reduction, C.c_float(pos_weight),
                                    P(scratch), P(loss), P(d[0]), P(d[1]), P(d[2]),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "msa_tacotron2_loss")
        ctx.save_for_backward(*d)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d = ctx.saved_tensors
        return g * d[0], g * d[1], g * d[2], None, None, None, None, None


class Tacotron2Loss:
    def __init__(self, n_frames_per_step: int = 1, reduction: str = "none", pos_weight: float = 10.0, device=None):
        if n_frames_per_step != 1:
            raise NotImplementedError("n_frames_per_step > 1 cannot train in the reference (SURVEY.md Q14)")
        self.reduction = {"none": 0, "mean": 1}[reduction]
        self.pos_weight = float(pos_weight)
        self.device = device

    def __call__(self, outputs, targets, mel_lengths):
        mel, post, gate = outputs[0], outputs[1], outputs[2]
        target, stop = targets
        dev = mel.device
        return _LossFn.apply(mel, post, gate, target.to(dev).float(), stop.to(dev).float(), mel_lengths.to(dev).long(),
                             self.reduction, self.pos_weight)
