"""mel -> vocoder hand-off of the inference script (msa_tts/infer.py:311-328).

The reference takes the post-net mel of ONE utterance as a numpy array, rebuilds a tensor, adds a batch axis and moves it to the
vocoder's device (``hifigan.inference(torch.tensor(melspec).unsqueeze(0).to(device))``; the WaveRNN / Griffin-Lim branches get the
same ``[1, n_mel, T']`` tensor).  ``Tacotron2NV.infer`` here returns the whole batch on the device, padded to the longest utterance,
with ``mel_lengths``; ``vocoder_inputs`` turns that into what the vocoders expect -- one ``[1, n_mel, len_b]`` tensor per utterance,
still on the device (views of the inference output: no device -> host -> device round trip).  The vocoders themselves (HiFi-GAN
generator ``utils/hifigan/models.py:75-126``, WaveRNN, Griffin-Lim) are outside the hot path.
"""
from __future__ import annotations

from typing import List

import torch


def vocoder_inputs(mel_post: torch.Tensor, mel_lengths: torch.Tensor) -> List[torch.Tensor]:
    """mel_post [B, n_mel, T'] (device), mel_lengths int32 [B] -> B tensors [1, n_mel, len_b] (views, same device)."""
    if mel_post.dim() != 3 or mel_lengths.dim() != 1 or mel_lengths.shape[0] != mel_post.shape[0]:
        raise ValueError("vocoder_inputs: expected mel_post [B, n_mel, T'] and mel_lengths [B]")
    lens = mel_lengths.tolist()                      # one small device -> host read for the whole batch
    T = mel_post.shape[2]
    return [mel_post[b:b + 1, :, :max(1, min(int(n), T))] for b, n in enumerate(lens)]


def vocoder_batch(mel_post: torch.Tensor, mel_lengths: torch.Tensor, pad_value: float = -11.5129) -> torch.Tensor:
    """The same hand-off as ONE padded batch [B, n_mel, max_len] for vocoders that run batched: frames past an utterance's length
    are overwritten with ``pad_value`` (log(1e-5), the silence floor HiFi-GAN's mel front end uses)."""
    lens = mel_lengths.to(mel_post.device).long().clamp(min=1, max=mel_post.shape[2])
    tmax = int(lens.max())
    out = mel_post[:, :, :tmax].clone()
    mask = torch.arange(tmax, device=mel_post.device)[None, :] >= lens[:, None]
    return out.masked_fill_(mask[:, None, :], pad_value)
