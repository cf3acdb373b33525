"""Host-side batch assembly in the reference's data contract.

``Collator`` / ``MetaCollator`` produce exactly the tuples the trainers consume (``MetaTrainer._unpack_batch``):

    (item_ids, transcripts int64 [B, L] zero-padded and SORTED by length descending, trans_lengths int64 [B],
     melspecs f32 [B, n_mels, T] zero-padded to a multiple of the reduction factor, melspec_lengths int64 [B],
     speaker_ids int64 [B], spk_embs f32 [B, Ds], stop_targets f32 [B, T] = 0...0,1 then padded with 1.0)

following msa_tts/dataloaders/dataloader_default.py:109-220 and dataloader_meta.py:124-243 (checked bit for bit against the
reference's own collators: tests/golden/collate.npz, oracle/gen_golden_collate.py).  The audio front end is out of scope, so an
item carries a ready mel-spectrogram where the reference carries a waveform: ``(item_id, transcript LongTensor [len], speaker_id,
melspec FloatTensor [1, n_mels, len] or [n_mels, len], spk_emb FloatTensor [Ds])``.

B200 side of the contract: with ``pin_memory=True`` every tensor of a batch is written straight into page-locked host memory, so
``MetaTrainer._stage`` can enqueue the host -> device copies of all batches of a meta-batch on a side stream: the copy engine moves
the later tasks' batches while the compute stream runs the earlier tasks' passes (each pass waits for its own batch's event).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


def _round_up(n: int, r: int) -> int:
    return n if n % r == 0 else n + (r - n % r)


class Collator:
    def __init__(self, reduction_factor: int = 1, pin_memory: bool = False):
        self.reduction_factor = int(reduction_factor)
        self.pin_memory = bool(pin_memory)

    def _new(self, shape, dtype, fill):
        t = torch.full(shape, fill, dtype=dtype)
        return t.pin_memory() if self.pin_memory else t

    def __call__(self, batch: Sequence[tuple]) -> tuple:
        # order: transcript length, descending -- the very torch.sort call of the reference decides ties (dataloader_default.py:127-128)
        lens = torch.LongTensor([len(it[1]) for it in batch])
        trans_lengths, order = torch.sort(lens, dim=0, descending=True)
        items = [batch[int(i)] for i in order]
        mels = [it[3] if it[3].dim() == 3 else it[3].unsqueeze(0) for it in items]
        B, n_mels = len(items), mels[0].shape[1]
        mel_lens = [m.shape[-1] for m in mels]
        L = int(trans_lengths[0])
        T = _round_up(max(mel_lens), self.reduction_factor)                       # dataloader_default.py:188-190
        transcripts = self._new((B, L), torch.int64, 0)                           # pad id 0
        melspecs = self._new((B, n_mels, T), torch.float32, 0.0)
        stop = self._new((B, T), torch.float32, 1.0)                              # padding of the stop targets is 1.0 (:209-215)
        spk_embs = self._new((B, items[0][4].shape[0]), torch.float32, 0.0)
        for b, it in enumerate(items):
            transcripts[b, :len(it[1])] = it[1]
            melspecs[b, :, :mel_lens[b]] = mels[b][0]
            stop[b, :mel_lens[b] - 1] = 0.0                                       # 0 ... 0, 1 at the last frame (:141-142)
            spk_embs[b] = it[4]
        melspec_lengths = torch.LongTensor(mel_lens)
        speaker_ids = torch.LongTensor([it[2] for it in items])
        if self.pin_memory:
            trans_lengths, melspec_lengths, speaker_ids = trans_lengths.pin_memory(), melspec_lengths.pin_memory(), speaker_ids.pin_memory()
        return ([it[0] for it in items], transcripts, trans_lengths, melspecs, melspec_lengths, speaker_ids, spk_embs, stop)


class MetaCollator(Collator):
    """[(speaker, {"train": items, "test": items}), ...] -> {speaker: {"train": batch, "test": batch}} (dataloader_meta.py:133-179)."""

    def __call__(self, batch_tuples: Sequence[Tuple[str, Dict[str, List[tuple]]]]) -> Dict[str, Dict[str, tuple]]:
        out: Dict[str, Dict[str, tuple]] = {}
        for speaker, splits in batch_tuples:
            out[speaker] = {mode: Collator.__call__(self, items) for mode, items in splits.items()}
        return out
