"""Seeded case definitions shared by oracle/gen_golden.py and the tests (TEST INFRASTRUCTURE).

Each entry returns (cfg, seed, dims, extra) so that the fixture generator (which
imports the reference) and the tests (which do not) build identical inputs."""
import copy

import torch

import msa_tts_b200 as pkg

CRIT = dict(reduction="none", pos_weight=10.0)


def _small(**attn):
    c = pkg.small_params()
    c["attention_params"].update(attn)
    return c


def _spklin():
    c = pkg.small_params()
    c["speaker_emb_type"] = "static+linear"
    return c


def _lookup():
    c = pkg.small_params()
    c["speaker_emb_type"] = "learnable_lookup"
    c["num_speakers"] = 5
    return c


def _opt(**kw):
    """tacotron2nv.py:88-121: freeze_charemb / freeze_encoder / freeze_decoder / use_residual_encoder."""
    c = pkg.small_params()
    c.update(kw)
    return c


def speaker_input(cfg, batch):
    """What the trainers pass as ``speaker_vecs`` (metatrainer.py:95-117): ids for "learnable_lookup", vectors otherwise."""
    return batch[5] if cfg["speaker_emb_type"] == "learnable_lookup" else batch[6]


CASES = {
    "small_train": lambda: (_small(), 11, (3, 12, 9), CRIT),
    "small_train_meanloss": lambda: (_small(), 12, (2, 10, 12), dict(reduction="mean", pos_weight=10.0)),
    "small_train_fwdattn_sigmoid": lambda: (_small(norm="sigmoid", forward_attn=True, trans_agent=True), 13, (3, 12, 9), CRIT),
    "small_train_spklin": lambda: (_spklin(), 14, (3, 11, 10), CRIT),
    # speaker_emb_type="learnable_lookup" (tacotron2nv.py:31-34,104-105): nn.Embedding over speaker ids, rows of several speakers
    "small_train_lookup": lambda: (_lookup(), 16, (4, 11, 10), CRIT),
    "small_train_sigmoid": lambda: (_small(norm="sigmoid"), 15, (4, 13, 10), CRIT),
    "small_train_residual": lambda: (_opt(use_residual_encoder=True), 17, (3, 12, 9), CRIT),
    "small_train_freeze_charemb": lambda: (_opt(freeze_charemb=True, use_residual_encoder=True), 18, (3, 11, 9), CRIT),
    "small_train_freeze_encoder": lambda: (_opt(freeze_encoder=True), 19, (3, 12, 10), CRIT),
    "small_train_freeze_decoder": lambda: (_opt(freeze_decoder=True), 20, (4, 12, 9), CRIT),
    # BASELINE.json configs[0]: default dims, batch 4, 200 mel frames, 80 mels
    "default_train_b4_t200": lambda: (pkg.default_params(), 0, (4, 200, 64), CRIT),
}


def _infer(early=False, thr=0.5, residual=False, **attn):
    c = _small(**attn)
    c["use_residual_encoder"] = residual
    c["max_decoder_steps"] = 24
    c["decoder_no_early_stopping"] = not early
    c["gate_threshold"] = thr
    return c


INFER_CASES = {
    "small_infer": lambda: (_infer(), 21, (3, 9), 24),
    # threshold picked by gen_golden.pick_threshold so that the rows stop at different steps
    "small_infer_earlystop": lambda: (_infer(early=True, thr=0.62), 22, (3, 9), 24),
    "small_infer_residual": lambda: (_infer(residual=True), 24, (3, 10), 24),
    "small_infer_window_fwdmask": lambda: (_infer(windowing=True, forward_attn=True, forward_attn_mask=True,
                                                  trans_agent=True), 23, (3, 14), 24),
}


def infer_stats(P, cfg, seed):
    """Non-trivial BN running statistics, as left by a few adaptation passes."""
    from oracle import model as OM
    g = torch.Generator().manual_seed(seed + 7)
    stats = OM.fresh_bn_stats(P, cfg)
    for k in stats:
        if k.endswith("running_mean"):
            stats[k] = 0.2 * torch.randn(stats[k].shape, generator=g)
        elif k.endswith("running_var"):
            stats[k] = 0.5 + torch.rand(stats[k].shape, generator=g)
    return stats


def collate_items(seed: int, n: int, n_mels: int = 8, n_symbols: int = 40, spk_dim: int = 6):
    """n raw dataset items in the reference's format (dataloader_meta.py:82-110): (item_id, transcript LongTensor [len],
    speaker_id int, "waveform" slot, spk_emb FloatTensor [Ds]) -- with a ready mel-spectrogram [1, n_mels, len] in the waveform
    slot (the audio front end is out of scope).  Ragged lengths, unsorted, with ties in the transcript length."""
    import torch
    g = torch.Generator().manual_seed(seed)
    items = []
    for i in range(n):
        tl = int(torch.randint(3, 9, (1,), generator=g))
        ml = int(torch.randint(5, 14, (1,), generator=g))
        items.append((f"item_{seed}_{i}.wav", torch.randint(1, n_symbols, (tl,), generator=g, dtype=torch.int64), (seed + i) % 3,
                      torch.randn(1, n_mels, ml, generator=g), torch.randn(spk_dim, generator=g)))
    return items
