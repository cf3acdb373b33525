"""Model configuration for the hot path.

Mirrors the ``params["model"]`` dict the reference passes to
``Tacotron2NV(params)`` (msa_tts/models/tacotron2nv.py:11-66; key list in
SURVEY.md Appendix D/G).  The reference ships no ``params.yml``; the defaults
below are the Tacotron-2 dimensions every BASELINE.json config is quoted on.
"""
from __future__ import annotations

import copy

DEFAULT_MODEL_PARAMS = {
    "mask_padding": False,            # True breaks backward in the reference
    "n_mel_channels": 80,
    "n_frames_per_step": 1,           # r > 1 cannot train in the reference
    "n_symbols": 123,                 # utils/g2p/char_list.py:15
    "symbols_embedding_dim": 512,
    "encoder_n_convolutions": 3,
    "encoder_embedding_dim": 512,
    "encoder_kernel_size": 5,
    "speaker_emb_type": "static",
    "num_speakers": 8,
    "speaker_embedding_dim": 256,
    "speaker_embedding_dim_lin": 64,
    "attention_params": {
        "attention_type": "ForwardAttention",   # "LSA" is broken in the reference
        "attention_dim": 128,
        "attention_location_n_filters": 32,
        "attention_location_kernel_size": 31,
        "windowing": False,
        "norm": "softmax",
        "forward_attn": False,
        "trans_agent": False,
        "forward_attn_mask": False,
    },
    "decoder_rnn_dim": 1024,
    "attention_rnn_dim": 1024,
    "prenet_dim": 256,
    "max_decoder_steps": 1000,
    "gate_threshold": 0.5,
    "p_attention_dropout": 0.1,
    "p_decoder_dropout": 0.1,
    "decoder_no_early_stopping": False,
    "postnet_embedding_dim": 512,
    "postnet_kernel_size": 5,
    "postnet_n_convolutions": 5,
    "freeze_charemb": False,
    "freeze_encoder": False,
    "freeze_decoder": False,
    "use_residual_encoder": False,
}

# A tiny instance used by the golden fixtures and the fast parity tests.
SMALL_MODEL_PARAMS = {
    **copy.deepcopy(DEFAULT_MODEL_PARAMS),
    "n_mel_channels": 8,
    "n_symbols": 20,
    "symbols_embedding_dim": 32,
    "encoder_embedding_dim": 32,
    "speaker_embedding_dim": 8,
    "speaker_embedding_dim_lin": 4,
    "decoder_rnn_dim": 32,
    "attention_rnn_dim": 32,
    "prenet_dim": 16,
    "postnet_embedding_dim": 32,
    "max_decoder_steps": 40,
    "attention_params": {
        **DEFAULT_MODEL_PARAMS["attention_params"],
        "attention_dim": 16,
        "attention_location_n_filters": 4,
        "attention_location_kernel_size": 7,
    },
}


def default_params() -> dict:
    return copy.deepcopy(DEFAULT_MODEL_PARAMS)


def small_params() -> dict:
    return copy.deepcopy(SMALL_MODEL_PARAMS)


def speaker_dim(cfg: dict) -> int:
    """Width the speaker vector adds to the encoder output (tacotron2nv.py:31-43)."""
    if cfg["speaker_emb_type"] == "static+linear":
        return cfg["speaker_embedding_dim_lin"]
    return cfg["speaker_embedding_dim"]


def memory_dim(cfg: dict) -> int:
    return cfg["encoder_embedding_dim"] + speaker_dim(cfg)


def rnn_dims(cfg: dict):
    """(attention-RNN hidden, decoder-RNN hidden) as the reference really builds them.

    Tacotron2NV passes ``decoder_rnn_dim`` into Decoder's ``attention_rnn_dim``
    slot and vice versa (tacotron2nv.py:53-54 vs decoder.py:81-82).
    """
    return cfg["decoder_rnn_dim"], cfg["attention_rnn_dim"]
