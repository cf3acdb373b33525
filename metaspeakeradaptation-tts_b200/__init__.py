"""B200-native hot path of MetaSpeakerAdaptation-TTS (meta-training inner loop of Tacotron2NV).

Import name: ``msa_tts_b200`` (the directory name required by the build spec,
``metaspeakeradaptation-tts_b200``, is not a valid Python identifier; the
``msa_tts_b200`` alias package at the repo root points its ``__path__`` here).

Pure-Python pieces (config, layout, synthetic data) import without a GPU.
Everything that computes goes through the C-ABI library ``libmsa_b200.so``
(include/msa_b200.h) and fails loudly when it is missing: there is no CPU or
PyTorch fallback.
"""
from .config import default_params, small_params  # noqa: F401
from .layout import FlatLayout  # noqa: F401

__all__ = ["default_params", "small_params", "FlatLayout"]
