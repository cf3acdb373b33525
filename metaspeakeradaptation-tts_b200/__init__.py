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

__all__ = ["default_params", "small_params", "FlatLayout", "Tacotron2NV", "Tacotron2Loss", "innerloop_ctx", "EWC"]


def __getattr__(name):      # GPU-facing classes import lazily: the pure-Python pieces above work on a CPU-only box
    if name == "Tacotron2NV":
        from .model import Tacotron2NV
        return Tacotron2NV
    if name == "Tacotron2Loss":
        from .loss import Tacotron2Loss
        return Tacotron2Loss
    if name == "innerloop_ctx":
        from .innerloop import innerloop_ctx
        return innerloop_ctx
    if name == "EWC":
        from .ewc import EWC
        return EWC
    raise AttributeError(name)
