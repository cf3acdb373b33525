"""Generate tests/golden/collate.npz from the UNMODIFIED reference collators (run in the build container only).

    python oracle/gen_golden_collate.py            # needs /root/reference

``MetaCollator`` (msa_tts/dataloaders/dataloader_meta.py:124-243) and ``Collator`` (dataloader_default.py:109-220) are imported
as they are.  Their module imports the audio front end (librosa, torchaudio's removed sox backend, espeak-based g2p), which is
out of scope and absent here, so those three modules are stubbed in ``sys.modules`` BEFORE the import and the collators get a
pass-through ``audio_processor`` whose ``get_melspec(w)`` returns ``w`` itself: the items carry a ready mel-spectrogram
[1, n_mels, len] in the waveform slot.  Everything the hot path depends on -- sorting by transcript length, zero / one padding,
padding to a multiple of the reduction factor, stop targets, dtypes -- is then the reference's own code.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

for name, attrs in (("msa_tts.utils.g2p.grapheme2phoneme", ["Grapheme2Phoneme"]), ("msa_tts.utils.ap", ["AudioProcessor"]),
                    ("msa_tts.utils.ap2", ["AudioProcessor2"])):
    mod = types.ModuleType(name)
    for a in attrs:
        setattr(mod, a, type(a, (), {}))
    sys.modules[name] = mod

from msa_tts.dataloaders.dataloader_default import Collator as RefCollator          # noqa: E402  (reference)
from msa_tts.dataloaders.dataloader_meta import MetaCollator as RefMetaCollator      # noqa: E402  (reference)

from oracle.gen_cases import collate_items                                           # noqa: E402


class _PassThrough:
    def get_melspec(self, w):
        return None, None, w


def _ref(cls, r):
    c = object.__new__(cls)             # the constructor only builds the audio processor
    c.reduction_factor = r
    c.audio_processor = _PassThrough()
    return c


def main():
    out = {}
    for r in (1, 2, 3):
        items = collate_items(seed=10 + r, n=5)
        b = _ref(RefCollator, r)(items)
        out[f"default_r{r}/item_ids"] = np.array(b[0])
        for k, t in zip(("transcripts", "trans_lengths", "melspecs", "melspec_lengths", "speaker_ids", "spk_embs", "stop_targets"), b[1:]):
            out[f"default_r{r}/{k}"] = t.numpy()
            out[f"default_r{r}/{k}/dtype"] = np.array(str(t.dtype))
        meta = [("spkA", {"train": collate_items(20 + r, 4), "test": collate_items(30 + r, 3)}),
                ("spkB", {"train": collate_items(40 + r, 2), "test": collate_items(50 + r, 4)})]
        d = _ref(RefMetaCollator, r)(meta)
        assert list(d.keys()) == ["spkA", "spkB"]
        for spk in d:
            for mode in ("train", "test"):
                b = d[spk][mode]
                out[f"meta_r{r}/{spk}/{mode}/item_ids"] = np.array(b[0])
                for k, t in zip(("transcripts", "trans_lengths", "melspecs", "melspec_lengths", "speaker_ids", "spk_embs", "stop_targets"), b[1:]):
                    out[f"meta_r{r}/{spk}/{mode}/{k}"] = t.numpy()
    path = os.path.join(ROOT, "tests", "golden", "collate.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
