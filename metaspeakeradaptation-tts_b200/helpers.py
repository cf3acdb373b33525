"""helpers.get_optimizer semantics (msa_tts/utils/helpers.py:20-26): the YAML gives an optimizer class name and
hyper-parameters as *strings* that the reference passes through eval()."""
from __future__ import annotations

import ast


def optimizer_hparams(spec: dict) -> dict:
    """{"optimizer_name": "SGD", "optim_params": {"lr": "1e-3", ...}} -> {"name": "SGD", "lr": 0.001, ...}.

    Strings are evaluated like the reference does, but with ast.literal_eval (numbers, tuples, booleans)."""
    out = {"name": spec["optimizer_name"]}
    for k, v in spec.get("optim_params", {}).items():
        out[k] = ast.literal_eval(v) if isinstance(v, str) else v
    if "lr" not in out:
        raise ValueError("optimizer needs an lr")
    return out
