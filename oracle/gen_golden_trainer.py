"""Generate tests/golden/trainer.npz from the UNMODIFIED reference trainer-level code (run in the build container only).

    python oracle/gen_golden_trainer.py            # needs /root/reference

What is imported from /root/reference as it is and run here:
  * ``mix_grad`` / ``apply_grad`` (msa_tts/utils/grad_utils.py:8-31) followed by the outer update exactly as
    maml.py:94-105 spells it -- ``model.zero_grad()``, uniform weights, ``clip_grad_norm_``, ``outer_optimizer.step()`` with the
    optimizer built by ``get_optimizer`` (utils/helpers.py:20-26) on the reference ``Tacotron2NV`` -- for SGD and for two
    consecutive Adam steps;
  * ``mcd_batch`` (msa_tts/utils/metrics.py:15-22), called as maml.py:78-80 calls it;
  * the ``EWC`` class (msa_tts/continual_ewc.py:28-89): ``_diag_fisher`` over a buffer of batches, the means snapshot,
    ``penalty`` at perturbed weights, and the BatchNorm running statistics the Fisher passes leave in the model.
    continual_ewc.py imports third-party / out-of-scope modules that are absent here (higher, tensorboard, matplotlib, the
    audio front end); they are stubbed in ``sys.modules`` BEFORE the import -- none of them is touched by the EWC class.
``higher`` itself (the inner-loop optimizer) is absent and stays "parity unpinned" (oracle/meta.py).

The fixture holds the reference's results; this script also asserts that the restatement in oracle/meta.py reproduces them.
Inputs are rebuilt in the tests from the same seeds (``trainer_inputs`` below), so nothing of the reference travels.
"""
from __future__ import annotations

import copy
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import msa_tts_b200 as pkg                      # noqa: E402
from msa_tts_b200 import synth                  # noqa: E402
from oracle import meta as OMeta                # noqa: E402
from oracle import model as OM                  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_TASKS, CLIP, LR_SGD, LR_ADAM = 3, 0.5, 0.05, 0.01
EWC_BATCHES, EWC_DIMS = 3, (3, 10, 8)
CRIT = dict(reduction="none", pos_weight=10.0)


def trainer_inputs():
    """Seeded inputs shared by this generator and the tests: small model, three task-gradient lists, two rounds (the second
    round feeds the second Adam step), arrays for mcd_batch, EWC buffer batches with their dropout masks."""
    cfg = pkg.small_params()
    P = synth.init_params(cfg, 51)
    names = list(P.keys())
    rounds = []
    for r in range(2):
        gl = []
        for i in range(N_TASKS):
            g = torch.Generator().manual_seed(700 + 10 * r + i)
            gl.append({n: 0.3 * torch.randn(P[n].shape, generator=g) for n in names})
        rounds.append(gl)
    g = torch.Generator().manual_seed(77)
    B, T, D = 4, 13, cfg["n_mel_channels"]
    mcd_in = (torch.randn(B, D, T, generator=g), torch.randn(B, D, T, generator=g), torch.tensor([13, 11, 7, 2]))
    B, T, L = EWC_DIMS
    buf = [synth.make_batch(cfg, B, T, L, 900 + i) for i in range(EWC_BATCHES)]
    buf_masks = [synth.make_masks(cfg, B, T, L, 950 + i) for i in range(EWC_BATCHES)]
    gp = torch.Generator().manual_seed(78)
    P_moved = {n: P[n] + 0.02 * torch.randn(P[n].shape, generator=gp) for n in names}
    return cfg, P, names, rounds, mcd_in, buf, buf_masks, P_moved


def _stub_modules():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    stub("higher")
    stub("matplotlib")
    stub("matplotlib.pyplot")
    stub("librosa")
    stub("torch.utils.tensorboard", SummaryWriter=type("SummaryWriter", (), {}))
    for name, attrs in (("msa_tts.utils.g2p.grapheme2phoneme", ["Grapheme2Phoneme"]), ("msa_tts.utils.ap", ["AudioProcessor"]),
                        ("msa_tts.utils.ap2", ["AudioProcessor2"]), ("msa_tts.utils.plot", ["plot_spec_attn_example"])):
        stub(name, **{a: type(a, (), {}) for a in attrs})


def main():
    sys.path.insert(0, "/root/reference")
    _stub_modules()
    from oracle.gen_golden import DropoutQueue, build_ref, train_mask_sequence, rel            # noqa: E402 (imports the reference model)
    from msa_tts.utils.grad_utils import apply_grad, mix_grad                                   # noqa: E402  (reference)
    from msa_tts.utils.helpers import get_optimizer                                             # noqa: E402  (reference)
    from msa_tts.utils.metrics import mcd_batch                                                 # noqa: E402  (reference)
    from msa_tts.continual_ewc import EWC                                                       # noqa: E402  (reference)
    from msa_tts.models.modules_tacotron2nv.tacotron2nv_loss import Tacotron2Loss               # noqa: E402  (reference)
    from torch.nn.utils import clip_grad_norm_

    torch.set_num_threads(8)
    cfg, P, names, rounds, mcd_in, buf, buf_masks, P_moved = trainer_inputs()
    d = {}

    # ---- A. outer update, maml.py:94-105 -----------------------------------------------------------------
    for opt_name, lr, n_rounds in (("SGD", LR_SGD, 1), ("Adam", LR_ADAM, 2)):
        model = build_ref(cfg, P)
        opt = get_optimizer(model, optimizer_name=opt_name, optim_params={"lr": repr(lr)})
        state = {}
        Po = {n: P[n].clone() for n in names}
        for r in range(n_rounds):
            grad_list = [[gl[n].clone() for n in names] for gl in rounds[r]]
            model.zero_grad()
            weight = torch.ones(len(grad_list))
            weight = weight / torch.sum(weight)
            mixed = mix_grad(grad_list, weight)
            # the restatement (oracle/meta.py) must give the same numbers (checked here: apply_grad assigns these very tensors
            # to p.grad and clip_grad_norm_ then scales them in place)
            o_mixed = OMeta.mix_grad(rounds[r], weight, names)
            assert max(rel(o_mixed[n], m) for n, m in zip(names, mixed)) < 1e-6
            mixed_copy = [m.clone() for m in mixed]
            grad_log = apply_grad(model, mixed)
            total = clip_grad_norm_(model.parameters(), CLIP)
            opt.step()
            assert abs(OMeta.grad_norm(o_mixed, names) - grad_log) < 1e-5 * grad_log
            Po = (OMeta.outer_sgd(Po, o_mixed, names, lr, clip=CLIP) if opt_name == "SGD"
                  else OMeta.outer_adam(Po, o_mixed, names, state, lr, clip=CLIP))
            sd = {n: p.detach().clone() for n, p in model.named_parameters()}
            e = max(rel(Po[n], sd[n]) for n in names)
            print(f"[outer {opt_name} step {r + 1}] grad norm {grad_log:.6f} (clip at {CLIP}, total {float(total):.6f})  "
                  f"oracle vs reference weights {e:.2e}")
            assert e < 2e-6
            if r == 0 and opt_name == "SGD":
                for n, m in zip(names, mixed_copy):
                    d["mixed/" + n] = m.numpy()
                d["grad_norm"] = np.float64(grad_log)
            for n in names:
                d[f"{opt_name.lower()}{r + 1}/" + n] = sd[n].numpy()

    # ---- B. mcd_batch as maml.py:78-80 calls it ----------------------------------------------------------
    out, mel, mel_len = mcd_in
    v = mcd_batch(out.transpose(1, 2).numpy(), mel.transpose(1, 2).numpy(), mel_len.numpy())
    o = OMeta.mcd_batch(out.transpose(1, 2), mel.transpose(1, 2), mel_len)
    print(f"[mcd_batch] reference {v:.6f}  oracle {o:.6f}")
    assert abs(v - o) < 1e-5 * abs(v)
    d["mcd"] = np.float64(v)

    # ---- C. EWC (continual_ewc.py:28-89) -----------------------------------------------------------------
    model = build_ref(cfg, P)
    crit = Tacotron2Loss(cfg["n_frames_per_step"], CRIT["reduction"], CRIT["pos_weight"], "cpu")
    seq = []
    for b, m in zip(buf, buf_masks):
        seq += train_mask_sequence(cfg, m, b[3].shape[2])
    q = DropoutQueue(seq)
    orig = torch.nn.functional.dropout
    torch.nn.functional.dropout = q
    try:
        ewc = EWC(model, buf, crit, "cpu")
    finally:
        torch.nn.functional.dropout = orig
    assert q.i == len(q.seq)
    o_F = OMeta.ewc_fisher(P, cfg, buf, buf_masks, CRIT, names)
    eF = max(float((o_F[n].double() - ewc._precision_matrices[n].double()).norm()) for n in names) / \
        float(torch.sqrt(sum((ewc._precision_matrices[n].double() ** 2).sum() for n in names)))
    moved = build_ref(cfg, P_moved)
    pen = float(ewc.penalty(moved))
    o_pen = float(OMeta.ewc_penalty(P_moved, o_F, P, names))
    print(f"[EWC] fisher oracle vs reference {eF:.2e}  penalty reference {pen:.6e} oracle {o_pen:.6e}")
    assert eF < 2e-5 and abs(pen - o_pen) < 2e-5 * abs(pen)
    for n in names:
        d["fisher/" + n] = ewc._precision_matrices[n].numpy()
        assert torch.equal(ewc._means[n], P[n])
    d["penalty"] = np.float64(pen)
    for k, v in model.state_dict().items():          # BN buffers of the model after the Fisher passes (they DO move)
        if "running" in k or "num_batches" in k:
            d["ewc_stat/" + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "trainer.npz"), **d)
    print("written", os.path.join(GOLD, "trainer.npz"))


if __name__ == "__main__":
    main()
