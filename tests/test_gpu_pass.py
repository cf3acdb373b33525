"""Parity of the CUDA teacher-forced pass (through the C ABI) against the oracle and the golden fixtures.

Tolerances: fp32 GEMM mode -- outputs 2e-4 relative per tensor, gradients 2e-4 of the global gradient
norm; TF32 mode -- 1e-3 / 2e-3 (north_star: rel 1e-3 with a TF32 path)."""
import os

import numpy as np
import pytest
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from oracle.gen_cases import CASES
from helpers import cuda_pass, intermediates_report, oracle_pass, rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CUDA_TRAIN_CASES = ["small_train", "small_train_meanloss", "small_train_spklin", "small_train_lookup", "small_train_sigmoid",
                    "small_train_fwdattn_sigmoid",
                    # tacotron2nv.py:88-121: use_residual_encoder, freeze_charemb (+ residual), freeze_encoder, freeze_decoder -- the
                    # detached sub-graphs of the reference leave zero gradients (its p.grad stays None)
                    "small_train_residual", "small_train_freeze_charemb", "small_train_freeze_encoder", "small_train_freeze_decoder"]


def _engine(cfg, crit, tf32=0):
    from msa_tts_b200.engine import Engine
    return Engine(cfg, reduction=crit["reduction"], pos_weight=crit["pos_weight"], gemm_tf32=tf32)


def _check(eng, cfg, seed, dims, crit, tol_out, tol_grad, gold=None, report=None):
    B, T, L = dims
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    o_out, o_loss, o_grads, o_stats, inter = oracle_pass(cfg, P, batch, masks, crit)
    c_out, c_loss, c_grads, c_bn = cuda_pass(eng, cfg, P, batch, masks)
    lines = [f"{k:14s} {e:.3e}" for k, e in intermediates_report(eng, cfg, inter, B, T, L)]
    errs = {}
    for key, a, b in zip(("mel", "mel_post", "gate", "align"), c_out, o_out):
        errs[key] = rel(a, b)
    errs["loss"] = abs(float(c_loss) - float(o_loss)) / abs(float(o_loss))
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in o_grads.values())))
    gerr = {n: float((c_grads[n].double().cpu() - o_grads[n].double()).norm()) / gn for n in o_grads}
    nan = [n for n in c_grads if not bool(torch.isfinite(c_grads[n]).all())]
    serr = {k: rel(c_bn[k], o_stats[k]) for k in c_bn}
    lines += [f"out/{k:10s} {v:.3e}" for k, v in errs.items()]
    lines += [f"grad/{n:70s} {v:.3e}  (|g|/|G| {float(o_grads[n].norm()) / gn:.2e})" for n, v in gerr.items()]
    lines += [f"bn/{k:60s} {v:.3e}" for k, v in serr.items()]
    txt = "\n".join(lines)
    if report:
        os.makedirs(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out"), exist_ok=True)
        with open(os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", report), "w") as f:
            f.write(txt + "\n")
    print(txt)
    assert not nan, f"non-finite / unwritten gradients: {nan}"
    assert max(errs.values()) < tol_out, errs
    assert max(gerr.values()) < tol_grad, max(gerr.items(), key=lambda kv: kv[1])
    assert max(serr.values()) < tol_out, serr
    if gold is not None:   # and against the reference's own outputs
        for key, a in zip(("mel", "mel_post", "gate"), c_out):
            assert rel(a, gold[key]) < tol_out, ("golden", key)
        assert abs(float(c_loss) - float(gold["loss"])) < tol_out * abs(float(gold["loss"]))
    return c_out, c_loss, c_grads


@pytest.mark.parametrize("name", CUDA_TRAIN_CASES)
def test_small_cases_fp32(name):
    cfg, seed, dims, crit = CASES[name]()
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    eng = _engine(cfg, crit)
    c_out, c_loss, c_grads = _check(eng, cfg, seed, dims, crit, 2e-4, 2e-4, gold, report=f"parity_{name}.txt")
    gn = np.sqrt(sum(float((gold["grad/" + n].astype(np.float64) ** 2).sum()) for n in c_grads))
    for n, g in c_grads.items():
        assert float((g.double().cpu() - torch.as_tensor(gold["grad/" + n]).double()).norm()) / gn < 2e-4, ("golden grad", n)


@pytest.mark.parametrize("attn", [dict(forward_attn=True), dict(forward_attn=True, trans_agent=True),
                                  dict(forward_attn=True, trans_agent=True, norm="sigmoid")])
def test_forward_attention_variants_fp32(attn):
    """Forward attention in the training chain kernels (recursion alpha' = ((1-u) alpha + u shift(alpha) + 1e-8) a, transition
    agent u = sigmoid(W_ta [ctx; h] + b), forward_attn.py:154-176,222-224) incl. their backward, ragged sizes, vs the oracle."""
    cfg = pkg.small_params()
    cfg["attention_params"].update(attn)
    crit = dict(reduction="none", pos_weight=10.0)
    _check(_engine(cfg, crit), cfg, 33 + len(attn), (5, 17, 13), crit, 2e-4, 2e-4, report=f"parity_fwdattn_{len(attn)}.txt")


def test_ragged_odd_sizes_fp32():
    cfg = pkg.small_params()
    crit = dict(reduction="none", pos_weight=10.0)
    _check(_engine(cfg, crit), cfg, 31, (5, 17, 13), crit, 2e-4, 2e-4, report="parity_ragged.txt")


def test_batch_of_one_row_lengths():
    """B=2 with very different lengths (packed BiLSTM semantics, padded positions in BN statistics, Q6)."""
    cfg = pkg.small_params()
    crit = dict(reduction="none", pos_weight=10.0)
    eng = _engine(cfg, crit)
    B, T, L = 2, 9, 11
    P = synth.init_params(cfg, 41)
    batch = list(synth.make_batch(cfg, B, T, L, 141))
    batch[2] = torch.tensor([11, 3]); batch[1][1, 3:] = 0
    batch[4] = torch.tensor([9, 4]); batch[3][1, :, 4:] = 0.0
    batch[7][1] = 0.0; batch[7][1, 3:] = 1.0
    masks = synth.make_masks(cfg, B, T, L, 241)
    o_out, o_loss, o_grads, _, _ = oracle_pass(cfg, P, tuple(batch), masks, crit)
    c_out, c_loss, c_grads, _ = cuda_pass(eng, cfg, P, tuple(batch), masks)
    for a, b in zip(c_out, o_out):
        assert rel(a, b) < 2e-4
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in o_grads.values())))
    for n in o_grads:
        assert float((c_grads[n].double().cpu() - o_grads[n].double()).norm()) / gn < 2e-4, n


def test_default_dims_config1_fp32_and_tf32():
    """BASELINE.json configs[0] shapes (B=4, T=200, L=64, 80 mels, default dims) against the committed golden
    outputs of the reference, fp32 GEMMs (tight) and TF32 GEMMs (north_star tolerance)."""
    name = "default_train_b4_t200"
    cfg, seed, dims, crit = CASES[name]()
    z = np.load(os.path.join(GOLD, name + ".npz"))
    B, T, L = dims
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    gn = float(np.sqrt((z["grad_norms"] ** 2).sum()))
    names = list(P.keys())
    # GEMM policy 0: fp32 everywhere; 1: fp32 forward + TF32 backward (bench default); 2: TF32 everywhere (reported, out of
    # the 1e-3 output tolerance because the postnet amplifies decoder-output error ~5x)
    # tol_g bounds BOTH the per-tensor norm error and the sampled-element error (direction / layout), each relative to the global
    # gradient norm scale; policy 1 is the bench policy and is held to north_star's 1e-3, policy 2 is reported only
    for tf32, tol_o, tol_g, tol_s in ((0, 3e-4, 3e-4, 3e-4), (1, 3e-4, 1e-3, 1e-3), (2, 5e-3, 2e-3, 3e-2)):
        eng = _engine(cfg, crit, tf32)
        c_out, c_loss, c_grads, c_bn = cuda_pass(eng, cfg, P, batch, masks)
        lines = []
        for key, a in zip(("mel", "mel_post", "gate"), c_out):
            lines.append(f"tf32={tf32} out/{key} {rel(a, z[key]):.3e}")
        lines.append(f"tf32={tf32} out/align {rel(c_out[3][:, ::8], z['align_sample']):.3e}")
        lines.append(f"tf32={tf32} loss {abs(float(c_loss) - float(z['loss'])) / abs(float(z['loss'])):.3e}")
        worst, worst_s = 0.0, 0.0
        for i, n in enumerate(names):
            g = c_grads[n]
            flat = g.flatten()
            samp = flat[:: max(1, flat.numel() // 64)][:64].cpu()
            e_norm = abs(float(g.double().norm()) - z["grad_norms"][i]) / gn
            e_samp = float((samp.double() - torch.as_tensor(z["gsample/" + n]).double()).norm()) / (float(np.linalg.norm(z["gsample/" + n])) + 1e-3 * gn)
            worst = max(worst, e_norm)
            worst_s = max(worst_s, e_samp)
            lines.append(f"tf32={tf32} grad/{n:70s} norm-err {e_norm:.3e} sample-rel {e_samp:.3e}")
        txt = "\n".join(lines)
        print(txt)
        odir = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
        os.makedirs(odir, exist_ok=True)
        with open(os.path.join(odir, f"parity_default_tf32_{int(tf32)}.txt"), "w") as f:
            f.write(txt + "\n")
        for key, a in zip(("mel", "mel_post", "gate"), c_out):
            assert rel(a, z[key]) < tol_o, (tf32, key)
        assert abs(float(c_loss) - float(z["loss"])) < tol_o * abs(float(z["loss"]))
        assert worst < tol_g, (tf32, worst)
        assert worst_s < tol_s, (tf32, worst_s)
        del eng
        torch.cuda.empty_cache()


def test_backward_accumulate_and_scale():
    cfg, seed, dims, crit = CASES["small_train"]()
    B, T, L = dims
    from msa_tts_b200.engine import batch_to_device
    eng = _engine(cfg, crit)
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, T, L, seed + 100)
    masks = synth.make_masks(cfg, B, T, L, seed + 200)
    flat = eng.flat_from_dict(P)
    bd = batch_to_device(batch, eng.device)
    mflat = eng.pack_masks(masks, B, T, L)
    g1, g2 = eng.new_flat(0.0), eng.new_flat(0.0)
    eng.forward(flat, None, bd, mflat)
    eng.backward(flat, g1)
    eng.forward(flat, None, bd, mflat)
    eng.backward(flat, g2, accumulate=False, scale=0.5)
    eng.backward(flat, g2, accumulate=True, scale=0.5)
    assert rel(g2, g1) < 1e-6
    g3 = eng.new_flat(0.0)
    eng.forward(flat, None, bd, mflat)
    eng.backward(flat, g3)
    assert torch.equal(g3, g1), "the pass must be deterministic (idempotence)"


@pytest.mark.parametrize("dims", [(2, 24, 150), (8, 16, 64), (1, 30, 64)])
def test_default_dims_other_shapes_fp32(dims):
    """Default model dimensions away from the bench shape: a long text (the MW slices no longer fit in shared memory: the
    global-memory fallbacks of the attention chain run), bigger batches (several batch tiles), a single row."""
    cfg = pkg.default_params()
    crit = dict(reduction="none", pos_weight=10.0)
    B, T, L = dims
    eng = _engine(cfg, crit)
    P = synth.init_params(cfg, 3)
    batch = synth.make_batch(cfg, B, T, L, 77)
    masks = synth.make_masks(cfg, B, T, L, 78)
    o_out, o_loss, o_grads, _, _ = oracle_pass(cfg, P, batch, masks, crit)
    c_out, c_loss, c_grads, _ = cuda_pass(eng, cfg, P, batch, masks)
    for key, a, b in zip(("mel", "mel_post", "gate", "align"), c_out, o_out):
        assert rel(a, b) < 3e-4, (key, rel(a, b))
    assert abs(float(c_loss) - float(o_loss)) < 3e-4 * abs(float(o_loss))
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in o_grads.values())))
    worst = max((float((c_grads[n].double().cpu() - o_grads[n].double()).norm()) / gn, n) for n in o_grads)
    assert worst[0] < 3e-4, worst


@pytest.mark.parametrize("dims", [(16, 10, 24), (32, 8, 16), (18, 6, 64)])
def test_default_dims_large_batches_run_on_the_grouped_tensor_core_kernels(dims):
    """Batches the single-task recurrent kernels cannot hold at the default dimensions (B > 8: their shared-memory staging of h)
    run on the grouped kernels (chain_mma.cu: h streamed from L2 straight into tensor-core fragments, bf16x3 products) under
    every GEMM policy.  B = 16, 18 and 32 against the oracle, fp32 GEMMs."""
    cfg = pkg.default_params()
    crit = dict(reduction="none", pos_weight=10.0)
    B, T, L = dims
    eng = _engine(cfg, crit)
    P = synth.init_params(cfg, 3)
    batch = synth.make_batch(cfg, B, T, L, 77)
    masks = synth.make_masks(cfg, B, T, L, 78)
    o_out, o_loss, o_grads, _, _ = oracle_pass(cfg, P, batch, masks, crit)
    c_out, c_loss, c_grads, _ = cuda_pass(eng, cfg, P, batch, masks)
    for key, a, b in zip(("mel", "mel_post", "gate", "align"), c_out, o_out):
        assert rel(a, b) < 3e-4, (key, rel(a, b))
    assert abs(float(c_loss) - float(o_loss)) < 3e-4 * abs(float(o_loss))
    gn = float(torch.sqrt(sum((g.double() ** 2).sum() for g in o_grads.values())))
    worst = max((float((c_grads[n].double().cpu() - o_grads[n].double()).norm()) / gn, n) for n in o_grads)
    print("worst gradient error / |G|:", worst)
    assert worst[0] < 3e-4, worst


def test_batch_beyond_every_kernel_is_a_loud_error():
    """More than 32 rows per recurrence launch fit neither kernel family: the library says so instead of computing something else."""
    cfg = pkg.default_params()
    crit = dict(reduction="none", pos_weight=10.0)
    eng = _engine(cfg, crit)
    B, T, L = 40, 4, 16
    P = synth.init_params(cfg, 3)
    with pytest.raises(RuntimeError, match="batch|shared memory|rows"):
        cuda_pass(eng, cfg, P, synth.make_batch(cfg, B, T, L, 77), synth.make_masks(cfg, B, T, L, 78), backward=False)
