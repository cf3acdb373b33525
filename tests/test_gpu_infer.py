"""Parity of the CUDA free-running inference path (msa_infer through the C ABI) against the committed golden outputs of the
reference (Tacotron2NV.infer, tacotron2nv.py:130-162) and the oracle restatement.

Tolerances: mel_post / alignments 5e-4 relative (fp32; the loop is autoregressive, so rounding differences of the GEMM
summation order feed back through <= 24 steps); mel_lengths and the number of produced steps bit-exact."""
import os

import numpy as np
import pytest
import torch

from msa_tts_b200 import synth
from oracle import model as OM
from oracle.gen_cases import INFER_CASES, infer_stats
from helpers import rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 5e-4


def _run(name):
    from msa_tts_b200.engine import Engine
    cfg, seed, (B, L), steps = INFER_CASES[name]()
    eng = Engine(cfg)
    P = synth.init_params(cfg, seed)
    batch = synth.make_batch(cfg, B, 8, L, seed + 100)
    _, inp, inp_len, _, _, _, spk, _ = batch
    stats = infer_stats(P, cfg, seed)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    flat = eng.flat_from_dict(P)
    bn = eng.bn_from_dict(stats)
    out = eng.infer(flat, bn, inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    ref = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    return out, ref, np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", ["small_infer", "small_infer_earlystop"])
def test_infer_small(name):
    (post, lens, align), (o_post, o_lens, o_align), z = _run(name)
    T = int(z["steps"])
    assert post.shape[2] == T == o_post.shape[2], "number of decoder steps must match the reference exactly"
    assert torch.equal(lens.cpu(), o_lens) and np.array_equal(lens.cpu().numpy(), z["mel_lengths"]), "mel_lengths are bit-exact"
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL
    assert rel(post, z["mel_post"][:, :, :T]) < TOL and rel(align, z["align"][:, :T]) < TOL


def test_infer_unsupported_variants_fail_loudly():
    from msa_tts_b200.engine import Engine
    cfg, seed, (B, L), steps = INFER_CASES["small_infer_window_fwdmask"]()
    eng = Engine(cfg)
    P = synth.init_params(cfg, seed)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, seed + 100)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    with pytest.raises(RuntimeError, match="not implemented"):
        eng.infer(eng.flat_from_dict(P), eng.new_bn_stats(), inp, inp_len, spk, pm, max_steps=steps)


def test_infer_graph_and_direct_launch_agree():
    """The CUDA-graph replay of the decoder step and plain per-step launches are the same computation."""
    (post, lens, align), _, _ = _run("small_infer")
    os.environ["MSA_INFER_GRAPH"] = "1"
    try:
        (post2, lens2, align2), _, _ = _run("small_infer")
    finally:
        del os.environ["MSA_INFER_GRAPH"]
    assert torch.equal(post, post2) and torch.equal(lens, lens2) and torch.equal(align, align2)
