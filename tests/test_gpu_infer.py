"""Parity of the CUDA free-running inference path (msa_infer through the C ABI) against the committed golden outputs of the
reference (Tacotron2NV.infer, tacotron2nv.py:130-162) and the oracle restatement.

Tolerances: mel_post / alignments 5e-4 relative (fp32; the loop is autoregressive, so rounding differences of the GEMM
summation order feed back through <= 24 steps); mel_lengths and the number of produced steps bit-exact."""
import os

import numpy as np
import pytest
import torch

import msa_tts_b200 as pkg
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


@pytest.mark.parametrize("name", ["small_infer", "small_infer_earlystop", "small_infer_residual"])
def test_infer_small(name):
    (post, lens, align), (o_post, o_lens, o_align), z = _run(name)
    T = int(z["steps"])
    assert post.shape[2] == T == o_post.shape[2], "number of decoder steps must match the reference exactly"
    assert torch.equal(lens.cpu(), o_lens) and np.array_equal(lens.cpu().numpy(), z["mel_lengths"]), "mel_lengths are bit-exact"
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL
    assert rel(post, z["mel_post"][:, :, :T]) < TOL and rel(align, z["align"][:, :T]) < TOL


def test_infer_window_fwdmask_transagent_golden():
    """Eval-only attention features on the device: windowing driven by batch row 0, forward attention with the transition agent
    and the forward-attention mask with the reference's Python slice / negative-index semantics (forward_attn.py:139-176,222-224,
    SURVEY Q16) -- against the golden output of the unmodified reference."""
    (post, lens, align), (o_post, o_lens, o_align), z = _run("small_infer_window_fwdmask")
    T = int(z["steps"])
    assert post.shape[2] == T == o_post.shape[2]
    assert torch.equal(lens.cpu(), o_lens) and np.array_equal(lens.cpu().numpy(), z["mel_lengths"])
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL
    assert rel(post, z["mel_post"][:, :, :T]) < TOL and rel(align, z["align"][:, :T]) < TOL
    # alignment-index handling is exact: the masked-out positions are exactly zero on both sides
    assert torch.equal(align.cpu() == 0, o_align == 0)


VARIANTS = {
    "fwd": dict(forward_attn=True),
    "fwd_ta": dict(forward_attn=True, trans_agent=True),
    "fwd_mask": dict(forward_attn=True, forward_attn_mask=True),
    "fwd_ta_mask_sigmoid": dict(forward_attn=True, trans_agent=True, forward_attn_mask=True, norm="sigmoid"),
    "window": dict(windowing=True),
    "window_sigmoid": dict(windowing=True, norm="sigmoid"),
    "window_fwd": dict(windowing=True, forward_attn=True),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
@pytest.mark.parametrize("B,L", [(3, 14), (5, 37)])
def test_infer_attention_variants_vs_oracle(name, B, L):
    """Every combination of the eval-time attention switches against the oracle restatement (no golden: same arithmetic, pinned
    through small_infer_window_fwdmask), incl. a text length that spans more than one warp pass and uneven cluster shares."""
    from msa_tts_b200.engine import Engine
    from oracle.gen_cases import _infer
    cfg, seed, steps = _infer(**VARIANTS[name]), 40 + len(name), 16
    cfg["max_decoder_steps"] = steps
    eng = Engine(cfg)
    P = synth.init_params(cfg, seed)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, seed + 100)
    stats = infer_stats(P, cfg, seed)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL
    assert torch.equal(align.cpu() == 0, o_align == 0)


@pytest.mark.parametrize("B,L,attn", [(32, 64, {}), (5, 61, dict(forward_attn=True, trans_agent=True, norm="sigmoid"))])
def test_infer_default_dims_vs_oracle(B, L, attn):
    """Default (Tacotron-2) dimensions: the multi-chunk K ring, 7 hidden units per CTA on all SMs, the two-segment LSTM products,
    the 81-row projection with its second weight matrix and the 4-CTA attention clusters -- none of which the small model
    exercises -- against the oracle for a few free-running steps (BASELINE configs[4] batch shape and a ragged one)."""
    from msa_tts_b200.engine import Engine
    steps = 6
    cfg = pkg.default_params()
    cfg["attention_params"].update(attn)
    cfg["max_decoder_steps"] = steps
    cfg["decoder_no_early_stopping"] = True
    eng = Engine(cfg)
    P = synth.init_params(cfg, 3)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, 77)
    stats = infer_stats(P, cfg, 3)
    pm = synth.make_infer_masks(cfg, B, steps, 78)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL, (rel(post, o_post), rel(align, o_align))


def test_infer_more_than_32_rows_is_decoded_tile_by_tile():
    """One msa_infer call takes one 32-row tile (BASELINE configs[4] is B = 32); the host layer decodes larger batches tile by tile
    (eval mode has no cross-row coupling): mel_lengths bit-exact and every frame below a row's length equal to the oracle's run of
    the whole batch."""
    from msa_tts_b200.engine import Engine
    from oracle.gen_cases import _infer
    cfg, steps, B = _infer(early=True, thr=0.62), 12, 40
    cfg["max_decoder_steps"] = steps
    eng = Engine(cfg)
    P = synth.init_params(cfg, 5)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, 9, 105)
    stats = infer_stats(P, cfg, 5)
    pm = synth.make_infer_masks(cfg, B, steps, 305)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert torch.equal(lens.cpu(), o_lens)
    assert post.shape[0] == B and post.shape[2] <= o_post.shape[2]
    for b in range(B):
        n = min(int(o_lens[b]), post.shape[2])
        assert rel(post[b, :, :n], o_post[b, :, :n]) < TOL and rel(align[b, :n], o_align[b, :n]) < TOL, b
    # the raw C entry point still refuses more than one tile loudly
    with pytest.raises(RuntimeError, match="batch 40"):
        Engine.INFER_ROWS, keep = 64, Engine.INFER_ROWS
        try:
            eng.infer(eng.flat_from_dict(P), eng.new_bn_stats(), inp, inp_len, spk, pm, max_steps=steps)
        finally:
            Engine.INFER_ROWS = keep


@pytest.mark.parametrize("B,L,early", [(32, 11, False), (17, 23, True), (1, 7, False)])
def test_infer_batch_shapes_small_model(B, L, early):
    """A full batch tile, a ragged one with early stopping, and a single row -- small model, vs the oracle."""
    from msa_tts_b200.engine import Engine
    from oracle.gen_cases import _infer
    cfg, seed, steps = _infer(early=early, thr=0.62), 61, 20
    cfg["max_decoder_steps"] = steps
    eng = Engine(cfg)
    P = synth.init_params(cfg, seed)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, seed + 100)
    stats = infer_stats(P, cfg, seed)
    pm = synth.make_infer_masks(cfg, B, steps, seed + 300)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL


def test_infer_long_horizon_default_dims():
    """150 free-running steps at the default dimensions: the tensor-core products of the decoder step are 3xTF32 (fp32-accurate),
    so the error against the fp32 oracle must not grow with the horizon (a single-TF32 step would drift beyond the tolerance)."""
    from msa_tts_b200.engine import Engine
    B, L, steps = 2, 40, 150
    cfg = pkg.default_params()
    cfg["max_decoder_steps"] = steps
    cfg["decoder_no_early_stopping"] = True
    eng = Engine(cfg)
    P = synth.init_params(cfg, 8)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, 91)
    stats = infer_stats(P, cfg, 8)
    pm = synth.make_infer_masks(cfg, B, steps, 92)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    first, last = slice(0, 30), slice(steps - 30, steps)
    e_first, e_last = rel(post[:, :, first], o_post[:, :, first]), rel(post[:, :, last], o_post[:, :, last])
    print(f"long horizon: rel err first 30 steps {e_first:.2e}, last 30 steps {e_last:.2e}, alignments {rel(align, o_align):.2e}")
    assert rel(post, o_post) < TOL and rel(align, o_align) < TOL and e_last < TOL


def test_infer_tf32_policy_within_stated_tolerance():
    """GEMM policy 2 ("TF32 everywhere"): the decoder LSTMCells use plain TF32 products.  north_star tolerance for a TF32 path:
    rel 1e-3 on the mel outputs; mel_lengths / number of steps stay bit-exact."""
    from msa_tts_b200.engine import Engine
    B, L, steps = 4, 48, 60
    cfg = pkg.default_params()
    cfg["max_decoder_steps"] = steps
    cfg["decoder_no_early_stopping"] = True
    eng = Engine(cfg, gemm_tf32=2)
    P = synth.init_params(cfg, 9)
    _, inp, inp_len, _, _, _, spk, _ = synth.make_batch(cfg, B, 8, L, 93)
    stats = infer_stats(P, cfg, 9)
    pm = synth.make_infer_masks(cfg, B, steps, 94)
    post, lens, align = eng.infer(eng.flat_from_dict(P), eng.bn_from_dict(stats), inp, inp_len, spk, pm, max_steps=steps)
    torch.cuda.synchronize()
    o_post, o_lens, o_align = OM.infer(P, cfg, inp, inp_len, spk, pm, stats)
    assert post.shape == o_post.shape and torch.equal(lens.cpu(), o_lens)
    e = rel(post, o_post)
    print(f"tf32 policy: rel err mel_post {e:.2e}, alignments {rel(align, o_align):.2e}")
    assert 1e-6 < e < 1e-3 and rel(align, o_align) < 1e-3      # (> 1e-6: the TF32 path really ran)


def test_infer_graph_and_direct_launch_agree():
    """The CUDA-graph replay of the decoder step and plain per-step launches are the same computation."""
    (post, lens, align), _, _ = _run("small_infer")
    os.environ["MSA_INFER_GRAPH"] = "1"
    try:
        (post2, lens2, align2), _, _ = _run("small_infer")
    finally:
        del os.environ["MSA_INFER_GRAPH"]
    assert torch.equal(post, post2) and torch.equal(lens, lens2) and torch.equal(align, align2)


def test_vocoder_hand_off_like_infer_py():
    """infer.py:311-328 hands each utterance's mel to the vocoder as a [1, n_mel, len] tensor on the vocoder's device: here the
    inference output stays on the GPU and is sliced per utterance (views), or packed as one padded batch."""
    from msa_tts_b200.vocoder import vocoder_batch, vocoder_inputs
    (post, lens, _), _, z = _run("small_infer_earlystop")
    items = vocoder_inputs(post, lens)
    assert len(items) == post.shape[0]
    for b, x in enumerate(items):
        n = min(int(z["mel_lengths"][b]), post.shape[2])
        assert x.is_cuda and tuple(x.shape) == (1, post.shape[1], n)
        assert x.data_ptr() == post[b:b + 1].data_ptr(), "a view of the inference output, no copy"
        assert rel(x[0], z["mel_post"][b, :, :n]) < TOL
    batch = vocoder_batch(post, lens, pad_value=-7.0)
    for b in range(post.shape[0]):
        n = min(int(z["mel_lengths"][b]), batch.shape[2])
        assert torch.equal(batch[b, :, :n], post[b, :, :n]) and bool((batch[b, :, n:] == -7.0).all())
