"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes (default dimensions, B=4, T=200, L=64, 8 tasks),
where the CPU oracle is too slow to be the checker for every case: bitwise repeatability, linearity of the backward pass in the
upstream gradients, and "fused accumulation == mean of separately computed task gradients" for the whole FOMAML meta-step.
(The same shapes against the reference's golden outputs: tests/test_gpu_pass.py::test_default_dims_config1_fp32_and_tf32.)
"""
import pytest
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from helpers import rel

pytestmark = pytest.mark.gpu
B, T, L = 4, 200, 64


def _engine(tf32):
    from msa_tts_b200.engine import Engine
    return Engine(pkg.default_params(), reduction="none", pos_weight=10.0, gemm_tf32=tf32)


def _inputs(eng, seed=11):
    from msa_tts_b200.engine import batch_to_device
    cfg = eng.cfg
    flat = eng.flat_from_dict(synth.init_params(cfg, 0))
    bd = batch_to_device(synth.make_batch(cfg, B, T, L, seed), eng.device)
    masks = eng.generate_masks(B, T, L, 77)
    return flat, bd, masks


@pytest.mark.parametrize("tf32", [0, 1])
def test_full_size_pass_is_bitwise_repeatable(tf32):
    """Same inputs, same masks -> the same bits, run to run and engine to engine: the persistent kernels hand data over in a fixed
    order, the column reductions combine partials in a fixed order, nothing uses floating-point atomics."""
    runs = []
    for _ in range(2):
        eng = _engine(tf32)
        flat, bd, masks = _inputs(eng)
        for _ in range(2):
            out, loss = eng.forward(flat, eng.new_bn_stats(), bd, masks)
            g = eng.new_flat(0.0)
            eng.backward(flat, g)
            torch.cuda.synchronize()
            eng.check_abort()
            runs.append(([o.clone() for o in out], loss.clone(), g))
        del eng
    ref = runs[0]
    assert bool(torch.isfinite(ref[2]).all()) and float(ref[2].abs().sum()) > 0
    for out, loss, g in runs[1:]:
        for a, b in zip(out, ref[0]):
            assert torch.equal(a, b)
        assert torch.equal(loss, ref[1]) and torch.equal(g, ref[2])


def test_full_size_backward_is_linear_in_the_upstream_gradients():
    """backward(a*d1 + b*d2) == a*backward(d1) + b*backward(d2) (fp32 GEMMs; 1e-5 of the gradient norm): every kernel of the
    backward pass -- both recurrent chains included -- is a linear map of the upstream gradients for a fixed forward pass."""
    eng = _engine(0)
    flat, bd, masks = _inputs(eng, seed=12)
    eng.forward(flat, eng.new_bn_stats(), bd, masks, outputs=False)
    gen = torch.Generator(device="cuda").manual_seed(5)
    M = eng.cfg["n_mel_channels"]

    def rand_d():
        return [torch.randn(B, M, T, device="cuda", generator=gen) * 1e-2, torch.randn(B, M, T, device="cuda", generator=gen) * 1e-2,
                torch.randn(B, T, device="cuda", generator=gen) * 1e-2]
    d1, d2 = rand_d(), rand_d()
    a, b = 0.75, -1.5

    def bwd(d):
        g = eng.new_flat(0.0)
        eng.backward(flat, g, d_outputs=d)
        return g
    g1, g2 = bwd(d1), bwd(d2)
    g12 = bwd([a * x + b * y for x, y in zip(d1, d2)])
    torch.cuda.synchronize()
    eng.check_abort()
    want = a * g1.double() + b * g2.double()
    assert float(want.norm()) > 0
    assert float((g12.double() - want).norm()) < 1e-5 * float(want.norm()), rel(g12, want)
    # and the scale argument is the same linear map applied to the gradient
    gs = eng.new_flat(0.0)
    eng.backward(flat, gs, scale=0.125, d_outputs=d1)
    assert rel(gs, 0.125 * g1) < 1e-6


@pytest.mark.parametrize("grouped", [False, True])
def test_full_size_fomaml_meta_gradient_is_the_mean_of_the_task_gradients(grouped):
    """BASELINE configs[1] (8 tasks, 1 inner SGD step, bench GEMM policy): the meta-gradient that the test-split backward passes
    accumulate in their epilogues (meta_grad += g/8) equals the mean of the 8 task gradients computed one by one into separate
    buffers (utils/grad_utils.py:23-31).  Without task grouping the task losses agree bit for bit and the gradients to 2e-6;
    with the first inner steps of the 8 tasks grouped into one pass (the bench's default: bf16x3 tensor-core recurrences for
    the train split) the adapted weights differ by ~1e-5 of the inner update, the test losses by < 1e-6 and the meta-gradient
    stays inside the TF32-path tolerance of the north star (1e-3; measured ~1e-4: TF32 operand rounding of the backward GEMMs)."""
    import bench
    from msa_tts_b200.maml import MAML
    params = bench.trainer_params(1)
    params["group_tasks"] = grouped
    params["optim_outer"] = {"optimizer_name": "SGD", "optim_params": {"lr": "0.0"}}       # keep theta: the tasks are re-run below
    tr = MAML(**params)
    items = bench.make_tasks(tr.model_params, pinned=False)
    log = tr._metatrain_step(items)
    torch.cuda.synchronize()
    tr.engine.check_abort()
    meta = tr.meta_grad.clone()
    losses = log["loss_test"].clone()
    tr.step_global = 0                                                                       # same dropout-mask keys as above
    eng, n = tr.engine, len(items)
    acc = torch.zeros_like(meta, dtype=torch.float64)
    g = eng.new_flat(0.0)
    for i, spk in enumerate(items):
        tr._adapt(i, items[spk]["train"], 1)
        inputs, _ = tr._unpack_batch(items[spk]["test"])
        _, loss = eng.forward(tr.fast, tr.task_bn, inputs, tr._masks(i, 1, B, T, L), outputs=False)
        eng.backward(tr.fast, g)
        torch.cuda.synchronize()
        if grouped:
            assert abs(float(loss) - float(losses[i])) < 1e-5 * abs(float(loss))
        else:
            assert torch.equal(loss, losses[i:i + 1])
        acc += g.double()
    eng.check_abort()
    want = acc / n
    err = float((meta.double() - want).norm()) / float(want.norm())
    print(f"grouped={grouped}: meta-gradient vs mean of one-by-one task gradients {err:.2e}")
    assert err < (1e-3 if grouped else 2e-6)
    assert abs(float(log["grad_sumsq"]) ** 0.5 - float(want.norm())) < (1e-3 if grouped else 1e-5) * float(want.norm())


def _group_case(cfg, G, B, T, L, tf32):
    from msa_tts_b200.engine import Engine, batch_to_device
    eng = Engine(cfg, gemm_tf32=tf32)
    flat = eng.flat_from_dict(synth.init_params(cfg, 3))
    bds = [batch_to_device(synth.make_batch(cfg, B, T, L, 500 + g), eng.device) for g in range(G)]
    masks = [eng.generate_masks(B, T, L, 900 + g) for g in range(G)]
    return eng, flat, bds, masks


@pytest.mark.parametrize("shape", [("small", 3, 3, 11, 9, 0, 1), ("small", 8, 4, 9, 8, 0, 4), ("small", 5, 3, 9, 8, 0, 4), ("small", 1, 5, 9, 8, 0, 4),
                                   ("default", 2, 4, 12, 20, 0, 4), ("default", 8, 4, 10, 16, 0, 4), ("default", 3, 6, 10, 16, 0, 4),
                                   ("default", 8, 4, 10, 16, 1, 1)])
def test_grouped_pass_equals_the_passes_one_by_one(shape, monkeypatch):
    """msa_train_forward_group / msa_train_backward_group (the theta_0 train passes of a meta-batch as ONE pass, maml.py:38-54):
    per-task losses, BatchNorm running statistics and parameter gradients equal those of G separate passes.
    * strict fp32 policy, default kernel choice: the recurrences run task by task inside the group -- equal to rounding;
    * fp32 GEMMs with the grouped tensor-core kernels FORCED (MSA_CHAIN_MMA=4, a test mode): their bf16x3 gate products against
      the fp32 FMA recurrences of the single-task kernels -- 5e-5 of the gradient norm;
    * the bench policy (TF32 backward GEMMs, grouped kernels by default): 1e-3, the north star's TF32-path tolerance (TF32
      operand rounding turns the 1e-5 differences of the recurrences into ~1e-4 differences of the weight gradients)."""
    which, G, B, T, L, tf32, mode = shape
    monkeypatch.setenv("MSA_CHAIN_MMA", str(mode))
    cfg = pkg.small_params() if which == "small" else pkg.default_params()
    eng, flat, bds, masks = _group_case(cfg, G, B, T, L, tf32)
    bn_g = [eng.new_bn_stats() for _ in range(G)]
    grads_g = [eng.new_flat() for _ in range(G)]
    loss_g = eng.forward_group(flat, bn_g, bds, masks)
    eng.backward_group(flat, grads_g)
    torch.cuda.synchronize()
    eng.check_abort()
    monkeypatch.setenv("MSA_CHAIN_MMA", "0")
    from msa_tts_b200.engine import Engine
    ref = Engine(cfg, gemm_tf32=tf32)             # single-task fp32-FMA recurrences
    for g in range(G):
        bn, gr = ref.new_bn_stats(), ref.new_flat()
        _, loss = ref.forward(flat, bn, bds[g], masks[g], outputs=False)
        ref.backward(flat, gr)
        torch.cuda.synchronize()
        gn = float(gr.double().norm())
        e_loss = abs(float(loss_g[g]) - float(loss)) / abs(float(loss))
        e_bn = float((bn_g[g] - bn).double().norm() / bn.double().norm())
        e_g = float((grads_g[g] - gr).double().norm()) / gn
        print(f"task {g}: loss {e_loss:.2e} bn {e_bn:.2e} grad {e_g:.2e}")
        tol = 1e-3 if tf32 else (5e-5 if mode == 4 else 1e-6)
        assert e_loss < tol and e_bn < tol and e_g < tol, (g, e_loss, e_bn, e_g)
    ref.check_abort()


@pytest.mark.parametrize("shape", [("small", 4, 3, 9, 8, 0, 4, 4), ("small", 8, 4, 9, 8, 0, 4, 4), ("small", 5, 3, 11, 9, 0, 4, 4),
                                   ("small", 3, 8, 9, 8, 0, 4, 3), ("small", 2, 1, 9, 8, 0, 4, 4),
                                   ("default", 4, 4, 10, 16, 0, 4, 4), ("default", 7, 4, 10, 16, 0, 4, 4), ("default", 6, 2, 12, 20, 0, 4, 3),
                                   ("default", 8, 4, 10, 16, 1, 1, 4), ("default", 8, 4, 10, 16, 1, 1, 2)])
def test_per_task_weight_groups_equal_the_passes_one_by_one(shape, monkeypatch):
    """Groups whose tasks have their OWN weights (the test-split passes / later inner steps of a meta-batch, maml.py:56-76): the
    attention chains of up to MSA_PT_GROUP tasks share a launch, every task's recurrent weights streamed as fragments
    (chain_mma.cu, PT variants).  Per-task losses, BatchNorm statistics and gradients equal those of G separate passes with the
    single-task fp32-FMA kernels: 5e-5 of the gradient norm with fp32 GEMMs (grouped kernels forced, MSA_CHAIN_MMA=4), 1e-3 under
    the bench policy (TF32 backward GEMMs)."""
    which, G, B, T, L, tf32, mode, ptg = shape
    monkeypatch.setenv("MSA_CHAIN_MMA", str(mode))
    monkeypatch.setenv("MSA_PT_GROUP", str(ptg))
    monkeypatch.setenv("MSA_PT_BWD", "1")         # the backward chains too (off by default: not faster than single-task launches)
    cfg = pkg.small_params() if which == "small" else pkg.default_params()
    eng, flat, bds, masks = _group_case(cfg, G, B, T, L, tf32)
    gen = torch.Generator(device="cpu").manual_seed(77)
    plist = []
    for g in range(G):          # every task its own weights: theta + a task-specific perturbation of every parameter
        noise = torch.randn(flat.numel(), generator=gen).to(flat.device)
        plist.append((flat * (1.0 + 0.05 * noise) + 0.002 * noise).contiguous())
    bn_g = [eng.new_bn_stats() for _ in range(G)]
    grads_g = [eng.new_flat() for _ in range(G)]
    eng.profile(True)
    loss_g = eng.forward_group(plist, bn_g, bds, masks)
    eng.backward_group(plist, grads_g)
    torch.cuda.synchronize()
    eng.check_abort()
    prof = eng.profile_read()
    eng.profile(False)
    # chunks of two or more tasks are ONE launch each of the per-task-weight kernels; the chunk size is MSA_PT_GROUP or the largest
    # smaller size whose rings fit into shared memory (default dims: four tasks do not)
    def launches(cap):
        n_chunks = -(-G // cap)
        sizes = [G // n_chunks + (1 if i < G % n_chunks else 0) for i in range(n_chunks)]
        return sum(1 for s_ in sizes if s_ >= 2)
    allowed = {launches(cap) for cap in range(2, ptg + 1)} if which == "default" else {launches(ptg)}
    assert prof["attn_chain_fwd_pt"][1] in allowed and prof["attn_chain_fwd_pt"][1] >= 1, (prof, allowed)
    assert prof["attn_chain_bwd_pt"][1] == prof["attn_chain_fwd_pt"][1], prof
    monkeypatch.setenv("MSA_CHAIN_MMA", "0")
    from msa_tts_b200.engine import Engine
    ref = Engine(cfg, gemm_tf32=tf32)             # single-task fp32-FMA recurrences
    for g in range(G):
        bn, gr = ref.new_bn_stats(), ref.new_flat()
        _, loss = ref.forward(plist[g], bn, bds[g], masks[g], outputs=False)
        ref.backward(plist[g], gr)
        torch.cuda.synchronize()
        gn = float(gr.double().norm())
        e_loss = abs(float(loss_g[g]) - float(loss)) / abs(float(loss))
        e_bn = float((bn_g[g] - bn).double().norm() / bn.double().norm())
        e_g = float((grads_g[g] - gr).double().norm()) / gn
        print(f"task {g}: loss {e_loss:.2e} bn {e_bn:.2e} grad {e_g:.2e}")
        tol = 1e-3 if tf32 else 5e-5
        assert e_loss < tol and e_bn < tol and e_g < tol, (g, e_loss, e_bn, e_g)
    ref.check_abort()
