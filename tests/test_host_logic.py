"""Host-side logic that needs no GPU: flat layout, optimizer spec parsing, task sharding, and the sharded meta-gradient
reduction over a 2-process gloo group (the N>1 path of parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import msa_tts_b200 as pkg
from msa_tts_b200.helpers import optimizer_hparams
from msa_tts_b200.layout import FlatLayout
from msa_tts_b200.parallel import ShardInfo
from oracle import model as OM


def test_flat_layout_matches_reference_parameter_order():
    for cfg in (pkg.small_params(), pkg.default_params()):
        lay = FlatLayout(cfg)
        assert lay.names() == OM.param_names(cfg)
        offs = [lay.offsets[n] for n in lay.names()]
        assert offs == sorted(offs) and all(o % 32 == 0 for o in offs), "tensors start on 128-byte boundaries, in order"
        assert lay.total >= lay.n_params
    assert FlatLayout(pkg.default_params()).n_params == 30331010


def test_optimizer_spec_strings_are_evaluated_like_helpers_get_optimizer():
    h = optimizer_hparams({"optimizer_name": "Adam", "optim_params": {"lr": "1e-3", "betas": "(0.9, 0.98)", "weight_decay": "0"}})
    assert h == {"name": "Adam", "lr": 1e-3, "betas": (0.9, 0.98), "weight_decay": 0}
    with pytest.raises(ValueError):
        optimizer_hparams({"optimizer_name": "SGD", "optim_params": {}})


def test_task_sharding_is_a_partition():
    for world in (1, 2, 3, 8):
        for n in (1, 8, 16, 5):
            seen = sorted(i for r in range(world) for i in ShardInfo(r, world).my_tasks(n))
            assert seen == list(range(n))
            assert max(len(ShardInfo(r, world).my_tasks(n)) for r in range(world)) - \
                   min(len(ShardInfo(r, world).my_tasks(n)) for r in range(world)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_tasks, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = ShardInfo.from_env()
        assert (shard.rank, shard.world) == (rank, world)
        acc = torch.zeros(n)
        losses = []
        for i in shard.my_tasks(n_tasks):                       # each rank adapts its own subset (maml.py:38-41)
            g = torch.Generator().manual_seed(100 + i)
            acc += torch.randn(n, generator=g) / n_tasks        # sum_local w_i g_i, w_i = 1/N (maml.py:94-98)
            losses.append(float(i))
        shard.allreduce_sum(acc)                                # the ONE collective of a meta-step
        full = shard.gather_scalars(torch.tensor(losses), n_tasks)
        mx = torch.tensor([float(rank)])
        shard.allreduce_max(mx)
        shard.barrier()
        if rank == 0:
            out.put((acc.numpy(), full.numpy(), float(mx)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_tasks", [8, 5])
def test_sharded_meta_gradient_equals_unsharded_gloo_world2(n_tasks):
    n, world = 4096, 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tasks, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    acc, losses, mx = out.get()
    ref = torch.zeros(n)
    for i in range(n_tasks):
        g = torch.Generator().manual_seed(100 + i)
        ref += torch.randn(n, generator=g) / n_tasks
    assert np.allclose(acc, ref.numpy(), rtol=0, atol=1e-6), "sharded sum = unsharded mix_grad up to fp32 summation order"
    assert losses.tolist() == [float(i) for i in range(n_tasks)]
    assert mx == 1.0


@pytest.mark.parametrize("r", [1, 2, 3])
def test_collators_match_the_reference_collators_bit_for_bit(r):
    """msa_tts_b200.data.Collator / MetaCollator against tests/golden/collate.npz, which holds the outputs of the reference's own
    collators on the same raw items (oracle/gen_golden_collate.py): order of the sorted items, padding values, padding to a
    multiple of the reduction factor, stop targets, dtypes -- all exact."""
    from msa_tts_b200.data import Collator, MetaCollator
    from oracle.gen_cases import collate_items
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "collate.npz"))
    keys = ("transcripts", "trans_lengths", "melspecs", "melspec_lengths", "speaker_ids", "spk_embs", "stop_targets")

    def check(batch, prefix):
        assert list(batch[0]) == list(z[prefix + "/item_ids"])
        for k, t in zip(keys, batch[1:]):
            ref = z[f"{prefix}/{k}"]
            assert tuple(t.shape) == ref.shape and np.array_equal(t.numpy(), ref), (prefix, k)
            if f"{prefix}/{k}/dtype" in z:
                assert str(t.dtype) == str(z[f"{prefix}/{k}/dtype"])
        assert batch[3].shape[2] % r == 0 and batch[7].shape[1] == batch[3].shape[2]

    check(Collator(r)(collate_items(10 + r, 5)), f"default_r{r}")
    meta = [("spkA", {"train": collate_items(20 + r, 4), "test": collate_items(30 + r, 3)}),
            ("spkB", {"train": collate_items(40 + r, 2), "test": collate_items(50 + r, 4)})]
    d = MetaCollator(r)(meta)
    assert list(d.keys()) == ["spkA", "spkB"]
    for spk in d:
        for mode in ("train", "test"):
            check(d[spk][mode], f"meta_r{r}/{spk}/{mode}")


def test_task_group_plan():
    """Host logic of the grouped passes: tasks of a rank are grouped by the shape of their train batch, at most 8 tasks and 32 rows
    per group; stateful inner optimizers with several inner steps, n_inner = 0 and group_tasks=False take the plain path."""
    from msa_tts_b200.metatrainer import plan_task_groups
    gs = lambda n, B: max(1, min(n, 8, 32 // max(B, 1)))
    same = {i: ((4, 64), 200) for i in range(8)}
    assert plan_task_groups(list(range(8)), same, 1, False, True, gs) == [list(range(8))]
    assert plan_task_groups([0, 2, 4, 6], same, 1, False, True, gs) == [[0, 2, 4, 6]]          # rank 0 of 2
    assert plan_task_groups(list(range(8)), same, 1, False, False, gs) == [[i] for i in range(8)]
    assert plan_task_groups(list(range(8)), same, 0, False, True, gs) == [[i] for i in range(8)]
    assert plan_task_groups(list(range(8)), same, 5, True, True, gs) == [[i] for i in range(8)]   # momentum / Adam, 5 inner steps
    assert plan_task_groups(list(range(8)), same, 1, True, True, gs) == [list(range(8))]          # ... one step: state starts at zero
    b16 = {i: ((16, 64), 200) for i in range(5)}
    assert plan_task_groups(list(range(5)), b16, 2, False, True, gs) == [[0, 1], [2, 3], [4]]      # 32 rows per launch
    mixed = {0: ((4, 64), 200), 1: ((4, 60), 200), 2: ((4, 64), 200), 3: ((4, 64), 180), 4: ((4, 60), 200)}
    assert plan_task_groups(list(range(5)), mixed, 1, False, True, gs) == [[0, 2], [1, 4], [3]]    # ragged speakers: by shape
    many = {i: ((4, 64), 200) for i in range(16)}
    assert plan_task_groups(list(range(16)), many, 5, False, True, gs) == [list(range(8)), list(range(8, 16))]


def test_reference_call_surface_names_and_keyword_defaults():
    """The drop-in surface keeps the names and keyword defaults of what it replaces: ``higher.innerloop_ctx(model, opt,
    copy_initial_weights=True, track_higher_grads=True)`` (maml.py:40-41 passes the flag explicitly; an omitted flag must not turn
    into a silent first-order run), ``mix_grad(grad_list, weight_list)`` / ``apply_grad(model, grad)`` (utils/grad_utils.py:8,23),
    ``run()`` / ``_metatrain(epoch)`` / ``_metatest(epoch)`` on both trainers (maml.py:19,33,115; reptile.py:19,33,108)."""
    import inspect
    import msa_tts_b200 as pkg
    from msa_tts_b200 import grad_utils
    from msa_tts_b200.maml import MAML
    from msa_tts_b200.reptile import Reptile
    sig = inspect.signature(pkg.innerloop_ctx)
    assert list(sig.parameters)[:2] == ["model", "opt"]
    assert sig.parameters["track_higher_grads"].default is True and sig.parameters["copy_initial_weights"].default is True
    mg = inspect.signature(grad_utils.mix_grad).parameters
    assert list(mg)[:2] == ["grad_list", "weight_list"] and all(p.default is not inspect.Parameter.empty for p in list(mg.values())[2:])
    assert list(inspect.signature(grad_utils.apply_grad).parameters) == ["model", "grad"]
    for cls in (MAML, Reptile):
        for name in ("run", "_metatrain", "_metatest", "_metatrain_step", "_metatest_step", "_unpack_batch", "_save_checkpoint",
                     "_load_checkpoint"):
            assert callable(getattr(cls, name)), (cls.__name__, name)
        for name in ("_metatrain", "_metatest"):       # callable as the reference calls them: (epoch) only
            ps = list(inspect.signature(getattr(cls, name)).parameters.values())
            assert ps[1].name == "epoch" and all(p.default is not inspect.Parameter.empty for p in ps[2:])
        assert all(p.default is not inspect.Parameter.empty for p in list(inspect.signature(cls.run).parameters.values())[1:])


class _RecordingEngine:
    """Test double of engine.Engine for the ORCHESTRATION of the trainers only (which call follows which, what is never called);
    it computes nothing.  The product path has no such thing: MetaTrainer.__init__ builds the real engine or fails."""

    def __init__(self):
        self.calls = []
        self.device = torch.device("cpu")

    def _rec(self, name, *a):
        self.calls.append((name,) + a)

    def forward(self, params, bn, bd, masks, outputs=True):
        self._rec("forward", params.data_ptr(), int(masks), bool(outputs))
        out = [torch.zeros(1)] * 4 if outputs else None
        return out, torch.full((1,), float(len(self.calls)))

    def backward(self, params, grads, accumulate=False, scale=1.0, **kw):
        self._rec("backward", params.data_ptr(), bool(accumulate), round(float(scale), 6))

    def sgd_step(self, p, g, p_out=None, **kw):
        self._rec("sgd_step", p.data_ptr(), (p if p_out is None else p_out).data_ptr())

    def mcd(self, lens, which=0):
        self._rec("mcd")
        return torch.zeros(1)

    def abort_poll(self):
        self._rec("abort_poll")

    def abort_flush(self):
        self._rec("abort_flush")

    def group_size(self, n, B):
        return 1

    def encoder_prefix(self):
        return 0

    def new_flat(self, fill=0.0):
        return torch.zeros(8)

    def new_bn_stats(self):
        return torch.zeros(4)

    def forward_group(self, params, bn_stats, batches_dev, masks):
        ps = [p.data_ptr() for p in params] if isinstance(params, (list, tuple)) else params.data_ptr()
        self._rec("forward_group", ps, [int(m) for m in masks])
        return torch.zeros(len(batches_dev))

    def backward_group(self, params, grads, accumulate=False, scale=1.0):
        ps = [p.data_ptr() for p in params] if isinstance(params, (list, tuple)) else params.data_ptr()
        self._rec("backward_group", ps, [g.data_ptr() for g in grads])

    def mcd_group(self, g, lens, which=0):
        self._rec("mcd_group", int(g))
        return torch.zeros(1)

    def axpy(self, acc, g, w, init):
        self._rec("axpy", g.data_ptr(), round(float(w), 6), bool(init))

    def reptile_delta(self, acc, p_T, p_0, w, init):
        self._rec("reptile_delta", p_T.data_ptr(), p_0.data_ptr(), round(float(w), 6), bool(init))

    def sumsq(self, g, out=None):
        self._rec("sumsq")
        return out

    def abort_guard(self, sumsq):
        self._rec("abort_guard")

    def clip_sgd(self, p, g, sumsq, lr, **kw):
        self._rec("clip_sgd", p.data_ptr(), bool(kw.get("first_step")))

    def clip_adam(self, p, g, m, v, sumsq, lr, step, **kw):
        self._rec("clip_adam", p.data_ptr(), int(step))


def _bare_trainer(cls, **params):
    """A trainer object with the attributes its epoch-level methods use, without the CUDA engine of __init__."""
    from msa_tts_b200.parallel import ShardInfo
    tr = object.__new__(cls)
    tr.params = dict(n_inner_train=1, group_tasks=False, **params)
    tr.engine, tr.shard, tr.device = _RecordingEngine(), ShardInfo(), torch.device("cpu")
    tr.speaker_emb_type = "static"
    tr.inner = {"name": "SGD", "lr": 0.1}
    tr.theta, tr.fast, tr.task_grad = torch.zeros(8), torch.zeros(8), torch.zeros(8)
    tr.base_bn, tr.task_bn = torch.ones(4), torch.zeros(4)
    tr.inner_buf = None
    tr.step_global, tr._outer_steps = 0, 0
    tr.outer = {"name": "Adam", "lr": 1e-3}
    tr.meta_grad, tr.sumsq, tr.outer_m, tr.outer_v = torch.zeros(8), torch.zeros(1), torch.zeros(8), torch.zeros(8)
    tr.injected_masks = None
    tr._masks = lambda i, p, B, T, L, slot=0: 1000 * i + p          # the mask KEY (task, pass) instead of a mask buffer
    return tr


def _toy_task(B=2, T=5, L=3):
    z = torch.zeros
    batch = (["a"] * B, z(B, L, dtype=torch.long), z(B, dtype=torch.long), z(B, 4, T), z(B, dtype=torch.long), z(B, dtype=torch.long),
             z(B, 2), z(B, T))
    return {"train": batch, "test": batch}


def test_metatest_orchestration_adapts_then_evaluates_without_a_gradient_or_an_outer_step():
    """maml.py:115-179 on a recording engine: per speaker n_inner_test x (forward, backward, inner step) on the train split -- the
    first step reading theta and writing the fast weights -- then ONE forward + MCD on the test split with the adapted weights, no
    backward after it, no outer update, theta / base BatchNorm statistics / step counters untouched; the dropout-mask stream of the
    meta-test passes starts at METATEST_PASS0."""
    from msa_tts_b200.maml import MAML
    tr = _bare_trainer(MAML, n_inner_test=2)
    items = {"s0": _toy_task(), "s1": _toy_task()}
    log = tr._metatest_step(items)
    th, fa, p0 = tr.theta.data_ptr(), tr.fast.data_ptr(), MAML.METATEST_PASS0
    per_task = lambda i: [("forward", th, 1000 * i + p0, False), ("backward", th, False, 1.0), ("sgd_step", th, fa),
                          ("forward", fa, 1000 * i + p0 + 1, False), ("backward", fa, False, 1.0), ("sgd_step", fa, fa),
                          ("forward", fa, 1000 * i + p0 + 2, False), ("mcd",)]
    assert tr.engine.calls == per_task(0) + per_task(1) + [("abort_poll",)]
    assert log["task_index"] == [0, 1] and log["speakers"] == ["s0", "s1"] and log["loss_test"].numel() == 2 and log["mcd"].numel() == 2
    assert tr.step_global == 0 and tr._outer_steps == 0 and float(tr.theta.abs().sum()) == 0.0
    assert torch.equal(tr.base_bn, torch.ones(4)) and torch.equal(tr.task_bn, tr.base_bn)      # the private copy was re-seeded
    # with the outputs asked for, the test pass returns them and still takes no gradient
    tr.engine.calls.clear()
    log = tr._metatest_step({"s0": _toy_task()}, return_outputs=True)
    assert [c[0] for c in tr.engine.calls].count("backward") == 2 and tr.engine.calls[-3][:1] == ("forward",) and tr.engine.calls[-3][3]
    assert len(log["outputs"]) == 1 and len(log["outputs"][0]) == 4


def test_run_schedules_checkpoints_and_metatests_like_the_reference_epoch_loop():
    """maml.py:19-31: ``run`` = per epoch ``_metatrain``; ``_save_checkpoint`` every ckpt_save_epoch_interval, ``_metatest`` every
    metatest_epoch_interval epochs; step_global restarts at 0."""
    from msa_tts_b200.reptile import Reptile
    tr = _bare_trainer(Reptile, n_epochs=4, ckpt_save_epoch_interval=2, metatest_epoch_interval=3)
    seen = []
    tr._metatrain_step = lambda items_b: seen.append(("train", items_b)) or {"i": items_b}
    tr._metatest_step = lambda items_b: seen.append(("test", items_b)) or {"t": items_b}
    tr._save_checkpoint = lambda path=None: seen.append(("ckpt",))
    tr.step_global = 17
    logs = tr.run(dataloader_metatrain=["b0", "b1"], dataloader_metatest=["m0"])
    want = []
    for epoch in range(1, 5):
        want += [("train", "b0"), ("train", "b1")]
        if epoch % 2 == 0:
            want.append(("ckpt",))
        if epoch % 3 == 0:
            want.append(("test", "m0"))
    assert seen == want and len(logs) == 8 and tr.last_metatest == [{"t": "m0"}]
    assert tr.step_global == 0          # the lambdas above do not step; run() reset the counter (maml.py:20)
    with __import__("pytest").raises(RuntimeError):
        _bare_trainer(Reptile)._metatest(1)                     # no loader: a loud error, not an empty epoch


def test_reptile_outer_loop_semantics_sequential_by_default_batched_on_request():
    """reptile.py:37-89 on a recording engine.  Default on one GPU = the reference's literal loop: an outer step after EACH speaker
    (the next speaker adapts from the updated weights), step_global and the optimizer's step count advancing once per speaker.
    ``reptile_sequential=False``: every speaker adapts from the same theta, the deltas are mixed with weights 1/N, ONE outer step."""
    from msa_tts_b200.reptile import Reptile
    items = {"s0": _toy_task(), "s1": _toy_task()}

    def names(tr):
        return [c[0] for c in tr.engine.calls]
    tr = _bare_trainer(Reptile)
    tr.params["n_inner_train"] = 2
    tr.sequential = True                                  # what Reptile.__init__ picks when world == 1 and the key is absent
    tr._metatrain_step(items)
    th, fa = tr.theta.data_ptr(), tr.fast.data_ptr()
    spk = ["forward", "backward", "sgd_step", "forward", "backward", "sgd_step", "forward", "mcd", "reptile_delta",
           "sumsq", "abort_guard", "clip_adam", "abort_poll"]
    assert names(tr) == spk + spk
    deltas = [c for c in tr.engine.calls if c[0] == "reptile_delta"]
    assert deltas == [("reptile_delta", fa, th, 1.0, True)] * 2                                  # reptile.py:75-77, weight 1
    assert [c[2] for c in tr.engine.calls if c[0] == "clip_adam"] == [1, 2] and tr.step_global == 2 and tr._outer_steps == 2
    tr = _bare_trainer(Reptile)
    tr.params["n_inner_train"] = 2
    tr.sequential = False
    tr._metatrain_step(items)
    task = spk[:9]
    assert names(tr) == task + task + spk[9:]
    assert [c[3:] for c in tr.engine.calls if c[0] == "reptile_delta"] == [(0.5, True), (0.5, False)]
    assert tr.step_global == 1 and tr._outer_steps == 1
    # Reptile.__init__'s choice of the default, without building an engine: sequential unless sharded
    import inspect
    src = inspect.getsource(Reptile.__init__)
    assert "seq = self.shard.world == 1" in src and "cannot be sharded" in src


def test_fomaml_meta_step_orchestration_accumulates_the_weighted_test_gradients():
    """maml.py:36-105 on a recording engine (plain path, copies not staged): per speaker the inner step(s) from theta, the test-split
    forward with the adapted weights, its MCD, and a backward that writes ``meta_grad (+)= g / N`` -- first task overwrites, the
    others accumulate (mix_grad with weights 1/N, maml.py:94-98) -- then ONE outer update."""
    from msa_tts_b200.maml import MAML
    tr = _bare_trainer(MAML)
    tr._stage = lambda batches: batches            # the side-stream H2D staging needs CUDA; _unpack_batch takes plain tuples too
    tr._bwd_event = None
    items = {f"s{i}": _toy_task() for i in range(3)}
    log = tr._metatrain_step(items)
    th, fa = tr.theta.data_ptr(), tr.fast.data_ptr()
    want = []
    for i in range(3):
        want += [("forward", th, 1000 * i, False), ("backward", th, False, 1.0), ("sgd_step", th, fa),
                 ("forward", fa, 1000 * i + 1, False), ("mcd",), ("backward", fa, i > 0, round(1.0 / 3, 6))]
    want += [("sumsq",), ("abort_guard",), ("clip_adam", th, 1), ("abort_poll",)]
    assert tr.engine.calls == want
    assert log["task_index"] == [0, 1, 2] and log["loss_test"].numel() == 3 and log["mcd"].numel() == 3
    assert tr.step_global == 1 and tr._outer_steps == 1


def test_grouped_fomaml_meta_step_orchestration():
    """The task-group path (metatrainer.plan_task_groups, include/msa_b200.h "task groups") on a recording engine: ONE grouped pass
    from theta for the first inner step of all tasks (maml.py:38-41: every task starts from the same weights), per-task functional
    updates into the slots' fast weights, ONE grouped test pass with per-task weights, and the slots' gradients mixed into the
    meta-gradient with weights 1/N (first overwrites)."""
    from msa_tts_b200.maml import MAML
    tr = _bare_trainer(MAML)
    tr.params["group_tasks"] = True
    tr.engine.group_size = lambda n, B: n
    tr._stage = lambda batches: batches
    tr._bwd_event, tr._slots = None, []
    tr._masks = lambda i, p, B, T, L, slot=0: 1000 * i + 10 * slot + p
    items = {f"s{i}": _toy_task() for i in range(3)}
    log = tr._metatrain_step(items)
    th = tr.theta.data_ptr()
    fast = [tr._slot(k)[0].data_ptr() for k in range(3)]
    grad = [tr._slot(k)[1].data_ptr() for k in range(3)]
    assert fast[0] == tr.fast.data_ptr() and len(set(fast)) == 3 and len(set(grad)) == 3       # slot 0 = the plain path's buffers
    want = [("forward_group", th, [0, 1010, 2020]), ("backward_group", th, grad)]
    want += [("sgd_step", th, fast[k]) for k in range(3)]
    want += [("forward_group", fast, [1, 1011, 2021]), ("mcd_group", 0), ("mcd_group", 1), ("mcd_group", 2),
             ("backward_group", fast, grad)]
    want += [("axpy", grad[k], round(1.0 / 3, 6), k == 0) for k in range(3)]
    want += [("sumsq",), ("abort_guard",), ("clip_adam", th, 1), ("abort_poll",)]
    assert tr.engine.calls == want
    assert log["task_index"] == [0, 1, 2] and log["loss_test"].numel() == 3 and log["mcd"].numel() == 3
