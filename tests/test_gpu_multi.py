"""Two GPUs, one process per GPU over NCCL: a sharded meta-step (task i -> rank i % 2, ONE allreduce of the flat meta-gradient,
replicated clip + outer update) equals the unsharded one.  Skipped on a single-GPU box; run with
    gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_multi.py -m gpu -q'
Tolerance: 2e-6 of the norm (only the fp32 summation order of the task gradients differs); the two ranks must end with
bit-identical weights (no broadcast exists, the update is replicated).  The outer optimizer is SGD here: Adam's normalised step
turns a sign change of a nearly cancelling gradient sum into a full +-lr difference, which says nothing about the sharding."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
B, T, L = 3, 12, 9


def _params(cfg, n_inner, outer):
    return {"model": cfg, "criterion": {"criterion_type": "Tacotron2Loss", "reduction": "none", "pos_weight": 10.0},
            "optim_inner": {"optimizer_name": "SGD", "optim_params": {"lr": "0.05"}},
            "optim_outer": {"optimizer_name": outer, "optim_params": {"lr": "0.01"}},
            "n_inner_train": n_inner, "track_higher_grads": False, "clip_grad_norm": True, "grad_clip_thresh": 0.5, "init_seed": 5,
            # the batched Reptile variant is the one that shards (on one GPU the reference's per-speaker sequential loop is the default)
            "reptile_sequential": False}


def _run(kind, n_tasks, n_inner, steps=2):
    """steps meta-steps of the trainer `kind` on the current device / process group; device-generated dropout masks (keyed by
    meta-step, task and pass -- not by rank)."""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.maml import MAML
    from msa_tts_b200.reptile import Reptile
    cfg = pkg.small_params()
    tasks = {f"spk{i}": synth.make_task(cfg, B, T, L, 60 + i) for i in range(n_tasks)}
    tr = (MAML if kind == "maml" else Reptile)(**_params(cfg, n_inner, "SGD"))
    logs = []
    for _ in range(steps):
        log = tr._metatrain_step(tasks)
        logs.append(tr.shard.gather_scalars(log["loss_test"], n_tasks).cpu())
    torch.cuda.synchronize()
    tr.engine.check_abort()
    return tr.theta.cpu(), tr.meta_grad.cpu(), torch.stack(logs)


def _worker(rank, world, port, kind, n_tasks, n_inner, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        theta, meta, losses = _run(kind, n_tasks, n_inner)
        out.put((rank, theta.numpy(), meta.numpy(), losses.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
# ("maml", 2, 1): one task per rank -- the plain path, whose last backward pass starts the allreduce of everything but the encoder
# gradients early (msa_backward_mark_event, parallel.py); ("maml", 3, 1): two tasks on rank 0 (grouped), one on rank 1 (plain, early
# start): both ranks must still issue the same two collectives in the same order
@pytest.mark.parametrize("kind,n_tasks,n_inner", [("maml", 5, 1), ("maml", 2, 1), ("maml", 3, 1), ("reptile", 4, 2)])
def test_sharded_meta_steps_equal_unsharded_nccl_world2(kind, n_tasks, n_inner):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n_tasks, n_inner, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, theta, meta, losses = out.get()
        got[rank] = (torch.from_numpy(theta), torch.from_numpy(meta), torch.from_numpy(losses))
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    theta1, meta1, losses1 = _run(kind, n_tasks, n_inner)                      # unsharded, this process, cuda:0
    assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1]), "replicated update must be bit-identical"
    assert torch.equal(got[0][2][0], losses1[0]), "per-task losses of the first meta-step do not depend on the sharding"
    assert torch.allclose(got[0][2][1], losses1[1], rtol=1e-5, atol=0)         # second step: theta differs by summation order
    gn = float(meta1.double().norm())
    assert float((got[0][1].double() - meta1.double()).norm()) < 2e-6 * gn
    upd = float((theta1.double() - _theta0().double()).norm())
    # the weights themselves round at 6e-8 relative per element: with updates this small (lr 0.01, clip 0.5) that rounding is
    # ~1e-5 of the update, so the bound is 1e-4 of the update and 1e-6 of the weights
    dth = float((got[0][0].double() - theta1.double()).norm())
    assert upd > 0 and dth < 1e-4 * upd and dth < 1e-6 * float(theta1.double().norm())


def _theta0():
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.layout import FlatLayout
    cfg = pkg.small_params()
    lay = FlatLayout(cfg)
    P = synth.init_params(cfg, 5)
    flat = torch.zeros(lay.total)
    for n in lay.names():
        flat[lay.offsets[n]:lay.offsets[n] + lay.numel(n)] = P[n].reshape(-1)
    return flat
