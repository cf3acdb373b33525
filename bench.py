#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1]): one first-order MAML meta-step over 8 synthetic speaker tasks, 1 inner SGD
step each, every split a batch of B=4 utterances x T=200 mel frames x L=64 tokens at the Tacotron-2 default
dimensions (30.33 M parameters): 16 teacher-forced forward+backward passes, 8 fused inner SGD steps, the fused
meta-gradient accumulation, clip + outer Adam step.  With N GPUs the 8 tasks are sharded (task i -> rank i % N) and
joined by one NCCL allreduce of the flat meta-gradient: total work is fixed => "scaling": "strong".

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference algorithm on the host CPU

`value`  : meta-steps/s, inputs already resident in HBM when the timed region starts.
`e2e`    : the same metric through the public API with HOST (pinned) batches: H2D of every batch and a D2H read of
           the per-task losses inside the timed region.
`roofline`, `cpu_baseline`, `clocks`, `gpu_launches`: see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TASKS, B, T, L = 8, 4, 200, 64
INNER_LR, N_INNER = 1e-3, 1
WORKLOAD = f"fomaml_meta_step_{N_TASKS}tasks_{N_INNER}inner_B{B}_T{T}_L{L}_default_dims"


def trainer_params(gemm_tf32: int):
    import msa_tts_b200 as pkg
    return {
        "model": pkg.default_params(),
        "criterion": {"criterion_type": "Tacotron2Loss", "reduction": "none", "pos_weight": 10.0},
        "optim_inner": {"optimizer_name": "SGD", "optim_params": {"lr": str(INNER_LR)}},
        "optim_outer": {"optimizer_name": "Adam", "optim_params": {"lr": "1e-4"}},
        "n_inner_train": N_INNER, "track_higher_grads": False, "clip_grad_norm": True, "grad_clip_thresh": 1.0,
        "meta_batch_size": N_TASKS, "dataset_random_seed": 1234, "gemm_tf32": gemm_tf32,
    }


def make_tasks(cfg, pinned=True):
    from msa_tts_b200 import synth
    items = {}
    for i in range(N_TASKS):
        task = synth.make_task(cfg, B, T, L, 1234 + i)
        if pinned:
            task = {k: tuple(x.pin_memory() if hasattr(x, "pin_memory") else x for x in v) for k, v in task.items()}
        items[f"spk{i}"] = task
    return items


def batch_bytes(batch):
    return sum(x.numel() * x.element_size() for x in batch[1:] if hasattr(x, "numel"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algo_bytes(cfg, kernel: str, rows: int = B) -> float:
    """Compulsory HBM bytes of one launch (every operand read once, every result written once).
    ``rows``: batch rows the launch carries (B for a single pass, G*B for a grouped launch, kernel names ending in "_grp")."""
    from msa_tts_b200.config import memory_dim, rnn_dims
    pt = kernel.endswith("_pt")
    kernel = kernel[:-4] if kernel.endswith("_grp") else (kernel[:-3] if pt else kernel)
    Ha, Hd = rnn_dims(cfg)
    E, A = memory_dim(cfg), cfg["attention_params"]["attention_dim"]
    F_, Hh = cfg["attention_params"]["attention_location_n_filters"], cfg["encoder_embedding_dim"] // 2
    TB, BL, TBL = T * rows, rows * L, T * rows * L
    f = 4.0
    if pt:      # per-task weights: every task's recurrent weight matrix (+ W_q) once
        G = max(1, rows // B)
        extra = f * (G - 1) * (4 * Ha * Ha + A * Ha)
        return algo_bytes(cfg, kernel, rows) + extra
    if kernel == "attn_chain_fwd":
        rd = TB * 4 * Ha + 4 * Ha * Ha + 4 * Ha * BL + A * Ha + BL * A
        wr = TB * (2 * Ha + 4 * Ha + A + 1) + 2 * TBL + TBL * A + TBL * F_
        return f * (rd + wr) + TB * Ha
    if kernel == "attn_chain_bwd":
        rd = 4 * Ha * Ha + BL * 4 * Ha + TB * (4 * Ha + Ha + Ha + 1) + 2 * TBL + TBL * A
        wr = TB * (4 * Ha + A) + TBL + TBL * A + TBL * F_
        return f * (rd + wr) + TB * Ha
    if kernel == "dec_lstm_fwd":
        return f * (TB * 4 * Hd + 4 * Hd * Hd + TB * (2 * Hd + 4 * Hd)) + TB * Hd
    if kernel == "dec_lstm_bwd":
        return f * (4 * Hd * Hd + TB * (4 * Hd + Hd + Hd) + TB * 4 * Hd) + TB * Hd
    if kernel == "enc_lstm_fwd":
        return f * 2 * (BL * 4 * Hh + 4 * Hh * Hh + BL * (2 * Hh + 4 * Hh))
    if kernel == "enc_lstm_bwd":
        return f * 2 * (4 * Hh * Hh + BL * (4 * Hh + 2 * Hh) + BL * 4 * Hh)
    raise KeyError(kernel)


# Cross-CTA hand-offs (all-gathers through L2 among the 148 co-resident CTAs) on the serial path of ONE recurrence step, and the
# bare cost of one such hand-off measured on B200 with nothing else in flight (tools/hop_bench.cu, mass polling of the data
# words, 2268 cycles at 1.965 GHz; profiles/r01_trace_attn_chain_bwd_v7.txt).  T x hand-offs x that latency is the floor of a
# persistent recurrent kernel whose weights are already on chip ("2 x grid-barrier latency" per step).
HAND_OFFS = {"attn_chain_fwd": 3, "attn_chain_bwd": 4, "dec_lstm_fwd": 1, "dec_lstm_bwd": 1, "enc_lstm_fwd": 1, "enc_lstm_bwd": 1}
HAND_OFF_US = 2268 / 1965.0


def latency_bound(kernel: str, ms_per_launch: float):
    kernel = kernel[:-4] if kernel.endswith("_grp") else (kernel[:-3] if kernel.endswith("_pt") else kernel)
    steps = L if kernel.startswith("enc_") else T
    bound_ms = steps * HAND_OFFS[kernel] * HAND_OFF_US * 1e-3
    return {"steps": steps, "hand_offs_per_step": HAND_OFFS[kernel], "hand_off_us": HAND_OFF_US, "bound_ms": bound_ms,
            "frac": bound_ms / ms_per_launch, "source": "tools/hop_bench.cu on B200: bare 148-CTA all-gather through L2"}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel)
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


OUTER_LR, CLIP = 1e-4, 1.0


def cpu_meta_step_fns(n_threads=None):
    """The FOMAML meta-step of the workload on the host CPU: the UNMODIFIED reference model / loss / mix_grad / apply_grad when
    baseline/_ref (oracle/install_reference.py) or /root/reference is importable ("reference"), else the oracle port ("port");
    the inner loop is the restatement of `higher` either way (oracle/ref_meta.py).  -> (task_fn, outer_fn, kind, threads)"""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from oracle import ref_meta
    threads = n_threads or ref_meta.host_threads()
    cfg = pkg.default_params()
    tasks = [synth.make_task(cfg, B, T, L, 1234 + i) for i in range(N_TASKS)]
    task_fn, outer_fn, kind = ref_meta.make_meta_step(cfg, tasks, INNER_LR, N_INNER, OUTER_LR, CLIP, threads)
    return task_fn, outer_fn, kind, threads


def run_reference(args):
    """--impl reference: WHOLE meta-steps (8 tasks x (inner SGD step + test forward/backward) + mix_grad / apply_grad / clip /
    Adam) of the reference on the host cores, rank 0 only, thread count set explicitly (torchrun exports OMP_NUM_THREADS=1).
    A meta-step takes ~15-20 s on the box's CPU, so the number of timed steps is bounded by a wall budget (MSA_REF_BUDGET_S,
    default 170 s) and the line prints the steps and warm-up steps that were REALLY run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("MSA_REF_BUDGET_S", "170"))
    t_start = time.perf_counter()
    task_fn, outer_fn, kind, threads = cpu_meta_step_fns()

    def meta_step():
        t0 = time.perf_counter()
        for i in range(N_TASKS):
            task_fn(i)
        outer_fn()
        return time.perf_counter() - t0
    warm = 1 if args.warmup >= 1 else 0
    t_warm = meta_step() if warm else 0.0
    ts = [meta_step()]
    while len(ts) < args.steps and (time.perf_counter() - t_start) + 1.1 * max(ts) < budget:
        ts.append(meta_step())
    t_step = sum(ts) / len(ts)
    value = 1.0 / t_step
    line = {"impl": "reference", "metric": "meta_steps_per_s", "value": value, "unit": "meta-steps/s", "n_gpus": args.gpus,
            "steps": len(ts), "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1000.0 * t_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tasks": N_TASKS, "inner_steps": N_INNER, "B": B, "T": T, "L": L},
            "cpu_baseline": {"value": value, "unit": "meta-steps/s", "cores": threads, "kind": kind,
                             "sample": f"whole meta-steps (8 tasks x (1 inner SGD step + test fwd/bwd) + mix_grad / apply_grad / clip / "
                                       f"Adam): {len(ts)} timed of {args.steps} requested after {warm} warm-up (wall budget {budget:.0f} s; "
                                       f"warm-up step took {t_warm:.1f} s)"},
            "e2e": {"value": value, "unit": "meta-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mel_frames_per_s": value * N_TASKS * 2 * B * T}
    print(json.dumps(line))


def cpu_infer_sample(B_=32, L_=64, steps=16):
    """BASELINE configs[4] on the host CPU: ``Tacotron2NV.infer`` of the unmodified reference model (B=32, L=64) for a bounded number
    of free-running decoder steps; mel-frames/s = rows x steps / time (encoder and postnet included, like the GPU leg)."""
    import copy
    import io
    from contextlib import redirect_stdout
    import torch
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from oracle import ref_meta
    ref = ref_meta.load_reference()
    if ref is None:
        return None
    threads = ref_meta.host_threads()
    torch.set_num_threads(threads)
    cfg = pkg.default_params()
    cfg["max_decoder_steps"] = steps
    cfg["decoder_no_early_stopping"] = True
    with redirect_stdout(io.StringIO()):
        model = ref[0](copy.deepcopy(cfg))
        sd = model.state_dict()
        for k, v in synth.init_params(cfg, 0).items():
            sd[k] = v.clone()
        model.load_state_dict(sd)
        model.eval()
    g = torch.Generator().manual_seed(4321)
    lens = torch.arange(L_, L_ - B_, -1)
    inp = torch.randint(1, 123, (B_, L_), generator=g)
    for b in range(B_):
        inp[b, lens[b]:] = 0
    spk = torch.randn(B_, cfg["speaker_embedding_dim"], generator=g)
    best = None
    with torch.no_grad(), redirect_stdout(io.StringIO()):
        for _ in range(2):
            t0 = time.perf_counter()
            out = model.infer(inp, lens, spk)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    n = int(out[0].shape[2])
    return {"value": B_ * n / best, "unit": "mel-frames/s", "us_per_step": best / n * 1e6, "cores": threads, "kind": "reference",
            "sample": f"Tacotron2NV.infer of the unmodified reference, B={B_}, L={L_}, {n} decoder steps (of the GPU leg's 1000), best of 2"}


def eager_gpu_reference_pass(torch, dev):
    """The 'library kernels on the same box' bar of SURVEY.md 8d: the unmodified reference model run by PyTorch eager ON THE SAME GPU,
    one forward + loss + backward pass at BASELINE configs[0] (B=4, T=200, L=64).  None if the reference is not importable."""
    import copy
    import io
    from contextlib import redirect_stdout
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from oracle import ref_meta
    ref = ref_meta.load_reference()
    if ref is None:
        return None
    cfg = pkg.default_params()
    with redirect_stdout(io.StringIO()):
        model = ref[0](copy.deepcopy(cfg))
        sd = model.state_dict()
        for k, v in synth.init_params(cfg, 0).items():
            sd[k] = v.clone()
        model.load_state_dict(sd)
    model.to(dev).train()
    crit = ref[1](cfg["n_frames_per_step"], "none", 10.0, dev)
    _, inp, inp_len, mels, mel_len, _, spk, stop = [x.to(dev) if hasattr(x, "to") else x for x in synth.make_batch(cfg, B, T, L, 100)]

    def one():
        model.zero_grad(set_to_none=True)
        out = model(inputs=inp, input_lengths=inp_len, melspecs=mels, melspec_lengths=mel_len, speaker_vecs=spk)
        crit(out, (mels, stop), mel_len).backward()
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        one()
    e1.record()
    torch.cuda.synchronize()
    return {"ms_per_pass": e0.elapsed_time(e1) / 3, "kind": "reference model, PyTorch eager (library kernels) on the same GPU"}


def cpu_baseline_sample():
    """Bounded sample (~10-30 s of CPU work) of the same meta-step for the `cpu_baseline` object of this repo's own line:
    1 warm-up task, 4 timed tasks and the outer update on their gradients; meta-step time = 8 x mean task + outer update.  Also the
    same sample under the reference's own 4-thread cap (utils/limit_threads.py:3-7)."""
    import torch
    task_fn, outer_fn, kind, threads = cpu_meta_step_fns()
    task_fn(0)
    ts = [task_fn(i) for i in range(1, 5)]
    t_outer = outer_fn()
    t_task = sum(ts) / len(ts)
    out = {"value": 1.0 / (N_TASKS * t_task + t_outer), "unit": "meta-steps/s", "cores": threads, "kind": kind,
           "sample": "4 of the 8 tasks (1 inner SGD step + test fwd/bwd each) timed after 1 warm-up task, plus mix_grad / apply_grad / "
                     "clip / Adam on their gradients; meta-step = 8 x mean task + outer update"}
    torch.set_num_threads(4)
    t4 = [task_fn(i) for i in range(5, 7)]
    t4_outer = outer_fn()
    torch.set_num_threads(threads)
    out["value_at_reference_thread_cap"] = {"value": 1.0 / (N_TASKS * (sum(t4) / len(t4)) + t4_outer), "cores": 4}
    return out


def bench_infer(torch, tr, world, sync, B=32, L_=64, steps=1000):
    """Free-running inference under both precision policies of the decoder LSTMCells: "tf32" (one TF32 product per step, inside the
    1e-3 the north star states for a TF32 path: tests/test_gpu_infer.py::test_infer_tf32_policy_within_stated_tolerance) is the
    headline of this leg, "fp32" (3xTF32 split, fp32-accurate to ~1e-6) is reported next to it."""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.engine import Engine
    cfg = pkg.default_params()
    cfg["max_decoder_steps"] = steps
    cfg["decoder_no_early_stopping"] = True
    lens = torch.arange(L_, L_ - B, -1)
    g = torch.Generator().manual_seed(4321)
    inp = torch.randint(1, 123, (B, L_), generator=g)
    for b in range(B):
        inp[b, lens[b]:] = 0
    spk = torch.randn(B, cfg["speaker_embedding_dim"], generator=g)
    pm = synth.make_infer_masks(cfg, B, steps, 5)
    res = {}
    for policy, code in (("tf32", 2), ("fp32", 0)):
        eng = Engine(cfg, tr.device, gemm_tf32=code)
        flat, bn = tr.theta, eng.new_bn_stats()
        eng.infer(flat, bn, inp, lens, spk, pm, max_steps=steps)          # warm-up (workspace, L2)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.infer(flat, bn, inp, lens, spk, pm, max_steps=steps)     # host tokens in, encoder + 1000 steps + postnet
        out[1].cpu()                                                       # D2H of mel_lengths
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=tr.device)
        tr.shard.allreduce_max(ms)
        n_steps = int(out[0].shape[2])
        res[policy] = {"value": world * B * n_steps / float(ms) * 1e3, "us_per_step": float(ms) * 1e3 / n_steps}
        del eng
    # stream bound of a decoder step: the 79.7 MB of fp32 LSTMCell / projection weights are read once per step
    # (L2-resident after the first step; the measured HBM copy peak stands in as the stream bound)
    from msa_tts_b200.config import memory_dim, rnn_dims
    Ha, Hd = rnn_dims(cfg)
    E = memory_dim(cfg)
    wbytes = 4.0 * (4 * Ha * (cfg["prenet_dim"] + E + Ha) + 4 * Hd * (Ha + E + Hd) + (cfg["n_mel_channels"] + 1) * (Hd + E) +
                    cfg["prenet_dim"] * (cfg["n_mel_channels"] + cfg["prenet_dim"]) + cfg["attention_params"]["attention_dim"] * Ha)
    peak, peak_src = measured_peak()
    roof = {"bound": "hbm", "achieved": wbytes / res["tf32"]["us_per_step"] / 1e3, "peak": peak, "unit": "GB/s",
            "frac": wbytes / res["tf32"]["us_per_step"] / 1e3 / peak, "traffic": None, "peak_source": peak_src,
            "algo_bytes_per_step": wbytes, "note": "decoder-step weights streamed once per step (L2-resident, five launches per step)"}
    return {"metric": "decoder_mel_frames_per_s", "value": res["tf32"]["value"], "unit": "mel-frames/s",
            "us_per_step": res["tf32"]["us_per_step"], "scaling": "weak", "gemm": "tf32 (rel 1e-3 path, tested)",
            "fp32_accurate": res["fp32"], "roofline": roof,
            "workload": f"free-running inference, B={B} per GPU, L={L_}, {n_steps} decoder steps, default dims, encoder and postnet included"}


def bench_other_configs(torch, tr, world, sync, gemm_mode):
    """BASELINE.json configs[0] (one forward+backward pass, B=4, T=200, 80 mels), configs[2] (Reptile meta-step, 16 tasks x 5 inner
    steps, batched variant, parameter-delta allreduce) and configs[3] (continual EWC step with the fused penalty + update kernel, ER-KD
    soft-target pass) on this repo's CUDA path; device-resident inputs, CUDA events, max over ranks."""
    import msa_tts_b200 as pkg
    from msa_tts_b200 import synth
    from msa_tts_b200.engine import batch_to_device
    from msa_tts_b200.reptile import Reptile
    eng, dev, cfg = tr.engine, tr.device, tr.model_params

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        tr.shard.allreduce_max(ms)
        return float(ms)
    out = {}
    bd = batch_to_device(synth.make_batch(cfg, B, T, L, 100), dev)
    masks = eng.generate_masks(B, T, L, 7)
    bn, g = eng.new_bn_stats(), tr.task_grad

    def one_pass():
        eng.forward(tr.fast, bn, bd, masks, outputs=False)
        eng.backward(tr.fast, g)
    ms = timed(one_pass, 10)
    out["config0_fwd_bwd_pass"] = {"ms_per_pass": ms, "mel_frames_per_s": world * B * T / ms * 1e3, "scaling": "replicas",
                                   "workload": f"one forward+backward pass per GPU, B={B}, T={T}, L={L}, default dims"}
    mu, fisher = tr.meta_grad, tr.outer_m          # any two flat buffers: the kernels are bandwidth-bound, the values do not matter
    fisher.abs_()

    def ewc_step():
        one_pass()
        eng.ewc_sgd_step(tr.fast, g, mu, fisher, 1e-9, 100.0)
    ms = timed(ewc_step, 10)
    ms_u = timed(lambda: eng.ewc_sgd_step(tr.fast, g, mu, fisher, 1e-9, 100.0), 20)
    n = eng.layout.total
    ms_kd = timed(lambda: eng.forward(tr.fast, bn, bd, masks, outputs=True), 10)
    out["config3_continual"] = {"ewc_step_ms": ms, "ewc_fused_penalty_update_ms": ms_u,
                                "ewc_fused_penalty_update_frac_of_hbm_peak": 20.0 * n / ms_u / 1e6 / measured_peak()[0],
                                "erkd_soft_target_pass_ms": ms_kd, "scaling": "replicas",
                                "workload": "EWC: forward+backward + fused penalty-gradient/SGD update; ER-KD: teacher-forced soft-target pass"}
    eng.check_abort()
    # Reptile: 16 tasks x 5 inner steps, all tasks from the same theta_0 (batched variant, the one that shards), one delta allreduce
    global N_TASKS
    keep = N_TASKS
    N_TASKS = 16
    try:
        params = trainer_params(gemm_mode)
        params["n_inner_train"] = 5
        params["reptile_sequential"] = False
        rp = Reptile(**params)
        items = make_tasks(cfg, pinned=False)
        items = {s_: {k: tuple(x.to(dev) if hasattr(x, "to") else x for x in v) for k, v in t_.items()} for s_, t_ in items.items()}
        ms = timed(lambda: rp._metatrain_step(items), 2, warm=1)
        rp.engine.abort_flush()
        rp.engine.check_abort()
        out["config2_reptile"] = {"ms_per_meta_step": ms, "meta_steps_per_s": 1e3 / ms, "scaling": "strong",
                                  "workload": "batched Reptile meta-step, 16 tasks x 5 inner SGD steps + test forward, parameter-delta allreduce"}
        del rp
    finally:
        N_TASKS = keep
    return out


def bench_gemm_tc(torch, eng, reps=20):
    import ctypes as C
    from msa_tts_b200.config import rnn_dims
    Ha, Hd = rnn_dims(eng.cfg)
    M, N, K = T * B, 4 * Hd, Ha
    lib = eng.lib
    dev = eng.device
    A_, B_, C_ = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
    scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device=dev)
    P = lambda t: C.c_void_p(t.data_ptr())

    def call():
        rc = lib.msa_gemm_nt(M, N, K, C.c_float(1.0), P(A_), K, P(B_), K, C.c_float(0.0), P(C_), N, 0, P(scratch),
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError(f"msa_gemm_nt: {rc}")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    side, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
            for _ in range(reps):
                call()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak_bf16 = 1636.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak_bf16 = float(json.load(open(p)).get("bf16_tflops", peak_bf16))
    flops = 2.0 * M * N * K
    return {"kernel": "gemm_tc_3xtf32", "shape": [M, N, K], "ms_per_launch": ms, "timed": "alone, CUDA graph of 20 launches",
            "fp32_equivalent_tflops": flops / ms / 1e9, "tensor_tflops": 3.0 * flops / ms / 1e9, "bound": "tensor",
            "peak_tflops": peak_bf16 / 2.0, "peak_source": "TF32 dense = half of the measured bf16 cuBLAS throughput (MEASURED_PEAKS.json)",
            "frac": 3.0 * flops / ms / 1e9 / (peak_bf16 / 2.0),
            "note": "3 TF32 MMAs per fp32-accurate product; bound by shared-memory operand bandwidth (DESIGN.md 4.5)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gemm", default="tf32x", choices=["fp32", "tf32x", "tf32"], help="GEMM precision policy (DESIGN.md section 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from msa_tts_b200.maml import MAML
    from msa_tts_b200.engine import batch_to_device

    gemm_mode = {"fp32": 0, "tf32x": 1, "tf32": 2}[args.gemm]
    params = trainer_params(gemm_mode)
    tr = MAML(**params)
    eng = tr.engine
    cfg = params["model"]
    host_items = make_tasks(cfg, pinned=True)
    dev_items = {s: {k: tuple(x.to(tr.device) if hasattr(x, "to") else x for x in v) for k, v in t.items()} for s, t in host_items.items()}
    mine = tr.shard.my_tasks(N_TASKS)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(items, steps, read_losses):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            log = tr._metatrain_step(items)
            if read_losses:
                log["loss_test"].cpu()            # D2H read of the step's result
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=tr.device)
        tr.shard.allreduce_max(ms)
        return float(ms)

    for _ in range(args.warmup):
        tr._metatrain_step(dev_items)
    # ---- device-resident leg (value) with live per-kernel event timing and clock sampling ----
    l0 = eng.kernel_launches()
    eng.profile(True)
    with ClockSampler(local_rank) as clk:
        ms_dev = timed(dev_items, args.steps, read_losses=False)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.kernel_launches() - l0
    # ---- end-to-end leg: pinned host batches, H2D inside, D2H of the losses ----
    for _ in range(2):
        tr._metatrain_step(host_items)
    ms_e2e = timed(host_items, args.steps, read_losses=True)
    h2d = sum(batch_bytes(host_items[f"spk{i}"][k]) for i in mine for k in ("train", "test"))
    d2h = 4 * len(mine)
    eng.abort_flush()          # a persistent kernel that gave up polling invalidates the numbers: raise here
    eng.check_abort()

    # ---- flat-buffer kernels timed alone (HBM roofline) ----
    def time_flat(fn, reps=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    n = eng.layout.total
    flat_ms = {
        "flat_sgd_step": (time_flat(lambda: eng.sgd_step(tr.fast, tr.task_grad, lr=1e-3)), 12.0 * n),
        "flat_clip_adam": (time_flat(lambda: eng.clip_adam(tr.fast, tr.task_grad, tr.outer_m, tr.outer_v, tr.sumsq, lr=1e-9, step=1, max_norm=1.0)), 28.0 * n),
        "flat_sumsq": (time_flat(lambda: eng.sumsq(tr.task_grad)), 4.0 * n),
    }

    # ---- the tensor-core shaped kernel of the pass: the hand-written tcgen05 / TMA GEMM on the decoder-RNN gate product
    # [T*B x 4Hd] x K = Ha in its fp32-accurate 3xTF32 mode, timed alone (CUDA graph of 20 calls: GPU time, no host gaps) ----
    gemm_line = None
    if world == 1:          # (single-GPU evidence line; no graph capture next to the NCCL watchdog thread)
        try:
            gemm_line = bench_gemm_tc(torch, eng)
        except Exception as e:
            gemm_line = {"kernel": "gemm_tc_3xtf32", "error": str(e)[:200]}

    # ---- decoder mel-frames/s of free-running inference (BASELINE configs[4]: B=32 per GPU, L=64, 1000 steps, no early stop);
    # every rank decodes its own batch (inference shards by batch rows, no collective), time = max over ranks ----
    infer_line = None
    try:
        infer_line = bench_infer(torch, tr, world, sync)
    except Exception as e:      # the meta-step line must not depend on this leg
        infer_line = {"error": str(e)[:200]}

    # ---- the other BASELINE configs as extra legs (configs[0], [2], [3]; the headline stays configs[1]) ----
    other = None
    try:
        other = bench_other_configs(torch, tr, world, sync, gemm_mode)
    except Exception as e:
        other = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak()
        kern = []
        for name, (ms, cnt) in prof.items():
            if cnt:
                rows = B * (eng.group_size(len(mine), B) if name.endswith("_grp") else 1)
                if name.endswith("_pt"):      # several tasks with their own weights per launch (3 + 3 + 2 for 8 local tasks)
                    rows = int(round(B * len(mine) * args.steps / cnt))
                ab = algo_bytes(cfg, name, rows)
                per = ms / cnt
                kern.append({"kernel": name, "rows_per_launch": rows, "us_per_step_per_row": per * 1e3 / (L if name.startswith("enc_") else T) / rows,
                             "launches_per_step": cnt / args.steps, "ms_per_launch": per,
                             "share_of_step": ms / ms_dev, "algo_bytes": ab, "achieved_gbs": ab / per / 1e6,
                             "frac": ab / per / 1e6 / peak, "latency_bound": latency_bound(name, per)})
                if name.endswith("_pt"):
                    # the launch streams every task's recurrent weights as bf16 hi/lo fragments from L2 on EVERY step
                    # (148 CTAs x 16 warps x ceil(Ha/256) k16 steps x 2 KB per task, chain_mma.cu): the stream that bounds it
                    from msa_tts_b200.config import rnn_dims
                    frag = 148 * 16 * -(-rnn_dims(cfg)[0] // 256) * 2048.0
                    kern[-1]["l2_stream"] = {"bytes_per_step": frag * rows / B, "achieved_gbs": frag * rows / B * T / per / 1e6,
                                             "note": "weight fragments of every task re-read from L2 each step (per-task weights cannot stay in shared memory)"}
        for name, (ms, ab) in flat_ms.items():
            kern.append({"kernel": name, "ms_per_launch": ms, "algo_bytes": ab, "achieved_gbs": ab / ms / 1e6,
                         "frac": ab / ms / 1e6 / peak, "timed": "alone, 20 reps"})
        if gemm_line is not None:
            kern.append(gemm_line)
        dom = max((k for k in kern if "share_of_step" in k), key=lambda k: k["share_of_step"])
        ms_step = ms_dev / args.steps
        value = 1000.0 / ms_step
        e2e_value = 1000.0 * args.steps / ms_e2e
        line = {
            "metric": "meta_steps_per_s", "value": value, "unit": "meta-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "tasks": N_TASKS, "inner_steps": N_INNER, "B": B, "T": T, "L": L,
                       "params": eng.layout.n_params, "gemm": args.gemm, "parallelism": f"task-shard x{world} + 1 allreduce",
                       "l2": "per-step working set (4 flat buffers 485 MB + workspace) > 126 MB L2, no explicit flush"},
            "mel_frames_per_s": value * N_TASKS * 2 * B * T,
            "e2e": {"value": e2e_value, "unit": "meta-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": dom["frac"], "traffic": ncu_traffic(dom["kernel"]), "peak_source": peak_src,
                         "note": "persistent recurrent kernel: the weight slices stay in shared memory for all T steps; its time is "
                                 "set by the cross-CTA hand-offs of every decoder step (L2 round trips among 148 CTAs) and, for "
                                 "the grouped launches, by the L2 -> SM streams of h / dz / MW -- not by HBM or tensor throughput "
                                 "(DESIGN.md section 4.2, profiles/r02_*)",
                         "ms_per_launch": dom["ms_per_launch"], "share_of_step": dom["share_of_step"],
                         "latency_bound": dom["latency_bound"]},
            "kernels": kern,
            "clocks": clk.summary(),
            "infer": infer_line,
            "other_configs": other,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
            try:
                if isinstance(infer_line, dict) and "error" not in infer_line:
                    infer_line["cpu_baseline"] = cpu_infer_sample()
                if isinstance(other, dict) and "config0_fwd_bwd_pass" in other:
                    other["config0_fwd_bwd_pass"]["eager_gpu_baseline"] = eager_gpu_reference_pass(torch, tr.device)
            except Exception as e:
                line["baseline_legs_error"] = str(e)[:200]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
