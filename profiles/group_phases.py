"""Per-phase cycles of the grouped persistent kernels (thread 0 of every CTA, summed over the T steps of a launch) for a group
of G tasks at the bench dimensions.    python profiles/group_phases.py 8
With "pt" among the arguments every task gets its own copy of the weights (the per-task-weight kernels of chain_mma.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import msa_tts_b200 as pkg
from msa_tts_b200 import synth
from msa_tts_b200.engine import Engine, batch_to_device

cfg = pkg.default_params()
dev = torch.device("cuda:0")
B, T, L = 4, 200, 64
eng = Engine(cfg, dev, gemm_tf32=1)
flat = eng.flat_from_dict(synth.init_params(cfg, 0))
PT = "pt" in sys.argv[1:]
EVENTS_ONLY = "events" in sys.argv[1:]      # plain kernels, CUDA-event times only
for G in [int(a) for a in sys.argv[1:] if a not in ("pt", "events")] or [8]:
    bds = [batch_to_device(synth.make_batch(cfg, B, T, L, 100 + g), dev) for g in range(G)]
    masks = [eng.generate_masks(B, T, L, 7 + g) for g in range(G)]
    bn = [eng.new_bn_stats() for _ in range(G)]
    grads = [eng.new_flat() for _ in range(G)]
    params = [flat.clone() for _ in range(G)] if PT else flat
    for _ in range(2):
        eng.forward_group(params, bn, bds, masks)
        eng.backward_group(params, grads)
    eng.profile(True, inkernel=not EVENTS_ONLY)
    eng.forward_group(params, bn, bds, masks)
    eng.backward_group(params, grads)
    torch.cuda.synchronize()
    eng.check_abort()
    ev = eng.profile_read()
    steps = {"enc_lstm_fwd": L, "enc_lstm_bwd": L, "attn_chain_fwd": T, "attn_chain_bwd": T, "dec_lstm_fwd": T, "dec_lstm_bwd": T}
    for name, (ms, cnt) in ev.items():
        if EVENTS_ONLY:
            if cnt:
                print(f"G={G} {name}: {ms * 1e3:.0f} us in {cnt} launches")
            continue
        ph = eng.profile_phases(name)
        act = [r for r in ph if sum(r) > 0]
        if not act:
            print(f"G={G} {name}: {ms * 1e3:.0f} us (no phase counters)")
            continue
        n = steps[name[:-4] if name.endswith('_grp') else (name[:-3] if name.endswith('_pt') else name)]
        tot = [sum(r) for r in act]
        print(f"G={G} {name}: {ms * 1e3:.0f} us (events), {len(act)} CTAs reporting, {sum(tot) / len(tot) / n:.0f} cycles/step")
        for j in range(8):
            col = [r[j] / n for r in act]
            if max(col) > 0:
                print(f"   phase {j}: mean {sum(col) / len(col):8.0f}  min {min(col):8.0f}  max {max(col):8.0f}  cycles/step")
    eng.profile(False)
