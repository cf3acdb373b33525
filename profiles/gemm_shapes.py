"""Every GEMM shape of one forward+backward pass (env MSA_GEMM_LOG) timed on the hand-written tcgen05 kernel (msa_gemm) and on
cuBLAS (torch.matmul, TF32 allowed for the backward shapes), operands L2-warm, CUDA events over `reps` back-to-back calls.
    python profiles/gemm_shapes.py [reps]"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from msa_tts_b200 import _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
env = dict(os.environ, MSA_GEMM_LOG="1")
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "pass_time.py"), "1"], env=env, capture_output=True, text=True).stderr
shapes = collections.OrderedDict()
for m in re.finditer(r"gemm ta=(\d) tb=(\d) M=(\d+) N=(\d+) K=(\d+) bwd=(\d) tf32=(\d) own=(\d)", out):
    key = tuple(int(x) for x in m.groups())
    shapes[key] = shapes.get(key, 0) + 1
n_pass = 4          # pass_time.py runs 3 warm-up passes + 1 timed
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print(f"{'ta tb':5s} {'M':>6s} {'N':>6s} {'K':>6s} bwd n/pass  cuBLAS us   own us (mode)  routed")
tot = collections.defaultdict(float)
for (ta, tb, M, N, K, bwd, tf32, own), cnt in shapes.items():
    cnt //= n_pass
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((N, K) if tb else (K, N), device="cuda")
    Cm = torch.zeros(M, N, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    Aop, Bop = (A.t() if ta else A), (B.t() if tb else B)
    t_blas = timed(lambda: torch.matmul(Aop, Bop, out=Cm))
    t_own = float("nan")
    mode = 2 if tf32 else 0
    if A.shape[1] % 4 == 0 and B.shape[1] % 4 == 0:
        scratch = torch.empty(int(lib.msa_gemm_nt_scratch_floats(M, N, K)) + 4, device="cuda")
        t_own = timed(lambda: _lib.check(lib.msa_gemm(ta, tb, M, N, K, C.c_float(1.0), P(A), A.shape[1], P(B), B.shape[1], C.c_float(0.0), P(Cm), N,
                                                      mode, P(scratch), st), "msa_gemm"))
    print(f"{ta}  {tb}   {M:6d} {N:6d} {K:6d}  {bwd}   {cnt:3d}   {t_blas:9.1f}  {t_own:9.1f} ({mode})     {own}")
    tot["blas"] += cnt * t_blas
    if t_own == t_own:
        tot["own"] += cnt * t_own
        tot["best"] += cnt * min(t_blas, t_own)
    else:
        tot["own"] += cnt * t_blas
        tot["best"] += cnt * t_blas
print(f"per pass: all cuBLAS {tot['blas']:.0f} us, all own {tot['own']:.0f} us, best of both {tot['best']:.0f} us")
