"""Development probe for operand layouts of msa_gemm: C = A . B with A = [I | 0] shows which elements of B the tensor core reads."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msa_tts_b200 import _lib
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = 128, 128, 32
for ta, tb in ((0, 0), (1, 1), (1, 0)):
    B = (torch.arange(K)[:, None] * 1000 + torch.arange(N)[None, :]).float().cuda() if not tb else (torch.arange(N)[:, None] + 1000 * torch.arange(K)[None, :]).float().cuda()
    A = torch.zeros(K, M, device="cuda") if ta else torch.zeros(M, K, device="cuda")
    for k in range(K):
        if ta: A[k, k] = 1.0
        else: A[k, k] = 1.0
    Cm = torch.full((M, N), -1.0, device="cuda")
    rc = lib.msa_gemm(ta, tb, M, N, K, C.c_float(1.0), P(A), A.shape[1], P(B), B.shape[1], C.c_float(0.0), P(Cm), N, 1, None, st)
    torch.cuda.synchronize()
    print("ta,tb", ta, tb, "rc", rc)
    print(Cm[:4, :8].cpu())
    print(Cm[8:10, 30:36].cpu(), Cm[31, 120:].cpu())
