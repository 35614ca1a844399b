"""Decode what the MN-major UMMA descriptors read (debug)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fsw_gnn_b200 import ops
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=250, sci_mode=False)
M, N, Kd = 2048, 128, 32
A = torch.zeros(M, Kd, device=dev)
for k in range(Kd):
    A[k, k] = 1.0
B = (torch.arange(Kd, device=dev)[:, None] * 1000 + torch.arange(N, device=dev)[None, :]).float()
expect = B.clone()
for cfg in ["16 1024 32 4096 512 1024 0 0", "16 1024 32 512 4096 1024 0 0", "16 1024 32 4096 1024 1024 0 0", "16 1024 32 1024 4096 1024 0 0"]:
    os.environ["FSW_UMMA_DBG"] = cfg
    C = ops.gemm(1, A, B, M, N, Kd, Kd, N)
    torch.cuda.synchronize()
    ok = bool((C[:Kd] == expect).all())
    print("NN cfg", cfg, "exact" if ok else "WRONG", " nonzero frac %.3f" % float((C[:Kd] != 0).float().mean()))
    if not ok:
        print(C[:12, :10].long())
        print(C[:12, 30:40].long())
# TN: A [Kd, M], B [Kd, N]: C[m, n] = sum_k A[k, m] B[k, n]
Kd, M, N = 4096, 128, 128
A = torch.zeros(Kd, M, device=dev)
A[:32] = (torch.arange(32, device=dev)[:, None] * 1000 + torch.arange(M, device=dev)[None, :]).float()
B = torch.zeros(Kd, N, device=dev)
for k in range(32):
    B[k, k] = 1.0
expect = A[:32].t().contiguous()   # C[m, n<32] = A[n, m]
for cfg in ["4096 512 1024 4096 512 1024 0 0", "512 4096 1024 4096 512 1024 0 0"]:
    os.environ["FSW_UMMA_DBG"] = cfg
    C = torch.zeros(M, N, device=dev)
    ops.gemm(2, A, B, M, N, Kd, M, N, out=C, ldc=N, accumulate=True)
    torch.cuda.synchronize()
    ok = bool((C[:, :32] == expect).all())
    print("TN cfg", cfg, "exact" if ok else "WRONG", " nonzero frac %.3f" % float((C[:, :32] != 0).float().mean()))
    if not ok:
        print(C[:12, :10].long())
        print(C[30:42, :10].long())
