"""K1 / MLP contractions at the configs[3] and configs[4] shapes: tensor-core kernel (fsw_umma.cu) against the FMA kernels and
torch (cuBLAS fp32).  CUDA events, 3 warm-up + 10 timed launches, inputs far larger than L2."""
import os, sys, json
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fsw_gnn_b200 import ops, _lib

dev = torch.device("cuda:0")
lib = _lib.load()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(name, fn_tc, fn_torch, flops, bytes_):
    lib.fsw_set_tensor_cores(1)
    t_tc = timeit(fn_tc)
    lib.fsw_set_tensor_cores(0)
    t_fma = timeit(fn_tc)
    lib.fsw_set_tensor_cores(1)
    t_torch = timeit(fn_torch) if fn_torch is not None else float("nan")
    print(json.dumps(dict(case=name, ms_tensor_core=round(t_tc, 4), ms_fma=round(t_fma, 4), ms_torch=round(t_torch, 4),
                          tflops_tc=round(flops / t_tc / 1e9, 2), gbs_tc=round(bytes_ / t_tc / 1e6, 1),
                          frac_hbm_6536=round(bytes_ / t_tc / 1e6 / 6536, 3))), flush=True)


torch.manual_seed(0)
for (N, d, K, tag) in [(2_400_000, 100, 199, "C4"), (1_000_000, 256, 511, "C5")]:
    ldp = (K + 7) // 8 * 8
    X = torch.randn(N, d, device=dev)
    theta = torch.randn(K, d, device=dev)
    Xp = torch.zeros(N, ldp, device=dev)
    report("%s project NT [%d,%d]x[%d,%d]^T" % (tag, N, d, K, d), lambda: ops.gemm(0, X, theta, N, K, d, d, d, out=Xp, ldc=ldp),
           lambda: torch.matmul(X, theta.t()), 2.0 * N * d * K, 4.0 * N * (d + ldp))
    dXp = torch.randn(N, ldp, device=dev)
    dX = torch.empty(N, d, device=dev)
    report("%s dX NN [%d,%d]x[%d,%d]" % (tag, N, K, K, d), lambda: ops.gemm(1, dXp, theta, N, d, K, ldp, d, out=dX, ldc=d),
           lambda: torch.matmul(dXp[:, :K], theta), 2.0 * N * d * K, 4.0 * N * (d + ldp))
    dth = torch.zeros(K, d, device=dev)
    report("%s dtheta TN [%d,%d]^Tx[%d,%d]" % (tag, N, K, N, d), lambda: ops.gemm(2, dXp, X, K, d, N, ldp, d, out=dth, ldc=d, accumulate=True),
           lambda: torch.matmul(dXp[:, :K].t(), X), 2.0 * N * d * K, 4.0 * N * (d + ldp))
    del X, Xp, dXp, dX
    torch.cuda.empty_cache()

# FSW_conv combine at C4: cat(emb [N,200], x [N,100]) . W[100,300]^T + b
N = 2_400_000
emb = torch.randn(N, 200, device=dev)
x = torch.randn(N, 100, device=dev)
W = torch.randn(100, 300, device=dev)
b = torch.randn(100, device=dev)
out = torch.empty(N, 100, device=dev)
report("C4 combine fused cat+Linear [N,200|100]x[100,300]^T+b", lambda: ops.gemm_fused([emb, x], [W[:, :200], W[:, 200:]], bias=b, out=out),
       lambda: torch.nn.functional.linear(torch.cat((emb, x), dim=1), W, b), 2.0 * N * 300 * 100, 4.0 * N * 400)
