"""K1 on the tensor cores (csrc/fsw_umma.cu: tcgen05.mma kind::tf32, TMA, hi/lo split) against fp64 matmul of the same fp32
inputs, beside the FMA kernels it replaces for large shapes.  Tolerance: the one test_gemm_strip_kernels_vs_fp64 holds the FMA
kernels to (2e-6 of the largest |result| for NT, 1e-6 / 4e-6 of the largest sum of absolute terms for NN / TN)."""
import numpy as np
import pytest
import torch

from parity import LOG

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _both(fn):
    """run fn with the tensor-core path and with the FMA kernels; returns (tc, fma)"""
    from fsw_gnn_b200 import _lib
    lib = _lib.load()
    try:
        lib.fsw_set_tensor_cores(1)
        a = fn()
        lib.fsw_set_tensor_cores(0)
        b = fn()
    finally:
        lib.fsw_set_tensor_cores(1)
    return a, b


def _ran_umma(fn, label):
    from fsw_gnn_b200 import _lib
    _lib.profile_enable(True)
    _lib.profile_read()
    try:
        fn()
        torch.cuda.synchronize()
        rec = _lib.profile_read()
    finally:
        _lib.profile_enable(False)
    return label in rec


@pytest.mark.parametrize("shape", [(20000, 199, 100), (4096, 511, 256), (3001, 64, 64), (2500, 100, 300), (70000, 256, 32)])
def test_umma_nt_vs_fp64(shape):
    from fsw_gnn_b200 import ops
    M, N, Kd = shape
    g = torch.Generator(device=dev()); g.manual_seed(M + N)
    A = torch.randn(M, Kd, device=dev(), generator=g)
    B = torch.randn(N, Kd, device=dev(), generator=g)
    ldc = (N + 7) // 8 * 8

    def run():
        C = torch.full((M, ldc), 7.0, device=dev())
        ops.gemm(0, A, B, M, N, Kd, Kd, Kd, out=C, ldc=ldc)
        return C
    assert _ran_umma(run, "umma_nt"), "the tensor-core kernel did not run for %s" % (shape,)
    C, Cf = _both(run)
    ref = A.double() @ B.double().T
    e_tc = float((C[:, :N].double() - ref).abs().max()) / float(ref.abs().max())
    e_fma = float((Cf[:, :N].double() - ref).abs().max()) / float(ref.abs().max())
    LOG.append("umma NT %s: max err / max|ref| tensor cores %.2e, FMA kernels %.2e" % (shape, e_tc, e_fma))
    assert e_tc <= 2e-6
    # padding columns: untouched, except that the TMA store's 16-byte granule zero-fills up to the next multiple of 4 columns
    n4 = (N + 3) // 4 * 4
    assert bool((C[:, n4:] == 7.0).all()) and bool(((C[:, N:n4] == 7.0) | (C[:, N:n4] == 0.0)).all())
    # accumulate (TMA reduce-add)
    C2 = C.clone()
    ops.gemm(0, A, B, M, N, Kd, Kd, Kd, out=C2, ldc=ldc, accumulate=True)
    assert float((C2[:, :N].double() - 2 * ref).abs().max()) <= 4e-6 * float(ref.abs().max())


@pytest.mark.parametrize("shape", [(20000, 100, 199), (4096, 256, 511), (2100, 300, 100), (5000, 50, 37)])
def test_umma_nn_vs_fp64(shape):
    """dX = dXp . theta: A [M, Kd] with a padded leading dimension, B [Kd, N] row-major (MN-major operand)"""
    from fsw_gnn_b200 import ops
    M, N, Kd = shape
    g = torch.Generator(device=dev()); g.manual_seed(M + N + 1)
    lda = (Kd + 7) // 8 * 8
    A = torch.zeros(M, lda, device=dev()); A[:, :Kd] = torch.randn(M, Kd, device=dev(), generator=g)
    if N % 4 != 0:
        pytest.skip("ldb must be a multiple of 4 for TMA")
    B = torch.randn(Kd, N, device=dev(), generator=g)

    def run():
        return ops.gemm(1, A, B, M, N, Kd, lda, N)
    assert _ran_umma(run, "umma_nn")
    D, Df = _both(run)
    ref = A[:, :Kd].double() @ B.double()
    scale = float((A[:, :Kd].double().abs() @ B.double().abs()).max())
    e_tc = float((D.double() - ref).abs().max()) / scale
    e_fma = float((Df.double() - ref).abs().max()) / scale
    LOG.append("umma NN %s: max err / max sum|terms| tensor cores %.2e, FMA kernels %.2e" % (shape, e_tc, e_fma))
    assert e_tc <= 1e-6


@pytest.mark.parametrize("shape", [(199, 100, 50000), (511, 256, 20000), (100, 300, 8000), (64, 64, 4100)])
def test_umma_tn_vs_fp64(shape):
    """dtheta += dXp^T . X: reduction over the long row axis, split over CTAs, TMEM drained every 16 k-blocks"""
    from fsw_gnn_b200 import ops
    M, N, Kd = shape
    g = torch.Generator(device=dev()); g.manual_seed(M + N + 2)
    lda = (M + 7) // 8 * 8
    A = torch.zeros(Kd, lda, device=dev()); A[:, :M] = torch.randn(Kd, M, device=dev(), generator=g)
    B = torch.randn(Kd, N, device=dev(), generator=g)

    def run():
        T = torch.zeros(M, N, device=dev())
        ops.gemm(2, A, B, M, N, Kd, lda, N, out=T, ldc=N, accumulate=True)
        return T
    assert _ran_umma(run, "umma_tn")
    T, Tf = _both(run)
    ref = A[:, :M].double().T @ B.double()
    scale = float((A[:, :M].double().abs().T @ B.double().abs()).max())
    e_tc = float((T.double() - ref).abs().max()) / scale
    e_fma = float((Tf.double() - ref).abs().max()) / scale
    LOG.append("umma TN %s: max err / max sum|terms| tensor cores %.2e, FMA kernels %.2e" % (shape, e_tc, e_fma))
    assert e_tc <= 4e-6


def test_umma_fused_concat_bias_vs_fp64():
    """cat(emb, x) . W^T + b as two contraction segments (fsw_conv.py:357-361), W sliced by views"""
    from fsw_gnn_b200 import ops
    M, K0, K1, N = 30000, 200, 100, 100
    g = torch.Generator(device=dev()); g.manual_seed(9)
    emb = torch.randn(M, K0, device=dev(), generator=g)
    x = torch.randn(M, K1, device=dev(), generator=g)
    W = torch.randn(N, K0 + K1, device=dev(), generator=g)
    b = torch.randn(N, device=dev(), generator=g)

    def run():
        return ops.gemm_fused([emb, x], [W[:, :K0], W[:, K0:]], bias=b)
    assert _ran_umma(run, "umma_nt")
    out, outf = _both(run)
    ref = torch.cat((emb, x), dim=1).double() @ W.double().T + b.double()
    e_tc = float((out.double() - ref).abs().max()) / float(ref.abs().max())
    e_fma = float((outf.double() - ref).abs().max()) / float(ref.abs().max())
    LOG.append("umma fused concat+bias: max err / max|ref| tensor cores %.2e, FMA kernels %.2e" % (e_tc, e_fma))
    assert e_tc <= 2e-6 and e_fma <= 2e-6


def test_umma_long_contraction_error_growth():
    """accuracy of the TMEM accumulation over a long contraction (NT, Kd = 8192): error relative to the sum of |terms|"""
    from fsw_gnn_b200 import ops
    M, N, Kd = 4096, 64, 8192
    g = torch.Generator(device=dev()); g.manual_seed(5)
    A = torch.rand(M, Kd, device=dev(), generator=g) + 0.5      # all positive: the accumulator grows monotonically
    B = torch.rand(N, Kd, device=dev(), generator=g) + 0.5

    def run():
        return ops.gemm(0, A, B, M, N, Kd, Kd, Kd)
    C, Cf = _both(run)
    ref = A.double() @ B.double().T
    e_tc = float(((C.double() - ref) / ref).abs().max())
    b_tc = float(((C.double() - ref) / ref).mean())
    e_fma = float(((Cf.double() - ref) / ref).abs().max())
    LOG.append("umma NT Kd=8192 positive terms: max rel err tensor cores %.2e (mean signed %.2e), FMA kernels %.2e" % (e_tc, b_tc, e_fma))
    assert e_tc <= 2e-5


@pytest.mark.parametrize("M", [300, 6000])
@pytest.mark.parametrize("widths", [(200, 100), (127, 64), (72,)])
def test_linear_cat_autograd_vs_torch_fp64(M, widths):
    """FSW_conv combine: Linear over cat(inputs) with the concatenation fused into the contraction; value and all gradients
    against torch in fp64 (small M: FMA kernels, large M: tensor cores)"""
    from fsw_gnn_b200 import ops
    g = torch.Generator(device=dev()); g.manual_seed(M + sum(widths))
    N = 100
    xs = [torch.randn(M, w, device=dev(), generator=g).requires_grad_(True) for w in widths]
    W = (torch.randn(N, sum(widths), device=dev(), generator=g) / 10).requires_grad_(True)
    b = torch.randn(N, device=dev(), generator=g).requires_grad_(True)
    gout = torch.randn(M, N, device=dev(), generator=g)
    out = ops.linear_cat(xs, W, b)
    (out * gout).sum().backward()
    xs64 = [x.detach().double().requires_grad_(True) for x in xs]
    W64, b64 = W.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    ref = torch.nn.functional.linear(torch.cat(xs64, dim=1), W64, b64)
    (ref * gout.double()).sum().backward()

    def rel(a, r):
        return float((a.double() - r).abs().max()) / max(float(r.abs().max()), 1e-30)
    errs = dict(out=rel(out, ref), dW=rel(W.grad, W64.grad), db=rel(b.grad, b64.grad))
    for i, (x, x64) in enumerate(zip(xs, xs64)):
        errs["dx%d" % i] = rel(x.grad, x64.grad)
    LOG.append("linear_cat M=%d widths=%s: max err / max|ref| %s" % (M, widths, {k: "%.1e" % v for k, v in errs.items()}))
    assert all(v <= 3e-6 for v in errs.values()), errs


def test_fsw_conv_large_graph_combine_runs_on_tensor_cores():
    """an FSW_conv step on a graph large enough for the tensor-core path: K1 + combine launch the UMMA kernel, and the result
    equals the FMA-kernel result to fp32 rounding"""
    from fsw_gnn_b200 import FSW_conv, _lib
    torch.manual_seed(1)
    N, E, d = 6000, 60000, 64
    conv = FSW_conv(d, d, device=dev())
    x = torch.randn(N, d, device=dev())
    ei = torch.randint(0, N, (2, E), device=dev())
    lib = _lib.load()

    def run():
        xx = x.clone().requires_grad_(True)
        for p_ in conv.parameters():
            p_.grad = None
        out = conv(xx, ei)
        out.square().sum().backward()
        return out.detach(), xx.grad, conv.mlp[0].weight.grad.clone(), conv.fsw_embed.projVecs.grad.clone()
    _lib.profile_enable(True)
    _lib.profile_read()
    a = run()
    torch.cuda.synchronize()
    rec = _lib.profile_read()
    _lib.profile_enable(False)
    assert "umma_nt" in rec and "umma_nn" in rec and "umma_tn" in rec, sorted(rec)
    try:
        lib.fsw_set_tensor_cores(0)
        b = run()
    finally:
        lib.fsw_set_tensor_cores(1)
    # the two projections differ in the last bits of the keys, so pairs of keys that agree to within an ulp may sort the other
    # way round: the value is continuous there, the gradient of the two affected source rows is not (tests/parity.py, KEYS)
    for name, u, v in zip(("out", "dx", "dW_mlp", "dtheta"), a, b):
        d = (u - v).abs()
        lim = 1e-6 + 1e-5 * v.abs() + 1e-5 * float(v.abs().max())
        frac = float((d > lim).float().mean())
        LOG.append("FSW_conv N=6000 tensor cores vs FMA kernels, %s: max diff / max %.2e, entries beyond tolerance %.4f%%"
                   % (name, float(d.max()) / float(v.abs().max()), 100 * frac))
        assert frac <= (0.0 if name == "out" else 2e-3), (name, frac)
