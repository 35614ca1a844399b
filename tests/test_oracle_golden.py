"""Pin the CPU oracle (oracle/fsw_oracle.py) against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  fp64 reference outputs are the truth; tolerance 1e-10."""
import numpy as np
import pytest

from conftest import coo_to_csr, load_golden
from oracle import fsw_oracle as O

TOL = dict(rtol=1e-9, atol=1e-10)


def emb_params(g):
    return dict(projVecs=g["param_projVecs"], freqs=g["param_freqs"], bias=g.get("param_bias"),
                total_mass_encoding_scale=g.get("param_total_mass_encoding_scale"))


DENSE_CASES = {
    "emb_dense_weighted": {},
    "emb_dense_unit": {},
    "emb_dense_uniform": {},
    "emb_dense_deficient_tm": dict(encode_total_mass=True, total_mass_encoding_function="sqrt"),
    "emb_dense_n1": {},
    "emb_dense_big": {},
    "emb_dense_homog_alt": dict(encode_total_mass=True, total_mass_encoding_method="homog_alt"),
}


def dense_inputs(g):
    X = g["X"]
    bd = X.shape[:-2]
    n, d = X.shape[-2:]
    B = int(np.prod(bd)) if bd else 1
    Xf = X.reshape(B * n, d)
    mode = str(g["Wmode"])
    if mode == "unit":
        W = None
    elif mode == "uniform":
        W = np.full(B * n, 1.0 / n)
    else:
        W = g["W"].reshape(B * n)
    rowptr, col = O.dense_to_csr(B, n)
    return Xf, rowptr, col, W, bd, n, d


@pytest.mark.parametrize("name", list(DENSE_CASES))
@pytest.mark.parametrize("form", ["prod", "diff"])
def test_dense_forward(name, form):
    g = load_golden(name)
    Xf, rowptr, col, W, bd, n, d = dense_inputs(g)
    out = O.fsw_embedding_forward(Xf, rowptr, col, W, emb_params(g), DENSE_CASES[name], form=form)
    np.testing.assert_allclose(out.reshape(g["out_f64"].shape), g["out_f64"], **TOL)


@pytest.mark.parametrize("name", ["emb_dense_weighted", "emb_dense_unit", "emb_dense_uniform", "emb_dense_n1"])
def test_dense_backward(name):
    """Closed-form backward (no total-mass channel in these cases) vs reference autograd."""
    g = load_golden(name)
    Xf, rowptr, col, W, bd, n, d = dense_inputs(g)
    K = g["param_projVecs"].shape[0]
    gout = g["gout"].reshape(-1, g["gout"].shape[-1])
    assert gout.shape[1] == K
    Wb = W
    if str(g["Wmode"]) == "unit":
        Wb = None
    r = O.fsw_embed_csr_backward(Xf, rowptr, col, Wb, g["param_projVecs"], g["param_freqs"], gout)
    np.testing.assert_allclose(r["dX"].reshape(g["dX_f64"].shape), g["dX_f64"], **TOL)
    np.testing.assert_allclose(r["dtheta"], g["dprojVecs_f64"], **TOL)
    np.testing.assert_allclose(r["dxi"], g["dfreqs_f64"], rtol=1e-8, atol=1e-9)
    if "dW_f64" in g:
        np.testing.assert_allclose(r["dW"].reshape(g["dW_f64"].shape), g["dW_f64"], rtol=1e-8, atol=1e-9)


def test_weight_gradient_at_exact_threshold():
    """rows with total mass below, EXACTLY at and above the pad threshold in one batch: the reference pads every row once
    one is deficient, and the row at the threshold then feels the (zero-weight) pad element in its weight gradient"""
    g = load_golden("emb_dense_exact_thresh")
    Xf, rowptr, col, W, bd, n, d = dense_inputs(g)
    T = W.reshape(-1, n).sum(axis=1)
    assert (T == 1.0).sum() >= 2 and (T < 1.0).any() and (T > 1.0).any()
    out = O.fsw_embedding_forward(Xf, rowptr, col, W, emb_params(g), {})
    np.testing.assert_allclose(out.reshape(g["out_f64"].shape), g["out_f64"], **TOL)
    r = O.fsw_embed_csr_backward(Xf, rowptr, col, W, g["param_projVecs"], g["param_freqs"], g["gout"].reshape(-1, g["gout"].shape[-1]))
    np.testing.assert_allclose(r["dW"].reshape(g["dW_f64"].shape), g["dW_f64"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["dX"].reshape(g["dX_f64"].shape), g["dX_f64"], **TOL)


def test_single_point_known_answer():
    g = load_golden("emb_dense_n1")
    X = g["X"]
    for b in range(X.shape[0]):
        ka = O.single_point_known_answer(X[b, 0], g["param_projVecs"], g["param_freqs"]) + g["param_bias"]
        np.testing.assert_allclose(ka, g["out_f64"][b], **TOL)


GRAPH_CASES = {
    "emb_graph_unit": {},
    "emb_graph_weighted": dict(encode_total_mass=True, total_mass_encoding_function="log"),
    "emb_graph_homog": dict(encode_total_mass=True, total_mass_encoding_method="homog"),
    "emb_graph_homog_alt": dict(encode_total_mass=True, total_mass_encoding_method="homog_alt", total_mass_encoding_function="sqrt"),
}


@pytest.mark.parametrize("name", list(GRAPH_CASES))
def test_graph_forward(name):
    g = load_golden(name)
    S, N = g["A_shape"]
    rowptr, col, W = coo_to_csr(g["A_indices"], g["A_values"], int(S))
    out = O.fsw_embedding_forward(g["X"], rowptr, col, W, emb_params(g), GRAPH_CASES[name])
    np.testing.assert_allclose(out, g["out_f64"], **TOL)


def test_graph_backward_unit():
    g = load_golden("emb_graph_unit")
    S, N = g["A_shape"]
    rowptr, col, W = coo_to_csr(g["A_indices"], g["A_values"], int(S))
    r = O.fsw_embed_csr_backward(g["X"], rowptr, col, W, g["param_projVecs"], g["param_freqs"], g["gout"])
    np.testing.assert_allclose(r["dX"], g["dX_f64"], **TOL)
    np.testing.assert_allclose(r["dtheta"], g["dprojVecs_f64"], **TOL)
    np.testing.assert_allclose(r["dxi"], g["dfreqs_f64"], rtol=1e-8, atol=1e-9)


def conv_params(g, prefix="param_"):
    p = dict(projVecs=g[prefix + "fsw_embed.projVecs"], freqs=g[prefix + "fsw_embed.freqs"],
             bias=g.get(prefix + "fsw_embed.bias"),
             total_mass_encoding_scale=g.get(prefix + "fsw_embed.total_mass_encoding_scale"))
    mlp = []
    i = 0
    while True:
        # Sequential indices of Linear layers: 0, 2, 4 ... (Linear, LeakyReLU pairs)
        key = prefix + "mlp.%d.weight" % (2 * i)
        if key not in g:
            break
        mlp.append((g[key], g.get(prefix + "mlp.%d.bias" % (2 * i))))
        i += 1
    # batchNorm_final: the Sequential is Linear, act, ..., Linear, BatchNorm1d, act (fsw_conv.py:298-304)
    bn_key = prefix + "mlp.%d.running_mean" % (2 * i - 1)
    if i > 0 and bn_key in g:
        b = prefix + "mlp.%d." % (2 * i - 1)
        p["bn_final"] = dict(mean=g[b + "running_mean"], var=g[b + "running_var"], weight=g[b + "weight"], bias=g[b + "bias"])
    p["mlp"] = mlp
    p["dim_reduct"] = g.get(prefix + "dim_reduct")
    return p


CONV_CASES = {
    "conv_default": dict(encode_total_mass=True),
    "conv_selfloop_gcn": dict(encode_total_mass=True, self_loop_weight=0.2, edge_weighting="gcn",
                              total_mass_encoding_function="log"),
    "conv_edgefeat": dict(encode_total_mass=True),
    "conv_homog_nomlp": dict(encode_total_mass=True, total_mass_encoding_method="homog"),
    "conv_wide": dict(encode_total_mass=True),
}


@pytest.mark.parametrize("name", list(CONV_CASES))
def test_conv_forward(name):
    g = load_golden(name)
    out = O.fsw_conv_forward(g["x"], g["edge_index"], conv_params(g), CONV_CASES[name],
                             edge_features=g.get("edge_features"))
    np.testing.assert_allclose(out, g["out_f64"], **TOL)


ACCEPT_CASES = {
    # test_conv.py:10-48
    "conv_acceptance_testconv": (dict(encode_total_mass=True, total_mass_encoding_method="homog", total_mass_encoding_function="log",
                                      self_loop_weight=0.2), "f64"),
    # demo_conv.py:10-38 (fp32 in the demo: the oracle is compared with the reference in fp64 at the fp32-rounded point)
    "conv_acceptance_democonv": (dict(encode_total_mass=True), "r64"),
}


@pytest.mark.parametrize("name", list(ACCEPT_CASES))
def test_conv_acceptance_forward(name):
    """the reference's own acceptance shapes (ER graph, 100 vertices, edge dim 11, 3 MLP layers)"""
    g = load_golden(name)
    cfg, tag = ACCEPT_CASES[name]
    params = conv_params(g)
    x, ef = g["x"], g["edge_features"]
    if tag == "r64":   # inputs and parameters rounded to fp32, arithmetic in fp64
        x, ef = x.astype(np.float32).astype(np.float64), ef.astype(np.float32).astype(np.float64)
    out = O.fsw_conv_forward(x, g["edge_index"], params, cfg, edge_features=ef)
    np.testing.assert_allclose(out, g["out_" + tag], **TOL)


def test_demo_embedding_forward_and_weight_gradient():
    """demo_fsw_embedding.py:10-25 (X [3,2,5,100,20], softmax weights, d_out = 1000): forward, dX and dW of the oracle
    against the reference's fp64 arithmetic at the fp32-rounded point.  Softmax weights sum to 1 up to an ulp, so the batch
    mixes padded rows (T < 1) and rows above the threshold: both branches of the normalisation chain occur."""
    g = load_golden("emb_acceptance_demo")
    X = g["X"].astype(np.float64)
    W = g["W"].astype(np.float64)
    bd = X.shape[:-2]
    n, d = X.shape[-2:]
    B = int(np.prod(bd))
    rowptr, col = O.dense_to_csr(B, n)
    theta, xi = g["param_projVecs"].astype(np.float64), g["param_freqs"].astype(np.float64)
    Kc = 40   # a column subset keeps the pure-python oracle fast; outputs are independent per slice
    out = O.fsw_embed_csr(X.reshape(B * n, d), rowptr, col, W.reshape(-1), theta[:Kc], xi[:Kc]) + g["param_bias"][:Kc].astype(np.float64)
    np.testing.assert_allclose(out, g["out_r64"].reshape(B, -1)[:, :Kc], rtol=1e-9, atol=1e-10)
    T = W.reshape(B, n).sum(axis=1)
    assert (T < 1).any() and (T > 1).any()
    # gradients need every slice: dX, dW of sum(gout * out) against the reference's autograd
    r = O.fsw_embed_csr_backward(X.reshape(B * n, d), rowptr, col, W.reshape(-1), theta, xi, g["gout"].reshape(B, -1).astype(np.float64))
    np.testing.assert_allclose(r["dX"].reshape(g["dX_r64"].shape), g["dX_r64"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(r["dW"].reshape(g["dW_r64"].shape), g["dW_r64"], rtol=1e-7, atol=1e-7)


def test_readout_forward():
    g = load_golden("readout_default")
    gi = g["graph_index"]
    B = g["out_f64"].shape[0]
    rowptr = np.zeros(B + 1, dtype=np.int64)
    np.add.at(rowptr, gi + 1, 1)
    rowptr = np.cumsum(rowptr)
    p = conv_params(g)
    emb = O.fsw_embedding_forward(g["x"], rowptr, None, None, p, dict(encode_total_mass=True))
    h = emb
    for (wt, b) in p["mlp"]:
        h = O.leaky_relu(h @ wt.T + (b if b is not None else 0.0))
    np.testing.assert_allclose(h, g["out_f64"], **TOL)


@pytest.mark.parametrize("tag,tol", [("f64", 1e-12), ("f32", 0.0)])
def test_segcumsum(tag, tol):
    g = load_golden("segcumsum")
    out = O.segcumsum(g["values_" + tag], g["segment_ids"])
    # the oracle follows segcumsum_slow's left-to-right association exactly => bit-identical
    if tol == 0.0:
        assert np.array_equal(out, g["out_slow_" + tag])
    else:
        np.testing.assert_allclose(out, g["out_slow_" + tag], rtol=tol, atol=tol)


@pytest.mark.parametrize("name", ["emb_cartesian", "emb_cartesian_collapse"])
def test_cartesian(name):
    g = load_golden(name)
    X = g["X"]
    B, n, d = X.shape
    rowptr, col = O.dense_to_csr(B, n)
    core = O.fsw_embed_csr(X.reshape(B * n, d), rowptr, col, g["W"].reshape(-1), g["param_projVecs"],
                           g["param_freqs"], cartesian=True)
    out = core + g["param_bias"].reshape((1,) + core.shape[1:]) if g["param_bias"].ndim == 2 else \
        core.reshape(B, -1) + g["param_bias"]
    np.testing.assert_allclose(out.reshape(g["out_f64"].shape), g["out_f64"], **TOL)


# ------------------------------------------------------------------------------------------------
# the C restatement (oracle/fsw_oracle.c, used as the CPU baseline of bench.py) against the same vectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["emb_dense_weighted", "emb_dense_unit", "emb_dense_big"])
def test_c_oracle_dense(name):
    from oracle import c_oracle as C
    g = load_golden(name)
    Xf, rowptr, col, W, bd, n, d = dense_inputs(g)
    gout = g["gout"].reshape(-1, g["gout"].shape[-1])
    out, mass, dX, dtheta, dxi = C.embed_forward_backward(Xf, rowptr, col, W, g["param_projVecs"], g["param_freqs"], gout)
    np.testing.assert_allclose((out + g["param_bias"]).reshape(g["out_f64"].shape), g["out_f64"], **TOL)
    np.testing.assert_allclose(dX.reshape(g["dX_f64"].shape), g["dX_f64"], **TOL)
    np.testing.assert_allclose(dtheta, g["dprojVecs_f64"], **TOL)
    np.testing.assert_allclose(dxi, g["dfreqs_f64"], rtol=1e-8, atol=1e-9)


def test_c_oracle_graph():
    from oracle import c_oracle as C
    g = load_golden("emb_graph_unit")
    S, N = g["A_shape"]
    rowptr, col, W = coo_to_csr(g["A_indices"], g["A_values"], int(S))
    out, mass, dX, dtheta, dxi = C.embed_forward_backward(g["X"], rowptr, col, W, g["param_projVecs"], g["param_freqs"], g["gout"])
    np.testing.assert_allclose(out + g["param_bias"], g["out_f64"], **TOL)
    np.testing.assert_allclose(dX, g["dX_f64"], **TOL)
    np.testing.assert_allclose(dtheta, g["dprojVecs_f64"], **TOL)
    np.testing.assert_allclose(dxi, g["dfreqs_f64"], rtol=1e-8, atol=1e-9)
    # float build: same algorithm in fp32, close to the reference's own fp32 output
    out32, _ = C.embed_forward_backward(g["X"], rowptr, col, W, g["param_projVecs"], g["param_freqs"], dtype=np.float32)
    np.testing.assert_allclose(out32 + g["param_bias"], g["out_f64"], rtol=1e-3, atol=1e-3)
    assert C.threads() >= 1
