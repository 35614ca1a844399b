"""CPU ORACLE (test infrastructure, NOT product code) for the Fourier Sliced-Wasserstein hot path.

This file is a plain numpy/fp64 restatement of the algorithm of tal-amir/fsw-gnn
(`/root/reference`).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and only as the checker.  The product path
(`fsw_gnn_b200/`) never imports anything from `oracle/`.

Parity status: PINNED.  The reference holds no golden vectors of its own, so the
oracle is pinned against outputs of the unmodified reference executed in this image
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).

Every function cites the reference file:line it follows (paths relative to /root/reference).

Data model (replaces the reference's coalesced sparse-COO tensors, fsw_embedding.py:2269-2278):
a batch of weighted multisets is a CSR structure
    rowptr[S+1], col[E] (index of the element's point in X; None = identity), W[E] (raw weights)
and `X[N, d]` holds the points; segment s is the multiset {(X[col[e]], W[e]) : rowptr[s] <= e < rowptr[s+1]}.
"""
import math

import numpy as np

# --------------------------------------------------------------------------------------------
# Segmented cumulative sum  (fsw_embedding.py:3016-3027 `segcumsum_slow`, :2866-2874)
# --------------------------------------------------------------------------------------------


def segcumsum(values, segment_ids):
    """Inclusive cumulative sum restarting whenever segment_ids changes between neighbours.

    Follows `segcumsum_slow` (fsw_embedding.py:3016-3027): a segment is a maximal run of equal
    consecutive ids.  Same dtype in and out (sums are carried in the value dtype like the reference).
    """
    values = np.asarray(values)
    segment_ids = np.asarray(segment_ids)
    out = np.empty_like(values)
    acc = values.dtype.type(0)
    for i in range(values.shape[0]):
        if i > 0 and segment_ids[i] == segment_ids[i - 1]:
            acc = values.dtype.type(acc + values[i])
        else:
            acc = values[i]
        out[i] = acc
    return out


# --------------------------------------------------------------------------------------------
# Core per-multiset evaluation
# --------------------------------------------------------------------------------------------


def _sinc(x):
    # torch.sinc / np.sinc: sin(pi x)/(pi x), 1 at 0 (fsw_embedding.py:1001-1002)
    return np.sinc(x)


def embed_sorted(ps, ws, xi, form="prod"):
    """Embedding of ONE multiset along ONE slice given projections sorted ascending.

    ps[n]: sorted projected values, ws[n]: normalised weights in the same order (sum == 1),
    xi: frequency.  Returns (1+xi) * sum_j ps[j] * D_j   (fsw_embedding.py:1101, :1109) with
      form='diff': D = diff_zeropad(2*C*sinc(2*xi*C))            (dense path, :999-1003, :1173-1177)
      form='prod': D = 2*w*sinc(xi*w)*cos(pi*xi*(2C - w))        (sparse path, :1047-1075)
    where C = cumsum(ws) (:999, :1032).
    """
    ps = np.asarray(ps, dtype=np.float64)
    ws = np.asarray(ws, dtype=np.float64)
    C = np.cumsum(ws)
    if form == "diff":
        s = 2.0 * C * _sinc(2.0 * xi * C)
        D = np.diff(s, prepend=0.0)
    else:
        D = 2.0 * ws * _sinc(xi * ws) * np.cos(np.pi * xi * (2.0 * C - ws))
    return (1.0 + xi) * float(np.dot(ps, D))


def mass_pad_normalise(W_seg, thresh):
    """Total mass, deficit padding and normalisation of one multiset's weights.

    fsw_embedding.py:778-829: T = sum W; if T < thresh a point at x = 0 with weight (thresh - T) is
    appended (:787-815); weights are divided by max(T, thresh) (:817-829).
    Returns (T, w_norm (with the pad weight appended when padded), padded: bool).
    NB the reference pads *every* row with a (possibly zero) weight once any row is deficient
    (:790-807); a zero-weight element changes nothing, so only deficient rows are padded here.
    """
    W_seg = np.asarray(W_seg, dtype=np.float64)
    T = float(W_seg.sum())
    if T < thresh:
        w = np.concatenate([W_seg, [thresh - T]]) / thresh
        return T, w, True
    return T, W_seg / T, False


def fsw_embed_csr(X, rowptr, col, W, theta, xi, thresh=1.0, E_feat=None, form="prod", cartesian=False):
    """[S, K] (or [S, K, F] when cartesian) embedding of S multisets before the epilogue.

    Follows FSW_embedding.forward part D + forward_helper (fsw_embedding.py:778-851, :894-1112):
    projection Xp = X . theta[:, :d]^T (:911/:913), plus per-element edge term E_feat . theta[:, d:]^T
    (:934-947), ascending sort per (multiset, slice) (:923-925, :954), weights carried along
    (:989-990, :1017-1025), cumulative weights, Fourier quantile evaluation, reduction, (1+xi).
    """
    X = np.asarray(X, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    xi = np.asarray(xi, dtype=np.float64)
    d = X.shape[1]
    S = len(rowptr) - 1
    K = theta.shape[0]
    Xp = X @ theta[:, :d].T  # [N, K]
    Ep = None
    if E_feat is not None:
        E_feat = np.asarray(E_feat, dtype=np.float64)
        if E_feat.ndim == 1:
            E_feat = E_feat[:, None]
        Ep = E_feat @ theta[:, d:].T  # [E, K]
    F = xi.shape[0]
    out = np.zeros((S, K, F) if cartesian else (S, K))
    for s in range(S):
        lo, hi = int(rowptr[s]), int(rowptr[s + 1])
        idx = np.arange(lo, hi) if col is None else np.asarray(col[lo:hi])
        Wseg = np.ones(hi - lo) if W is None else np.asarray(W[lo:hi], dtype=np.float64)
        T, w, padded = mass_pad_normalise(Wseg, thresh)
        P = Xp[idx, :]
        if Ep is not None:
            P = P + Ep[lo:hi, :]
        if padded:
            P = np.concatenate([P, np.zeros((1, K))], axis=0)
        for k in range(K):
            order = np.argsort(P[:, k], kind="stable")
            ps, ws = P[order, k], w[order]
            if cartesian:
                for f in range(F):
                    out[s, k, f] = embed_sorted(ps, ws, xi[f], form)
            else:
                out[s, k] = embed_sorted(ps, ws, xi[k], form)
    return out


def total_mass_function(T, name):
    """fsw_embedding.py:857-865."""
    T = np.asarray(T, dtype=np.float64)
    if name == "identity":
        return T.copy()
    if name == "sqrt":
        return 2.0 * (T / (np.sqrt(T + 1.0) + 1.0))
    if name == "log":
        return np.log1p(T)
    raise ValueError(name)


def epilogue(core, T, *, encode_total_mass, tm_function="identity", tm_scale=1.0, tm_method="plain", bias=None):
    """Total-mass channel (PREPENDED as element 0) and bias, fsw_embedding.py:856-888, :1137-1144."""
    out = core
    if encode_total_mass:
        tm = total_mass_function(T, tm_function)[:, None] * tm_scale
        if tm_method == "plain":
            out = np.concatenate([tm, out], axis=-1)
        elif tm_method == "homog":
            nrm = np.mean(np.abs(out), axis=-1, keepdims=True)
            out = np.concatenate([tm * nrm, out], axis=-1)
        elif tm_method == "homog_alt":
            nrm = np.mean(np.abs(out), axis=-1, keepdims=True)
            p1 = np.where(tm <= 1, tm * (2 - tm), 1.0)
            p2 = np.where(tm <= 1, tm**2, 2 * tm - 1)
            out = np.concatenate([p1 * nrm, p2 * out], axis=-1)
        else:
            raise ValueError(tm_method)
    if bias is not None:
        out = out + np.asarray(bias, dtype=np.float64)
    return out


def segment_mass(rowptr, W):
    S = len(rowptr) - 1
    T = np.zeros(S)
    for s in range(S):
        lo, hi = int(rowptr[s]), int(rowptr[s + 1])
        T[s] = (hi - lo) if W is None else float(np.sum(np.asarray(W[lo:hi], dtype=np.float64)))
    return T


def fsw_embedding_forward(X, rowptr, col, W, params, cfg, E_feat=None, form="prod"):
    """Whole FSW_embedding.forward on CSR input (fsw_embedding.py:587-890).

    params: dict(projVecs [K, d+d_edge], freqs [K], bias or None, total_mass_encoding_scale or None)
    cfg: dict(total_mass_pad_thresh, encode_total_mass, total_mass_encoding_function,
              total_mass_encoding_method)
    Returns [S, d_out].
    """
    core = fsw_embed_csr(X, rowptr, col, W, params["projVecs"], params["freqs"],
                         thresh=cfg.get("total_mass_pad_thresh", 1.0), E_feat=E_feat, form=form)
    T = segment_mass(rowptr, W)
    return epilogue(core, T,
                    encode_total_mass=cfg.get("encode_total_mass", False),
                    tm_function=cfg.get("total_mass_encoding_function", "identity"),
                    tm_scale=(params.get("total_mass_encoding_scale") if params.get("total_mass_encoding_scale") is not None else 1.0),
                    tm_method=cfg.get("total_mass_encoding_method", "plain"),
                    bias=params.get("bias"))


# --------------------------------------------------------------------------------------------
# Closed-form backward of the core (mirrors ag.*.backward, fsw_embedding.py:1286-2258,
# and torch autograd of the dense path :911-1109)
# --------------------------------------------------------------------------------------------


def fsw_embed_csr_backward(X, rowptr, col, W, theta, xi, g, thresh=1.0, E_feat=None):
    """Gradients of sum(g * fsw_embed_csr(...)) w.r.t. X, theta, xi, W (raw weights), E_feat.

    g: [S, K] upstream gradient.  Returns dict(dX, dtheta, dxi, dW, dE).
    """
    X = np.asarray(X, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    xi = np.asarray(xi, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    N, d = X.shape
    S = len(rowptr) - 1
    K = theta.shape[0]
    Xp = X @ theta[:, :d].T
    Ep = None
    if E_feat is not None:
        E_feat = np.asarray(E_feat, dtype=np.float64)
        if E_feat.ndim == 1:
            E_feat = E_feat[:, None]
        Ep = E_feat @ theta[:, d:].T
    nE = int(rowptr[-1])
    dXp = np.zeros((N, K))
    dEp = np.zeros((nE, K))
    dxi = np.zeros(K)
    dW = np.zeros(nE)
    # the reference pads EVERY row once any row is deficient (fsw_embedding.py:790-807); the extra element has weight
    # max(thresh - T, 0).  It changes no value, but for a row with T == thresh EXACTLY its low clamp is still 'active'
    # (custom_lowclamp, :1735-1744: input >= thresh with input = thresh - T = 0), so the weight gradient of such a row
    # sees the pad element
    any_deficient = any((float(np.sum(np.asarray(W[int(rowptr[s]):int(rowptr[s + 1])], dtype=np.float64))) if W is not None
                         else float(rowptr[s + 1] - rowptr[s])) < thresh for s in range(S))
    for s in range(S):
        lo, hi = int(rowptr[s]), int(rowptr[s + 1])
        n = hi - lo
        idx = np.arange(lo, hi) if col is None else np.asarray(col[lo:hi])
        Wseg = np.ones(n) if W is None else np.asarray(W[lo:hi], dtype=np.float64)
        T, w, padded = mass_pad_normalise(Wseg, thresh)
        zero_pad = (not padded) and any_deficient and T == thresh and n > 0
        if zero_pad:
            w = np.concatenate([w, [0.0]])
        Sp = max(T, thresh)
        P = Xp[idx, :]
        if Ep is not None:
            P = P + Ep[lo:hi, :]
        if padded or zero_pad:
            P = np.concatenate([P, np.zeros((1, K))], axis=0)
        dw_total = np.zeros(P.shape[0])  # dL/dw (normalised weights, unsorted order)
        for k in range(K):
            x = xi[k]
            G = g[s, k] * (1.0 + x)
            order = np.argsort(P[:, k], kind="stable")
            ps, ws = P[order, k], w[order]
            C = np.cumsum(ws)
            C0 = C - ws
            if x != 0.0:
                D = (np.sin(2 * np.pi * x * C) - np.sin(2 * np.pi * x * C0)) / (np.pi * x)
                dD = (2 * C * np.cos(2 * np.pi * x * C) - 2 * C0 * np.cos(2 * np.pi * x * C0)) / x - D / x
            else:
                D = 2.0 * ws
                dD = np.zeros_like(ws)
            # projections
            dps = G * D
            dP = np.zeros_like(dps)
            dP[order] = dps
            m = n  # real elements; the pad point (if any) has a constant projection 0
            np.add.at(dXp[:, k], idx, dP[:m])
            if Ep is not None:
                dEp[lo:hi, k] += dP[:m]
            # frequency
            dxi[k] += g[s, k] * (float(np.dot(ps, D)) + (1.0 + x) * float(np.dot(ps, dD)))
            # cumulative weights -> weights (reverse cumulative sum)
            pnext = np.concatenate([ps[1:], [0.0]])
            dC = G * 2.0 * np.cos(2 * np.pi * x * C) * (ps - pnext)
            dws = np.cumsum(dC[::-1])[::-1]
            tmp = np.zeros_like(dws)
            tmp[order] = dws
            dw_total += tmp
        # normalisation chain (custom_lowclamp passes the gradient iff input >= thresh, :1738-1743)
        if padded:
            # w_i = W_i/thresh, w_pad = (thresh - T)/thresh
            dW[lo:hi] = (dw_total[:n] - dw_total[n]) / thresh
        elif zero_pad:
            # w_i = W_i/T, w_pad = (thresh - T)/T with both clamps active at T == thresh
            dW[lo:hi] = dw_total[:n] / Sp - float(np.dot(dw_total[:n], Wseg)) / (Sp * Sp) - dw_total[n] / Sp
        else:
            dW[lo:hi] = dw_total[:n] / Sp - float(np.dot(dw_total[:n], Wseg)) / (Sp * Sp)
    dX = dXp @ theta[:, :d]
    dtheta = np.zeros_like(theta)
    dtheta[:, :d] = dXp.T @ X
    dE = None
    if Ep is not None:
        dE = dEp @ theta[:, d:]
        dtheta[:, d:] = dEp.T @ E_feat
    return dict(dX=dX, dtheta=dtheta, dxi=dxi, dW=dW, dE=dE, dXp=dXp)


# --------------------------------------------------------------------------------------------
# FSW_conv (fsw_conv.py:331-447)
# --------------------------------------------------------------------------------------------


def edge_index_to_csr(edge_index, num_vertices, self_loop_weight=0.0, edge_weighting="unit", edge_features=None):
    """`FSW_conv.edge_index_to_adj` (fsw_conv.py:384-447) restated on CSR.

    rows = destinations, cols = sources (`edge_index.flip(0)`, :387); optional self loops (:390-395);
    duplicate (dst, src) pairs are SUMMED by `coalesce` (:397-398) - weights and edge features alike
    (:438-439); 'gcn' weighting divides by sqrt(deg_dst) and sqrt(deg_src) with deg = weighted
    in-degree after coalescing (:400-409).  Returns rowptr, col, W, in_degrees, E_feat (or None).
    """
    src = np.asarray(edge_index[0], dtype=np.int64)
    dst = np.asarray(edge_index[1], dtype=np.int64)
    vals = np.ones(src.shape[0])
    ef = None
    if edge_features is not None:
        ef = np.asarray(edge_features, dtype=np.float64)
        if ef.ndim == 1:
            ef = ef[:, None]
    if self_loop_weight > 0:
        loops = np.arange(num_vertices, dtype=np.int64)
        src = np.concatenate([src, loops])
        dst = np.concatenate([dst, loops])
        vals = np.concatenate([vals, np.full(num_vertices, float(self_loop_weight))])
        if ef is not None:
            ef = np.concatenate([ef, np.zeros((num_vertices, ef.shape[1]))], axis=0)
    key = dst * num_vertices + src
    order = np.argsort(key, kind="stable")
    key_s = key[order]
    uniq, start = np.unique(key_s, return_index=True)
    W = np.add.reduceat(vals[order], start) if len(start) else np.zeros(0)
    E_feat = None
    if ef is not None:
        E_feat = np.add.reduceat(ef[order], start, axis=0) if len(start) else np.zeros((0, ef.shape[1]))
    rows = uniq // num_vertices
    col = uniq % num_vertices
    rowptr = np.zeros(num_vertices + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    deg = np.zeros(num_vertices)
    np.add.at(deg, rows, W)
    if edge_weighting == "gcn":
        W = W / np.sqrt(deg[rows]) / np.sqrt(deg[col])
    elif edge_weighting != "unit":
        raise ValueError(edge_weighting)
    return rowptr, col, W, deg, E_feat


def leaky_relu(x, slope=0.2):
    return np.where(x >= 0, x, slope * x)


def fsw_conv_forward(x, edge_index, params, cfg, edge_features=None):
    """`FSW_conv.forward` (fsw_conv.py:331-371) with the default LeakyReLU(0.2) activations, no
    batch-norm / dropout.  params: embedding params + 'mlp' = [(weight [out,in], bias or None), ...]
    or 'dim_reduct'.  cfg: embedding cfg + self_loop_weight, edge_weighting, concat_self,
    message_weight_vs_self."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    rowptr, col, W, deg, E_feat = edge_index_to_csr(edge_index, n, cfg.get("self_loop_weight", 0.0),
                                                     cfg.get("edge_weighting", "unit"), edge_features)
    emb = fsw_embedding_forward(x, rowptr, col, W, params, cfg, E_feat=E_feat)
    if cfg.get("concat_self", True):
        emb = np.concatenate([cfg.get("message_weight_vs_self", 1.0) * emb, x], axis=-1)
    if params.get("mlp"):
        h = emb
        nl = len(params["mlp"])
        for li, (wt, b) in enumerate(params["mlp"]):
            h = h @ np.asarray(wt, dtype=np.float64).T
            if b is not None:
                h = h + np.asarray(b, dtype=np.float64)
            bn = params.get("bn_final") if li == nl - 1 else None
            if bn is not None:
                # BatchNorm1d in eval mode after the last Linear (fsw_conv.py:300-301): running statistics
                h = (h - bn["mean"]) / np.sqrt(bn["var"] + bn.get("eps", 1e-5)) * bn["weight"] + bn["bias"]
            h = leaky_relu(h)
        return h
    if params.get("dim_reduct") is not None:
        return emb @ np.asarray(params["dim_reduct"], dtype=np.float64).T
    return emb


# --------------------------------------------------------------------------------------------
# Convenience: dense inputs -> CSR
# --------------------------------------------------------------------------------------------


def dense_to_csr(batch, n):
    """`batch` multisets of `n` points each, points stored contiguously: identity columns."""
    rowptr = np.arange(batch + 1, dtype=np.int64) * n
    return rowptr, None


def single_point_known_answer(x, theta, xi):
    """Known answer: one point x with weight 1 => out_k = (1+xi_k) <x,theta_k> 2 sinc(2 xi_k)."""
    x = np.asarray(x, dtype=np.float64)
    return (1.0 + xi) * (theta @ x) * 2.0 * np.sinc(2.0 * xi)
