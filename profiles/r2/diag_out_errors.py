"""Where do the forward values deviate most from the fp64 oracle?  (diagnostic, GPU)  Prints the worst (segment, slice) pairs
of the configs[3]-shaped test graph with their segment size and frequency."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from fsw_gnn_b200 import FSW_conv, ops, synthetic as syn
from fsw_gnn_b200.ops import SegmentPlan
from oracle import c_oracle as C

dev = torch.device("cuda:0")
torch.manual_seed(3)
N = 50_000
deg = syn.products_like_degrees(N, int(N * 25.8), seed=5, device=dev).cpu().numpy()
deg[:6] = [17000, 9000, 4097, 2049, 1025, 513]
conv = FSW_conv(100, 100, device=dev)
emb = conv.fsw_embed
rng = np.random.default_rng(40)
S = len(deg)
rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
E = int(rowptr[-1])
col = rng.integers(0, N, E).astype(np.int32)
X = rng.standard_normal((N, 100)).astype(np.float32)
plan = SegmentPlan(S, E, torch.as_tensor(rowptr.astype(np.int32), device=dev), 0, torch.as_tensor(col, device=dev), None, 1.0, torch.float32, dev)
Xt = torch.as_tensor(X, device=dev)
theta = emb.projVecs.detach().cpu().numpy().astype(np.float64)
xi = emb.freqs.detach().cpu().numpy().astype(np.float64)
ref, mass = C.embed_forward_backward(X.astype(np.float64), rowptr, col, None, theta, xi)
K = theta.shape[0]
with torch.no_grad():
    Xp = ops.project(Xt, emb.projVecs.detach()[:, :100], ops.round_up(K, 8))[:, :K]
kref, _ = C.embed_forward_backward(Xp.cpu().numpy().astype(np.float64), rowptr, col, None, np.eye(K), xi)
for mode in ("train", "eval"):
    if mode == "train":
        out = emb.embed_plan(Xt.clone().requires_grad_(True), plan).detach()
    else:
        with torch.no_grad():
            out = emb.embed_plan(Xt, plan)
    core = out[:, 1:].cpu().numpy().astype(np.float64)
    for nm, r in (("fp64-projection oracle", ref), ("keys oracle", kref)):
        err = np.abs(core - r)
        lim = 1e-6 + 1e-5 * np.abs(r)
        bad = np.argwhere(err > lim)
        print("== %s, %s: %d strict failures, max err %.3g" % (mode, nm, len(bad), err.max()))
        if len(bad):
            ds = deg[bad[:, 0]]
            print("   by segment size: ", {int(b): int((ds >= a) & (ds < b)).sum() if False else int(((ds >= a) & (ds < b)).sum())
                                           for a, b in [(0, 33), (33, 129), (129, 513), (513, 1025), (1025, 4097), (4097, 1 << 30)]})
            ks = bad[:, 1]
            print("   by slice index quartile:", np.histogram(ks, bins=[0, 50, 100, 150, 200])[0])
            order = np.argsort(-(err / lim)[bad[:, 0], bad[:, 1]])[:10]
            for i in order:
                s_, k_ = bad[i]
                print("   seg %d n=%d slice %d xi=%.2f ref=%.6g got=%.6g err=%.3g" % (s_, deg[s_], k_, xi[k_], r[s_, k_], core[s_, k_], err[s_, k_]))
    # error statistics per size class vs keys oracle
    err = np.abs(core - kref)
    for a, b in [(1, 33), (33, 129), (129, 513), (513, 1025), (1025, 4097), (4097, 1 << 30)]:
        m = (deg >= a) & (deg < b)
        if m.any():
            print("   n in [%d,%d): %d segs, rms err %.3g, max err %.3g" % (a, b, m.sum(), np.sqrt((err[m] ** 2).mean()), err[m].max()))
