"""World-size-2 gloo tests (CPU) of the host-side logic of the destination-sharded FSW_conv:
edge-balanced partition, source re-indexing into the all-gathered layout, the autograd all-gather /
reduce-scatter pair and the gradient all-reduce.  The embedding itself is computed by the CPU oracle here
(the CUDA kernels need a GPU); what is checked is that shards + exchange reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fsw_gnn_b200 import dist as fdist
from fsw_gnn_b200 import synthetic as syn
from oracle import fsw_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _csr(ei_src_rows, dst_local, n_rows):
    order = np.argsort(dst_local, kind="stable")
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(dst_local, minlength=n_rows))])
    return rowptr, ei_src_rows[order]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        N, E, d, K = 300, 4000, 4, 6
        deg = syn.lognormal_degrees(N, E / N, 1.0, 1, 200, 0, dev)
        ranges = syn.balanced_row_ranges(deg, world)
        lo, hi = ranges[rank]
        ei = syn.edges_for_rows(deg, lo, hi, N, seed=0, device=dev)
        # (1) the shards are the rows of the same global graph
        full = syn.edges_for_rows(deg, 0, N, N, seed=0, device=dev, shuffle=False)
        mine = syn.edges_for_rows(deg, lo, hi, N, seed=0, device=dev, shuffle=False)
        sel = (full[1] >= lo) & (full[1] < hi)
        assert torch.equal(full[:, sel], mine)
        counts = [int(deg[a:b].sum()) for a, b in ranges]
        assert max(counts) - min(counts) <= int(deg.max())  # edge balanced up to one row
        # (2) sources re-indexed into the padded gathered layout
        max_rows = max(b - a for a, b in ranges)
        col = fdist.remap_sources(ei[0], ranges, max_rows)
        gen = torch.Generator().manual_seed(5)
        X = torch.randn(N, d, generator=gen, dtype=torch.float64)
        x_local = X[lo:hi].clone().requires_grad_(True)
        # (3) autograd all-gather: forward layout and backward = reduce-scatter
        x_all = fdist.all_gather_rows(x_local, max_rows)
        assert x_all.shape[0] == world * max_rows
        assert torch.equal(x_all[col], X[ei[0]])
        wgt = torch.arange(1, world * max_rows * d + 1, dtype=torch.float64).reshape(world * max_rows, d) * (rank + 1)
        (x_all * wgt).sum().backward()
        tot = sum(torch.arange(1, world * max_rows * d + 1, dtype=torch.float64).reshape(world * max_rows, d) * (r + 1) for r in range(world))
        assert torch.allclose(x_local.grad, tot[rank * max_rows: rank * max_rows + (hi - lo)])
        # (3b) the non-autograd exchange used around the fused kernels (projected rows): gather layout and adjoint
        ex = fdist.RowExchange(max_rows)
        xp_all = ex.gather(X[lo:hi])
        assert xp_all.shape[0] == world * max_rows and torch.equal(xp_all[col], X[ei[0]])
        back = ex.scatter(wgt.clone(), hi - lo)
        assert back.shape[0] == hi - lo and torch.allclose(back, tot[rank * max_rows: rank * max_rows + (hi - lo)])
        # (4) sharded embedding (oracle as the compute) == rows lo..hi of the single-process embedding
        theta = np.random.default_rng(1).standard_normal((K, d))
        theta /= np.linalg.norm(theta, axis=1, keepdims=True)
        xi = np.linspace(0.1, 5.0, K)
        rp_l, col_l = _csr(col.numpy(), (ei[1] - lo).numpy(), hi - lo)
        emb_l = O.fsw_embed_csr(x_all.detach().numpy(), rp_l, col_l, None, theta, xi)
        rp_f, col_f = _csr(full[0].numpy(), full[1].numpy(), N)
        emb_f = O.fsw_embed_csr(X.numpy(), rp_f, col_f, None, theta, xi)
        np.testing.assert_allclose(emb_l, emb_f[lo:hi], rtol=1e-12, atol=1e-12)
        # (5) flat gradient all-reduce
        lin = torch.nn.Linear(3, 2)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        fdist.all_reduce_gradients([lin])
        for p in lin.parameters():
            assert torch.all(p.grad == sum(range(1, world + 1)))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_sharded_conv_host_logic_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_balanced_ranges_cover_everything():
    deg = torch.tensor([5, 1, 1, 1, 10, 2, 2, 2, 2, 30, 1, 1])
    for parts in (1, 2, 3, 4):
        r = syn.balanced_row_ranges(deg, parts)
        assert r[0][0] == 0 and r[-1][1] == deg.numel()
        assert all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
