"""Multi-GPU FSW_conv: destination vertices sharded across the GPUs of one box.

Recipient rows of the adjacency are independent, so each rank owns a contiguous range of destination
vertices chosen to hold an equal share of the EDGES, keeps the CSR rows / plan of that range only, and
produces the output rows of that range.  Each rank projects its own vertices onto the slices (K1), and the only
exchange per layer is an all-gather of those PROJECTED rows (NCCL over NVLink / NVSwitch); its adjoint in the
backward pass is a reduce-scatter of the projected gradient, after which the contractions dX, dtheta are local;
parameter gradients are all-reduced once per step.  Point-cloud batches need no
exchange at all (shard the batch dimension; see bench.py).

One process per GPU; torch.distributed provides the communicator (backend nccl on GPUs, gloo in the
CPU tests of the host-side logic).
"""
import os

import torch
import torch.distributed as dist

from . import graph as _graph


# Optional timing of the exchange as the compute stream sees it (bench.py `nccl_exposed_ms_per_step`): a CUDA event before a
# collective is issued and one after the compute stream has waited for it.  Off by default (no events, no overhead).
EXCHANGE_TIMING = {"on": False, "events": []}


class _TimedWork:
    """async work handle whose wait() also records the end event on the waiting stream"""

    def __init__(self, work, start, kind):
        self.work, self.start, self.kind = work, start, kind

    def wait(self):
        self.work.wait()
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        EXCHANGE_TIMING["events"].append((self.kind, self.start, end))


def _timing_start():
    if not EXCHANGE_TIMING["on"]:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def exchange_timing_read():
    """{kind: total ms} of the collectives recorded since the last read (synchronises their events)"""
    out = {}
    for kind, a, b in EXCHANGE_TIMING["events"]:
        b.synchronize()
        out[kind] = out.get(kind, 0.0) + a.elapsed_time(b)
    EXCHANGE_TIMING["events"] = []
    return out


def remap_sources(src, row_ranges, max_rows):
    """Global source id -> row of the padded all-gathered feature matrix [G * max_rows, d]
    (rank r's vertices occupy rows r*max_rows .. r*max_rows + n_r)."""
    lows = torch.tensor([lo for lo, hi in row_ranges], device=src.device, dtype=torch.int64)
    owner = torch.searchsorted(lows, src, right=True) - 1
    return owner * max_rows + (src - lows[owner])


class _AllGatherRows(torch.autograd.Function):
    """[max_rows, d] per rank -> [G * max_rows, d]; backward = reduce-scatter (sum) of the gradient."""

    @staticmethod
    def forward(ctx, x_pad, group):
        ctx.group = group
        G = dist.get_world_size(group)
        out = torch.empty((G * x_pad.shape[0],) + tuple(x_pad.shape[1:]), dtype=x_pad.dtype, device=x_pad.device)
        dist.all_gather_into_tensor(out, x_pad.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        group = ctx.group
        G = dist.get_world_size(group)
        rows = g.shape[0] // G
        g = g.contiguous()
        if dist.get_backend(group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep our block
            dist.all_reduce(g, group=group)
            r = dist.get_rank(group)
            return g[r * rows:(r + 1) * rows].clone(), None
        out = torch.empty((rows,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g, group=group)
        return out, None


def all_gather_rows(x_local, max_rows, group=None):
    """Pad the local block to max_rows rows, all-gather, return [G * max_rows, d] (autograd-aware)."""
    n = x_local.shape[0]
    if n < max_rows:
        pad = torch.zeros((max_rows - n,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        x_local = torch.cat((x_local, pad), dim=0)
    return _AllGatherRows.apply(x_local, group)


class RowExchange:
    """Exchange of PROJECTED rows around the fused kernels (ops.FSWEmbedFunction looks for `plan.exchange`).

    Each rank projects only its own vertices (K1 scales with the number of GPUs instead of being replicated);
    `gather` all-gathers the padded [max_rows, ldp] blocks into the [G * max_rows, ldp] matrix the re-indexed
    sources point into, `scatter` is its adjoint: reduce-scatter (sum) of the gradient, cut back to the local rows."""

    def __init__(self, max_rows, group=None, chunks=1):
        self.max_rows = int(max_rows)
        self.group = group
        # column chunks of the slices: with > 1 the exchange of one chunk runs under the kernels of another.  Measured
        # on 8 B200s (profiles/r1/README.md): 84 ms/step with 1 chunk, 92 with 2, 122 with 4 - the concurrent NCCL
        # kernels slow the sort kernels down by more than the overlap hides, and narrow chunks waste lanes.
        self.chunks = int(chunks)

    def gather(self, xp_local):
        n = xp_local.shape[0]
        G = dist.get_world_size(self.group)
        if n < self.max_rows:
            pad = torch.zeros((self.max_rows - n,) + tuple(xp_local.shape[1:]), dtype=xp_local.dtype, device=xp_local.device)
            xp_local = torch.cat((xp_local, pad), dim=0)
        out = torch.empty((G * self.max_rows,) + tuple(xp_local.shape[1:]), dtype=xp_local.dtype, device=xp_local.device)
        dist.all_gather_into_tensor(out, xp_local.contiguous(), group=self.group)
        return out

    def gather_async(self, xp_local):
        """-> (gathered [G * max_rows, ld], work): the collective runs on NCCL's stream after everything queued so far
        on the current stream; `work.wait()` makes the current stream wait for it."""
        n = xp_local.shape[0]
        G = dist.get_world_size(self.group)
        if n < self.max_rows:
            pad = torch.zeros((self.max_rows - n,) + tuple(xp_local.shape[1:]), dtype=xp_local.dtype, device=xp_local.device)
            xp_local = torch.cat((xp_local, pad), dim=0)
        out = torch.empty((G * self.max_rows,) + tuple(xp_local.shape[1:]), dtype=xp_local.dtype, device=xp_local.device)
        t0 = _timing_start()
        work = dist.all_gather_into_tensor(out, xp_local.contiguous(), group=self.group, async_op=True)
        return out, (work if t0 is None else _TimedWork(work, t0, "all_gather_rows"))

    def scatter_async(self, g_all, n_local):
        """-> (summed block [max_rows, ld] (valid after work.wait(); take [:n_local]), work)"""
        G = dist.get_world_size(self.group)
        rows = g_all.shape[0] // G
        g_all = g_all.contiguous()
        if dist.get_backend(self.group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep our block
            work = dist.all_reduce(g_all, group=self.group, async_op=True)
            r = dist.get_rank(self.group)
            return g_all[r * rows:(r + 1) * rows], work
        out = torch.empty((rows,) + tuple(g_all.shape[1:]), dtype=g_all.dtype, device=g_all.device)
        t0 = _timing_start()
        work = dist.reduce_scatter_tensor(out, g_all, group=self.group, async_op=True)
        return out, (work if t0 is None else _TimedWork(work, t0, "reduce_scatter_rows"))

    def scatter(self, g_all, n_local):
        G = dist.get_world_size(self.group)
        rows = g_all.shape[0] // G
        g_all = g_all.contiguous()
        if dist.get_backend(self.group) == "gloo":  # gloo has no reduce_scatter: all-reduce and keep our block
            dist.all_reduce(g_all, group=self.group)
            r = dist.get_rank(self.group)
            return g_all[r * rows:r * rows + n_local].clone()
        out = torch.empty((rows,) + tuple(g_all.shape[1:]), dtype=g_all.dtype, device=g_all.device)
        dist.reduce_scatter_tensor(out, g_all, group=self.group)
        return out[:n_local]


class ShardedGraph:
    """This rank's share of a graph: destination rows [row_lo, row_hi) with all their in-edges."""

    def __init__(self, edge_index_local, row_ranges, rank, thresh, dtype, group=None):
        self.group = group
        self.rank = rank
        self.row_ranges = list(row_ranges)
        self.row_lo, self.row_hi = self.row_ranges[rank]
        self.n_local = self.row_hi - self.row_lo
        self.max_rows = max(hi - lo for lo, hi in self.row_ranges)
        src, dst = edge_index_local[0], edge_index_local[1]
        assert bool(((dst >= self.row_lo) & (dst < self.row_hi)).all()), "edge_index_local holds foreign destinations"
        col = remap_sources(src, self.row_ranges, self.max_rows)
        ei = torch.stack((col, dst - self.row_lo), dim=0).contiguous()
        self.num_edges = int(ei.shape[1])
        self.csr = _graph.GraphCSR(ei, self.n_local, 0, "unit", dtype, num_sources=len(self.row_ranges) * self.max_rows)
        self.plan = self.csr.plan(thresh, dtype)
        # project locally, all-gather the projected rows; FSW_EXCHANGE_CHUNKS: column chunks whose exchange overlaps the kernels
        self.plan.exchange = RowExchange(self.max_rows, group, chunks=int(os.environ.get("FSW_EXCHANGE_CHUNKS", "1")))


def sharded_conv_forward(conv, x_local, sg):
    """One FSW_conv layer on this rank's destination rows (unit edge weights, no edge features)."""
    assert conv.edgefeat_dim == 0 and conv.edge_weighting == "unit" and not (conv.self_loop_weight > 0), \
        "the sharded path covers the default FSW_conv configuration (unit weights, no self loops / edge features)"
    emb = conv.fsw_embed.embed_plan(x_local, sg.plan, None)   # the row exchange happens inside (plan.exchange)
    return conv._combine(emb, x_local)


def all_reduce_gradients(modules, group=None):
    """Sum the parameter gradients over the ranks (one flat all-reduce).  Every trainable parameter takes part on every
    rank - a rank whose shard produced no gradient for one (e.g. an empty shard) contributes zeros - so that the flat
    buffers have the same layout everywhere."""
    params = [p for m in modules for p in m.parameters() if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    t0 = _timing_start()
    dist.all_reduce(flat, group=group)
    if t0 is not None:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        EXCHANGE_TIMING["events"].append(("all_reduce_params", t0, end))
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
