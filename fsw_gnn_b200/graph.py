"""Builders of SegmentPlan from the input formats the reference accepts.

 * dense batches of point clouds (fsw_embedding.py:713-732),
 * coalesced sparse-COO weight tensors, graph mode or not (fsw_embedding.py:664-668, :734-757),
 * an `edge_index` [2, E] as in FSW_conv.edge_index_to_adj (fsw_conv.py:384-447).
"""
import torch

from . import _lib
from ._lib import dtype_code, ptr, stream_ptr
from .ops import SegmentPlan, _ws


def plan_dense(batch, n, W_flat, thresh, dtype, device):
    """`batch` multisets of exactly n points, stored contiguously; W_flat [batch*n] or None (unit)."""
    return SegmentPlan(batch, batch * n, None, n, None, W_flat, thresh, dtype, device)


def rowptr_from_sorted_rows(rows, S):
    lib = _lib.load()
    rows = rows.contiguous()
    rowptr = torch.empty(S + 1, dtype=torch.int32, device=rows.device)
    _lib.call(rows.device, "fsw_rowptr_from_sorted_rows", ptr(rows), rows.numel(), S, ptr(rowptr), stream_ptr(rows.device))
    return rowptr


def plan_from_coo(indices, values, shape, graph_mode, thresh, dtype):
    """Coalesced COO weights of shape (<batch>, n) or, in graph mode, (<batch>, S, n).

    Segment id = ravel of every index but the last (lexicographically sorted because the tensor is
    coalesced), element's point row = batch_flat * n + last index.  This is what
    sp.get_slice_info (fsw_embedding.py:2586-2678) recomputes with a stable sort on every call."""
    device = values.device
    nd = len(shape)
    n = int(shape[-1])
    seg = torch.zeros(indices.shape[1], dtype=torch.int64, device=device)
    nseg = 1
    for a in range(nd - 1):
        seg = seg * int(shape[a]) + indices[a]
        nseg *= int(shape[a])
    if graph_mode:
        batch_flat = torch.zeros_like(seg)
        for a in range(nd - 2):
            batch_flat = batch_flat * int(shape[a]) + indices[a]
    else:
        batch_flat = seg
    col64 = batch_flat * n + indices[nd - 1]
    nrows_points = (nseg // int(shape[-2]) if graph_mode else nseg) * n
    if nrows_points >= 2 ** 31 or indices.shape[1] >= 2 ** 31:
        raise RuntimeError("more than 2^31 points / nonzeros are not supported")
    rowptr = rowptr_from_sorted_rows(seg, nseg)
    col = col64.to(torch.int32)
    W = values.contiguous()
    return SegmentPlan(nseg, W.numel(), rowptr, 0, col, W, thresh, dtype, device)


def validate_edge_index(edge_index, num_vertices, num_sources=None):
    """Every id must address an existing row: destinations in [0, num_vertices), sources in [0, num_sources).
    The reference raises here as well (sparse_coo_tensor / coalesce check the indices against the size,
    fsw_conv.py:397-398); the CSR kernels themselves do not bounds-check.  One pass over the edge list and one small
    D2H read, on the cache-miss path only (a new graph is followed by the plan's own D2H read anyway)."""
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError("edge_index must have shape [2, E], got %s" % (tuple(edge_index.shape),))
    E = int(edge_index.shape[1])
    if E >= 2 ** 31 - int(num_vertices):
        raise RuntimeError("more than 2^31 edges are not supported (int32 CSR)")
    if E == 0:
        return
    num_sources = int(num_vertices) if num_sources is None else int(num_sources)
    lo, hi = torch.aminmax(edge_index, dim=1)
    src_lo, dst_lo, src_hi, dst_hi = [int(v) for v in torch.stack((lo, hi)).reshape(-1).tolist()]
    if src_lo < 0 or dst_lo < 0 or src_hi >= num_sources or dst_hi >= int(num_vertices):
        raise RuntimeError("edge_index out of range: sources in [%d, %d] (need [0, %d)), destinations in [%d, %d] (need [0, %d))"
                           % (src_lo, src_hi, num_sources, dst_lo, dst_hi, int(num_vertices)))


class GraphCSR:
    """Destination-major CSR of a graph given as edge_index (fsw_conv.py:384-409)."""

    def __init__(self, edge_index, num_vertices, self_loop_weight, edge_weighting, dtype, coalesce=False, num_sources=None):
        """coalesce=False: duplicate (dst, src) pairs stay separate elements (exact for the embedding and
        its gradients, see include/fsw_embedding.h section 3).  coalesce=True merges them like the
        reference's `coalesce()` (fsw_conv.py:397-398, :438-439) - needed with edge features, where
        the merged element carries the SUM of the duplicates' feature vectors.
        num_sources: rows of the point matrix the sources index (default num_vertices; the sharded path re-indexes its
        sources into the gathered layout of G * max_rows rows)."""
        self.coalesced = bool(coalesce)
        _lib.require_cuda(edge_index, "edge_index")
        validate_edge_index(edge_index, num_vertices, num_sources)
        if coalesce:
            self._init_coalesced(edge_index, num_vertices, self_loop_weight, edge_weighting, dtype)
            return
        lib = _lib.load()
        assert edge_weighting in {"unit", "gcn"}, "invalid value passed in argument <edge_weighting>"
        device = edge_index.device
        ei = edge_index.contiguous()
        if ei.dtype != torch.int64:
            ei = ei.to(torch.int64)
        E = int(ei.shape[1])
        N = int(num_vertices)
        self_loops = 1 if self_loop_weight > 0 else 0
        Etot = E + (N if self_loops else 0)
        self.N, self.E, self.Etot = N, E, Etot
        self.rowptr = torch.empty(N + 1, dtype=torch.int32, device=device)
        self.col = torch.empty(max(Etot, 1), dtype=torch.int32, device=device)[:Etot]
        self.eid = torch.empty(max(Etot, 1), dtype=torch.int32, device=device)[:Etot]
        ws = _ws(lib.fsw_csr_workspace_bytes(N, E), device)
        _lib.call(device, "fsw_csr_from_edge_index", ptr(ei), E, N, self_loops, ptr(self.rowptr), ptr(self.col), ptr(self.eid),
                  ptr(ws), ws.numel(), stream_ptr(device))
        gcn = 1 if edge_weighting == "gcn" else 0
        self.in_degrees = torch.empty(N, dtype=dtype, device=device)
        self.W = torch.empty(Etot, dtype=dtype, device=device) if (gcn or self_loops) else None
        _lib.call(device, "fsw_edge_weights", dtype_code(dtype), ptr(self.rowptr), ptr(self.col), ptr(self.eid), N, E, self_loops,
                  float(self_loop_weight), gcn, ptr(self.in_degrees), ptr(self.W), stream_ptr(device))

    def _init_coalesced(self, edge_index, num_vertices, self_loop_weight, edge_weighting, dtype):
        """fsw_csr_coalesce: stable 64-bit radix sort of dst * N + src, one element per distinct pair with the summed base weight
        (replaces the reference's sparse_coo_tensor(...).coalesce(), fsw_conv.py:397-398); one small D2H read for the element count"""
        _lib.require_cuda(edge_index, "edge_index")
        assert edge_weighting in {"unit", "gcn"}, "invalid value passed in argument <edge_weighting>"
        lib = _lib.load()
        device = edge_index.device
        N, E = int(num_vertices), int(edge_index.shape[1])
        ei = edge_index.contiguous()
        if ei.dtype != torch.int64:
            ei = ei.to(torch.int64)
        self_loops = 1 if self_loop_weight > 0 else 0
        cap = max(E + (N if self_loops else 0), 1)
        self.rowptr = torch.empty(N + 1, dtype=torch.int32, device=device)
        col = torch.empty(cap, dtype=torch.int32, device=device)
        W = torch.empty(cap, dtype=dtype, device=device)
        slot_of_elem = torch.empty(cap, dtype=torch.int32, device=device)
        self.in_degrees = torch.empty(N, dtype=dtype, device=device)
        nslots_dev = torch.zeros(1, dtype=torch.int32, device=device)
        ws = _ws(lib.fsw_csr_coalesce_workspace_bytes(N, E), device)
        _lib.call(device, "fsw_csr_coalesce", dtype_code(dtype), ptr(ei), E, N, self_loops, float(self_loop_weight),
                  1 if edge_weighting == "gcn" else 0, ptr(self.rowptr), ptr(col), ptr(W), ptr(slot_of_elem), ptr(self.in_degrees),
                  ptr(nslots_dev), ptr(ws), ws.numel(), stream_ptr(device))
        nslots = int(nslots_dev.item())
        self.N, self.E, self.Etot = N, E, nslots
        self.col = col[:nslots]
        self.W = W[:nslots]
        self.eid = None
        self.slot_of_edge = slot_of_elem[:E].to(torch.int64)

    def plan(self, thresh, dtype):
        return SegmentPlan(self.N, self.Etot, self.rowptr, 0, self.col, self.W, thresh, dtype, self.rowptr.device)

    def edge_features_in_slot_order(self, edge_features):
        """[E, d_edge] per input edge -> [E', d_edge] per CSR slot (zeros for self loops,
        fsw_conv.py:430-439)."""
        if edge_features.dim() == 1:
            edge_features = edge_features.unsqueeze(-1)
        if self.coalesced:
            out = torch.zeros((self.Etot, edge_features.shape[1]), dtype=edge_features.dtype, device=edge_features.device)
            return out.index_add(0, self.slot_of_edge, edge_features)
        if self.Etot > self.E:
            pad = torch.zeros((self.N, edge_features.shape[1]), dtype=edge_features.dtype, device=edge_features.device)
            edge_features = torch.cat((edge_features, pad), dim=0)
        return edge_features.index_select(0, self.eid.to(torch.int64))


# ------------------------------------------------------------------------------------------------
# Small cache so that the layers of one network (and successive steps) share the graph preparation.
# The reference rebuilds the adjacency on every forward (fsw_conv.py:352).  Entries keep a reference
# to the edge_index tensor, so a cached data_ptr can never be recycled for different contents, and
# `_version` catches in-place edits.
# ------------------------------------------------------------------------------------------------
_GRAPH_CACHE = []
_GRAPH_CACHE_SIZE = 2   # the layers of a network share one graph; two entries cover alternating train / validation graphs


def cached_graph(edge_index, num_vertices, self_loop_weight, edge_weighting, thresh, dtype, use_cache=True, coalesce=False):
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), edge_index.dtype, int(num_vertices),
           float(self_loop_weight), edge_weighting, float(thresh), dtype, edge_index.device, bool(coalesce))
    if use_cache:
        for i, (k, ref, csr, plan) in enumerate(_GRAPH_CACHE):
            if k == key:
                _GRAPH_CACHE.append(_GRAPH_CACHE.pop(i))
                return csr, plan
    csr = GraphCSR(edge_index, num_vertices, self_loop_weight, edge_weighting, dtype, coalesce=coalesce)
    plan = csr.plan(thresh, dtype)
    if use_cache:
        _GRAPH_CACHE.append((key, edge_index, csr, plan))
        while len(_GRAPH_CACHE) > _GRAPH_CACHE_SIZE:
            _GRAPH_CACHE.pop(0)
    return csr, plan


def clear_graph_cache():
    del _GRAPH_CACHE[:]
