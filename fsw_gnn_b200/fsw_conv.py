"""FSW_conv / FSW_readout - host-side mirror of the reference layers (fsw_conv.py:54-517).

Same constructor keywords, parameter names (`fsw_embed.*`, `mlp.*`, `dim_reduct`, `bn_final.*`,
`size_coeff`) and `forward(vertex_features, edge_index, edge_features=None)` /
`forward(vertex_features, graph_index=None, batch_size=None)`.

The graph preparation (edge list -> destination-major CSR + plan, K0) runs once per distinct
edge_index and is shared by all layers through a small cache (the reference rebuilds a coalesced
sparse adjacency on every forward, fsw_conv.py:352, :384-447); the neighbourhood embedding is the
fused K1/K2/K3 path; the concat + Linear layers after it run the K1 contraction kernels with the concatenation
fused into the first contraction; BatchNorm / activations / dropout stay torch modules.

torch_geometric is optional: when it is installed the classes derive from MessagePassing and are
registered in GraphGym under the reference's names ('fsw_conv', 'fsw_readout'); when it is not
(as in this image) they derive from torch.nn.Module.  Only the base class and the two registration
decorators of PyG are used by the reference (fsw_conv.py:4-9, :374-381).
"""
import inspect

import numpy as np
import torch

from . import graph as _graph
from .fsw_embedding import FSW_embedding, minimize_mutual_coherence, sp  # noqa: F401  (re-exported like fsw_conv.py:33-36)

try:  # pragma: no cover - not installed in this image
    from torch_geometric.nn import MessagePassing as _Base
    from torch_geometric.graphgym.register import register_layer, register_pooling
    _HAVE_PYG = True
except Exception:  # ImportError or a broken install
    _HAVE_PYG = False

    class _Base(torch.nn.Module):
        def __init__(self, aggr=None, **kwargs):
            super().__init__()

    def register_layer(name):
        return lambda cls: cls

    def register_pooling(name):
        return lambda cls: cls


@register_layer("fsw_conv")
class FSW_conv(_Base):
    def __init__(self,
                 in_channels, out_channels, edgefeat_dim=0,
                 embed_dim=None, learnable_embedding=True,
                 encode_vertex_degrees=True, vertex_degree_encoding_function="identity",
                 vertex_degree_encoding_scale=1.0, learnable_vertex_degree_encoding_scale=False, homog_degree_encoding=False,
                 vertex_degree_pad_thresh=1.0,
                 concat_self=True, message_weight_vs_self=1.0,
                 bias=True,
                 mlp_layers=1, mlp_hidden_dim=None,
                 mlp_activation_final=torch.nn.LeakyReLU(negative_slope=0.2),
                 mlp_activation_hidden=torch.nn.LeakyReLU(negative_slope=0.2),
                 mlp_init=None,
                 batchNorm_final=False, batchNorm_hidden=False,
                 dropout_final=0, dropout_hidden=0,
                 self_loop_weight=0, edge_weighting="unit",
                 device=None, dtype=torch.float32,
                 config=None):
        super().__init__(aggr=None)
        config = dict(config) if config is not None else {}
        arg_names = {p.name for p in inspect.signature(FSW_conv.__init__).parameters.values()} - {"config", "self"}
        for key in config:
            if key not in arg_names:
                raise ValueError(f"Invalid argument '{key}' in config")
        given = locals()
        for name in arg_names:
            if name not in config:
                config[name] = given[name]
        self.init_helper(**config)

    def init_helper(self, in_channels, out_channels, edgefeat_dim, embed_dim, learnable_embedding,
                    encode_vertex_degrees, vertex_degree_encoding_function, vertex_degree_encoding_scale,
                    learnable_vertex_degree_encoding_scale, homog_degree_encoding, vertex_degree_pad_thresh,
                    concat_self, message_weight_vs_self, bias, mlp_layers, mlp_hidden_dim, mlp_activation_final,
                    mlp_activation_hidden, mlp_init, batchNorm_final, batchNorm_hidden, dropout_final, dropout_hidden,
                    self_loop_weight, edge_weighting, device, dtype):
        assert edge_weighting in {"unit", "gcn"}, "invalid value passed in argument <edge_weighting>"
        assert vertex_degree_encoding_function in {"identity", "sqrt", "log"}, \
            "invalid value passed in argument <vertex_degree_encoding_function>"
        if mlp_hidden_dim is None:
            mlp_hidden_dim = max(in_channels, out_channels)
        if (mlp_layers == 0) and (not concat_self):
            embed_dim = out_channels
        elif embed_dim is None:
            embed_dim = 2 * max(in_channels, out_channels)
        embedding_bias = bool(bias and mlp_layers == 0)
        tm_method = "homog" if homog_degree_encoding else "plain"

        self.edgefeat_dim = edgefeat_dim
        self.concat_self = concat_self
        self.edge_weighting = edge_weighting
        self.self_loop_weight = self_loop_weight
        self.message_weight_vs_self = message_weight_vs_self
        self.vertex_degree_pad_thresh = float(vertex_degree_pad_thresh)
        self.cache_graph = True  # share K0 between layers / steps (see graph.cached_graph)

        mlp_input_dim = in_channels + embed_dim if concat_self else embed_dim
        if mlp_layers == 0:
            self.mlp = None
            if concat_self:
                with torch.no_grad():
                    dim_reduct = torch.randn(size=(out_channels, mlp_input_dim), device=device, dtype=dtype)
                    dim_reduct = minimize_mutual_coherence(dim_reduct, report=False)
                self.dim_reduct = torch.nn.Parameter(dim_reduct, requires_grad=learnable_embedding)
            self.bn_final = torch.nn.BatchNorm1d(num_features=out_channels, device=device, dtype=dtype) if batchNorm_final else None
        else:
            self.bn_final = None
            mods = []
            for i in range(mlp_layers):
                last = (i == mlp_layers - 1)
                in_curr = mlp_input_dim if i == 0 else mlp_hidden_dim
                out_curr = out_channels if last else mlp_hidden_dim
                lin = torch.nn.Linear(in_curr, out_curr, bias=bias, device=device, dtype=dtype)
                if mlp_init is None:
                    pass
                elif mlp_init == "xavier_uniform":
                    torch.nn.init.xavier_uniform_(lin.weight)
                elif mlp_init == "xavier_normal":
                    torch.nn.init.xavier_normal_(lin.weight)
                elif mlp_init == "kaiming_uniform":
                    torch.nn.init.kaiming_uniform_(lin.weight)
                elif mlp_init == "kaiming_normal":
                    torch.nn.init.kaiming_normal_(lin.weight)
                else:
                    raise RuntimeError("Invalid value passed at argument mlp_init")
                if (mlp_init is not None) and bias:
                    torch.nn.init.zeros_(lin.bias)
                mods.append(lin)
                if (batchNorm_final if last else batchNorm_hidden):
                    mods.append(torch.nn.BatchNorm1d(num_features=out_curr, device=device, dtype=dtype))
                act = mlp_activation_final if last else mlp_activation_hidden
                if act is not None:
                    mods.append(act)
                drop = dropout_final if last else dropout_hidden
                if drop > 0:
                    mods.append(torch.nn.Dropout(p=drop))
            self.mlp = torch.nn.Sequential(*mods)

        self.size_coeff = torch.nn.Parameter(torch.ones(1, device=device, dtype=dtype) / np.sqrt(embed_dim),
                                             requires_grad=learnable_embedding)
        self.fsw_embed = FSW_embedding(d_in=in_channels, d_out=embed_dim, d_edge=edgefeat_dim,
                                       learnable_slices=learnable_embedding, learnable_freqs=learnable_embedding,
                                       learnable_total_mass_encoding_scale=learnable_vertex_degree_encoding_scale,
                                       encode_total_mass=encode_vertex_degrees,
                                       total_mass_encoding_function=vertex_degree_encoding_function,
                                       total_mass_encoding_scale=vertex_degree_encoding_scale,
                                       total_mass_encoding_method=tm_method,
                                       total_mass_pad_thresh=vertex_degree_pad_thresh,
                                       minimize_slice_coherence=True, freqs_init="spread",
                                       enable_bias=embedding_bias, device=device, dtype=dtype)
        device = device if device is not None else self.fsw_embed.get_device()
        dtype = dtype if dtype is not None else self.fsw_embed.get_dtype()
        self.to(device=device, dtype=dtype)

    # ------------------------------------------------------------------------------------------
    def forward(self, vertex_features, edge_index, edge_features=None):
        emb_mod = self.fsw_embed
        assert vertex_features.dtype == emb_mod.get_dtype(), \
            "vertex_features has incorrect dtype (expected %s, got %s)" % (emb_mod.get_dtype(), vertex_features.dtype)
        assert vertex_features.device == emb_mod.get_device(), \
            "vertex_features has incorrect device (expected %s, got %s)" % (emb_mod.get_device(), vertex_features.device)
        assert edge_index.device == emb_mod.get_device(), \
            "edge_index has incorrect device (expected %s, got %s)" % (emb_mod.get_device(), edge_index.device)
        n = vertex_features.size(0)
        num_edges = edge_index.shape[1]
        if self.edgefeat_dim > 0:
            assert edge_features is not None, "Edge features must be provided since edgefeat_dim > 0"
            assert edge_features.dim() in (1, 2), "edge_features should have the shape (num_edges, edegfeat_dim)"
            if self.edgefeat_dim == 1:
                assert tuple(edge_features.shape) in {(num_edges,), (num_edges, 1)}, \
                    "edge_features should have the shape (num_edges, edegfeat_dim) (or optionally (num_edges,) in the case edgefeat_dim=1)"
            else:
                assert tuple(edge_features.shape) == (num_edges, self.edgefeat_dim), \
                    "edge_features must have the shape (num_edges, edgefeat_dim)"
        else:
            assert edge_features is None, "Edge features should not be provided since edgefeat_dim = 0"

        csr, plan = _graph.cached_graph(edge_index, n, self.self_loop_weight, self.edge_weighting,
                                        emb_mod.total_mass_pad_thresh, vertex_features.dtype, use_cache=self.cache_graph,
                                        coalesce=(edge_features is not None))
        E_feat = csr.edge_features_in_slot_order(edge_features) if edge_features is not None else None
        if torch.is_grad_enabled() and vertex_features.dtype == torch.float32 and \
                (vertex_features.requires_grad or any(p.requires_grad for p in emb_mod.parameters())):
            plan.transpose_async(n)   # training: the backward's transposed structure is built under the forward kernels
        emb = emb_mod.embed_plan(vertex_features, plan, E_feat)
        return self._combine(emb, vertex_features)

    def _combine(self, emb, vertex_features):
        """fsw_conv.py:357-369.  fp32: the concatenation is never materialised - the first Linear (or `dim_reduct`) contracts over
        the columns of `emb` and of `vertex_features` in one K1 launch (tensor cores for large graphs), and the remaining Linear
        layers run the same kernels; BatchNorm / activation / dropout stay torch modules.  fp64 keeps torch's matmul."""
        from . import ops as _ops
        fused = emb.dtype == torch.float32 and emb.is_cuda
        parts = [emb]
        if self.concat_self:
            if self.message_weight_vs_self != 1.0:
                emb = self.message_weight_vs_self * emb
            parts = [emb, vertex_features]
        if not fused:
            parts = [torch.cat(parts, dim=-1)] if len(parts) > 1 else parts
        if self.mlp is not None:
            out = None
            for mod in self.mlp:
                if fused and isinstance(mod, torch.nn.Linear):
                    out = _ops.linear_cat(parts if out is None else [out], mod.weight, mod.bias)
                else:
                    out = mod(parts[0] if out is None else out)
            if out is None:   # an empty Sequential cannot occur (mlp_layers >= 1), kept for safety
                out = parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)
        elif self.concat_self:
            if fused:
                out = _ops.linear_cat(parts, self.dim_reduct, None)
            else:
                out = torch.matmul(parts[0], self.dim_reduct.transpose(0, 1))
        else:
            out = emb
        if self.bn_final is not None:
            out = self.bn_final(out)
        return out

    # stubs kept for interface parity with the reference (fsw_conv.py:374-381)
    def aggregate(self, inputs, index):
        return

    def message(self, x_j):
        return

    def update(self, aggr_out):
        return

    @staticmethod
    def edge_index_to_adj(edge_index, edge_features, num_vertices, edgefeat_dim, dtype, self_loop_weight=0,
                          edge_weighting="unit"):
        """Same return triple as the reference (fsw_conv.py:384-447): coalesced sparse adjacency
        (rows = destinations), sparse edge-feature tensor or None, in-degrees [N, 1].  Provided for
        callers that used the static method; FSW_conv.forward itself goes through the cached CSR."""
        num_edges = edge_index.shape[1]
        inds = edge_index.flip(0)
        vals = torch.ones(num_edges, device=edge_index.device, dtype=dtype)
        if self_loop_weight > 0:
            loops = torch.arange(num_vertices, device=edge_index.device).reshape(1, -1).repeat(2, 1)
            inds = torch.cat((inds, loops), dim=1)
            vals = torch.cat((vals, self_loop_weight * torch.ones(num_vertices, device=edge_index.device, dtype=dtype)))
        adj = torch.sparse_coo_tensor(indices=inds, values=vals, size=(num_vertices, num_vertices)).coalesce()
        rows, cols = adj.indices()
        in_degrees = torch.zeros(num_vertices, device=edge_index.device, dtype=dtype).index_add_(0, rows, adj.values())
        if edge_weighting == "gcn":
            d = torch.sqrt(in_degrees)
            adj = torch.sparse_coo_tensor(adj.indices(), adj.values() / d[rows] / d[cols], adj.shape, is_coalesced=True)
        elif edge_weighting != "unit":
            raise RuntimeError("Invalid weighting method passed in argument <edge_weighting>")
        X_edge = None
        if edgefeat_dim > 0:
            assert edge_features is not None, "Edge features must be provided since edgefeat_dim > 0"
            shape = tuple(adj.shape) if edge_features.dim() == 1 else tuple(adj.shape) + (edgefeat_dim,)
            if self_loop_weight > 0:
                s = list(edge_features.shape)
                s[0] = num_vertices
                edge_features = torch.cat((edge_features, torch.zeros(s, device=edge_index.device, dtype=dtype)), dim=0)
            X_edge = torch.sparse_coo_tensor(indices=inds, values=edge_features, size=shape).coalesce()
        else:
            assert edge_features is None, "Edge features should not be provided since edgefeat_dim = 0"
        return adj, X_edge, in_degrees.unsqueeze(-1)


@register_pooling("fsw_readout")
class FSW_readout(FSW_conv):
    """Global pooling: one multiset per graph of the batch (fsw_conv.py:451-517)."""

    def forward(self, vertex_features, graph_index=None, batch_size=None):
        emb_mod = self.fsw_embed
        assert self.edgefeat_dim == 0, "edgefeat_dim should equal zero in a global readout layer"
        num_vertices = vertex_features.shape[0]
        if graph_index is None:
            assert batch_size is None, "batch_size must be None when graph_index is None"
            graph_index = torch.zeros(num_vertices, device=vertex_features.device, dtype=torch.int64)
        else:
            assert tuple(graph_index.shape) == (num_vertices,), \
                "graph_index should be of shape (num_vertices,), where vertex_features is of shape (num_features, vertex_feature_dimension)"
            assert is_monotone_increasing(graph_index), "for efficiency, graph_index should be monotone non-decreasing"
        batch_size = graph_index.max().item() + 1 if batch_size is None else batch_size
        assert (graph_index < batch_size).all(), "all entries of graph_index must be in the range 0,...,batch_size-1"
        assert (graph_index >= 0).all(), "all entries of graph_index must be in the range 0,...,batch_size-1"
        assert vertex_features.device == emb_mod.get_device(), "invalid device given in vertex_features"
        assert graph_index.device == emb_mod.get_device(), "invalid device given in graph_index"
        assert vertex_features.dtype == emb_mod.get_dtype(), "invalid dtype given in vertex_features"
        assert graph_index.dtype == torch.int64, "invalid dtype given in graph_index (expected torch.int64)"

        # vertex v belongs to segment graph_index[v]; vertices of one graph are contiguous, so the
        # CSR is just rowptr over the sorted graph_index with identity columns
        dtype, device = vertex_features.dtype, vertex_features.device
        key = ("readout", graph_index.data_ptr(), graph_index._version, num_vertices, int(batch_size),
               emb_mod.total_mass_pad_thresh, dtype)

        def build():
            rowptr = _graph.rowptr_from_sorted_rows(graph_index, int(batch_size))
            from .ops import SegmentPlan
            return SegmentPlan(int(batch_size), num_vertices, rowptr, 0, None, None, emb_mod.total_mass_pad_thresh, dtype, device)

        plan = emb_mod._cached_plan(key, graph_index, build)
        emb = emb_mod.embed_plan(vertex_features, plan, None)
        if self.mlp is not None:
            out = self.mlp(emb)
        elif self.concat_self:
            out = torch.matmul(emb, self.dim_reduct.transpose(0, 1))
        else:
            out = emb
        return out


def is_monotone_increasing(tensor):
    return torch.all(tensor[1:] - tensor[:-1] >= 0)
