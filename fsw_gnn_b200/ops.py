"""Host-side wrappers around the C ABI: segment plans, projection, fused embed forward/backward.

Everything here launches kernels of libfsw_embedding.so on torch's current CUDA stream; torch is
used for memory (allocation) and autograd bookkeeping only.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, dtype_code, ptr, stream_ptr


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device)


# Scratch of the embedding kernels (coefficient tables, pre-scaled gradient, merge tiles; up to several GB at configs[3]) is only
# live inside one library call, and calls on one stream are ordered: every plan on a (device, stream) shares one buffer instead
# of holding its own (a training loop that meets a new graph every step otherwise keeps one set per cached graph).  Buffers are
# never released or resized in place - a larger request adds a new one - so pointers baked into a captured CUDA graph stay valid.
_SCRATCH_POOL = {}


def _shared_scratch(nbytes, device):
    dev = torch.device(device)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    bufs = _SCRATCH_POOL.setdefault(key, [])
    for b in bufs:
        if b.numel() >= nbytes:
            return b
    b = _ws(nbytes, device)
    bufs.append(b)
    bufs.sort(key=lambda t: t.numel())
    return b


def round_up(x, m):
    return (x + m - 1) // m * m


RANKT_NMAX = 32768   # FSW_RANKT_NMAX (include/fsw_embedding.h): largest segment served by the source-major backward


_SIDE_STREAMS = {}


def _side_stream(device):
    """One extra stream per device for work that is not on the critical path (fsw_csr_transpose under the forward)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class SegmentPlan:
    """CSR description of a batch of weighted multisets + the launch plan of the fused kernels.

    Replaces the reference's coalesced sparse-COO weight tensor and the `slice_info` dictionaries
    that `sp.get_slice_info` rebuilds 3-4 times per forward (fsw_embedding.py:2586-2678, :2657-2663).

    rowptr [S+1] int32 or None (then every segment has n_fixed elements, stored contiguously)
    col    [E] int32 row of the point matrix for each element, or None (identity)
    W      [E] raw weights in the compute dtype, or None (unit weights)
    """

    def __init__(self, S, E, rowptr, n_fixed, col, W, thresh, dtype, device):
        lib = _lib.load()
        self.S, self.E = int(S), int(E)
        self.rowptr, self.n_fixed, self.col, self.W = rowptr, int(n_fixed), col, W
        self.thresh = float(thresh)
        self.dtype, self.device = dtype, device
        self.mass = torch.empty(self.S, dtype=torch.float64, device=device)
        self.info = torch.empty(self.S, dtype=torch.int32, device=device)
        self.order = torch.empty(self.S, dtype=torch.int32, device=device)
        self._bucket_dev = torch.empty(_lib.PLAN_BUCKETS + 2, dtype=torch.int32, device=device)
        self._elems_dev = torch.empty(_lib.PLAN_BUCKETS, dtype=torch.int64, device=device)
        ws = _ws(lib.fsw_plan_workspace_bytes(self.S), device)
        _lib.call(device, "fsw_segment_plan", dtype_code(dtype), ptr(rowptr), self.n_fixed, ptr(W), self.S, self.thresh,
                  ptr(self.mass), ptr(self.info), ptr(self.order), ptr(self._bucket_dev), ptr(self._elems_dev), ptr(ws),
                  ws.numel(), stream_ptr(device))
        # one small D2H read per new plan (the reference syncs in every get_slice_info, :2655)
        host = self._bucket_dev.cpu()
        self.bucket_offsets = (ctypes.c_int * (_lib.PLAN_BUCKETS + 2))(*host.tolist())
        self.max_n_eff = int(host[_lib.PLAN_BUCKETS + 1])
        self.bucket_counts = [int(host[i + 1] - host[i]) for i in range(_lib.PLAN_BUCKETS)]
        self.bucket_elems = self._elems_dev.cpu().tolist()
        self._scratch = {}
        self._mass_t = None

    def mass_as(self, dtype):
        """Total mass per segment (W_sum of fsw_embedding.py:778-784) in the compute dtype, [S]."""
        if self._mass_t is None or self._mass_t.dtype != dtype:
            self._mass_t = self.mass.to(dtype)
        return self._mass_t

    def scratch(self, K, backward):
        """scratch buffer for one fsw_embed_forward / fsw_embed_backward call (shared per device and stream, see _shared_scratch)"""
        key = (int(K), bool(backward))
        if key not in self._scratch:   # the SIZE is cached per plan, the buffer is shared
            lib = _lib.load()
            nbytes = lib.fsw_embed_scratch_bytes(dtype_code(self.dtype), self.bucket_offsets, int(K), self.max_n_eff,
                                                 1 if backward else 0)
            if backward and self.col is not None:
                nbytes += lib.fsw_embed_backward_extra_bytes(dtype_code(self.dtype), self.S, int(K))
            self._scratch[key] = int(nbytes)
        nbytes = self._scratch[key]
        return _shared_scratch(nbytes, self.device) if nbytes > 0 else None

    def _build_transpose(self, key, side):
        lib = _lib.load()
        dev = self.device
        tptr = torch.empty(key + 1, dtype=torch.int32, device=dev)
        tseg = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)
        tslot = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)
        tn = torch.empty(max(self.E, 1), dtype=torch.int32, device=dev)
        ws = _ws(lib.fsw_transpose_workspace_bytes(key, self.E), dev)

        def launch():
            _lib.call(dev, "fsw_csr_transpose", ptr(self.rowptr), ptr(self.col), ptr(self.info), self.S, key, self.E,
                      RANKT_NMAX,   # FSW_RANKT_ELIGIBLE
                      ptr(tptr), ptr(tseg), ptr(tslot), ptr(tn), ptr(ws), ws.numel(), stream_ptr(dev))

        self._transpose_event = None
        if side is None:
            launch()
        else:
            # the buffers were allocated on the current stream; the side stream starts after everything queued so far
            # (rowptr / col / info are ready there) and every buffer is told about its second stream
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch()
                ev = torch.cuda.Event()
                ev.record(side)
            for t in (tptr, tseg, tslot, tn, ws):
                t.record_stream(side)
            self._transpose_event = ev
        self._transpose = (key, tptr, tseg, tslot, tn)

    def transpose_async(self, nrows):
        """Start fsw_csr_transpose on a side stream: the transposition is first read by the BACKWARD of the last layer, a whole
        forward pass after the graph arrives, so its ~9 ms (configs[3]) need not sit between the graph and the first sort.
        `transpose()` makes the consuming stream wait for it."""
        if self.col is None or self.rowptr is None or self.dtype != torch.float32:
            return
        key = int(nrows)
        if getattr(self, "_transpose", None) is None or self._transpose[0] != key:
            self._build_transpose(key, _side_stream(self.device))

    def transpose(self, nrows):
        """(tptr, tseg, tslot, tn) of fsw_csr_transpose, built lazily and cached (graphs with explicit columns)."""
        if self.col is None or self.rowptr is None:
            return None
        key = int(nrows)
        if getattr(self, "_transpose", None) is None or self._transpose[0] != key:
            self._build_transpose(key, None)
        ev = getattr(self, "_transpose_event", None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            self._transpose_event = None
        return self._transpose[1:]

    def any_deficient(self):
        """True when some segment has total mass below the pad threshold (one small D2H read, cached)."""
        if getattr(self, "_any_def", None) is None:
            self._any_def = bool((self.mass < self.thresh).any().item()) if self.S > 0 else False
        return self._any_def

    def segment_ids(self):
        """[E] int64 segment of every element (for element-wise broadcasts / segment sums in torch)."""
        if getattr(self, "_seg_ids", None) is None:
            if self.rowptr is None:
                self._seg_ids = torch.arange(self.S, device=self.device).repeat_interleave(self.n_fixed)
            else:
                counts = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
                self._seg_ids = torch.repeat_interleave(torch.arange(self.S, device=self.device), counts, output_size=self.E)
        return self._seg_ids

    def expand_to_elements(self, per_segment):
        """[S] -> [E]: the segment's value for each of its elements"""
        return per_segment.index_select(0, self.segment_ids())

    def differentiable_mass(self, W_vals):
        """Total mass per segment as a torch function of the element weights (autograd flows into W)"""
        if self.rowptr is None:
            return W_vals.reshape(self.S, self.n_fixed).sum(dim=1)
        return torch.zeros(self.S, dtype=W_vals.dtype, device=W_vals.device).index_add(0, self.segment_ids(), W_vals)

    def uniform_fraction(self):
        bo = list(self.bucket_offsets)
        return (bo[_lib.PLAN_BUCKETS_PER_KIND] - bo[0]) / max(self.S, 1)


def gemm(op, A, B, M, N, Kd, lda, ldb, out=None, ldc=None, accumulate=False):
    """C = A.B in the input precision (op 0: NT, 1: NN, 2: TN, see include/fsw_embedding.h section 4)."""
    lib = _lib.load()
    if out is None:
        ldc = N if ldc is None else ldc
        out = torch.empty((M, ldc), dtype=A.dtype, device=A.device)
    _lib.call(A.device, "fsw_gemm", dtype_code(A.dtype), op, M, N, Kd, ptr(A), lda, ptr(B), ldb, ptr(out), ldc,
              1 if accumulate else 0, stream_ptr(A.device))
    return out


def gemm_fused(As, Bs, bias=None, out=None, accumulate=False):
    """out[M, N] (+)= sum_s As[s] . Bs[s]^T (+ bias): the first Linear of FSW_conv over cat(emb, x) without the concatenated
    copy (fsw_conv.py:357-361).  As[s] [M, Kd_s], Bs[s] [N, Kd_s] (row-strided views are fine); fp32."""
    lib = _lib.load()
    nseg = len(As)
    M, N = As[0].shape[0], Bs[0].shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=As[0].dtype, device=As[0].device)
    I64 = ctypes.c_int64 * nseg
    VP = ctypes.c_void_p * nseg
    for a, b in zip(As, Bs):
        assert a.stride(1) == 1 and b.stride(1) == 1 and a.shape[1] == b.shape[1] and a.shape[0] == M and b.shape[0] == N
    _lib.call(out.device, "fsw_gemm_fused", dtype_code(out.dtype), M, N, nseg, I64(*[a.shape[1] for a in As]),
              VP(*[a.data_ptr() for a in As]), I64(*[a.stride(0) for a in As]), VP(*[b.data_ptr() for b in Bs]),
              I64(*[b.stride(0) for b in Bs]), ptr(out), out.stride(0), ptr(bias), 1 if accumulate else 0, stream_ptr(out.device))
    return out


class FSWLinearFunction(torch.autograd.Function):
    """out = cat(inputs, dim=1) . W^T + b without the concatenated copy (FSW_conv combine, fsw_conv.py:357-361), fp32.
    Forward: one K1 launch with the contraction axis made of the inputs' column ranges.  Backward: d input_s = g . W[:, seg_s]
    (op 1), dW[:, seg_s] = g^T . input_s (op 2), db = column sums of g."""

    @staticmethod
    def forward(ctx, W, b, *inputs):
        inputs = tuple(x if (x.stride(1) == 1 and x.stride(0) % 4 == 0) else x.contiguous() for x in inputs)
        widths = [x.shape[1] for x in inputs]
        bounds = [0]
        for w_ in widths:
            bounds.append(bounds[-1] + w_)
        Wc = W if W.stride(1) == 1 else W.contiguous()
        segs = [Wc[:, bounds[i]:bounds[i + 1]] for i in range(len(inputs))]
        M, N = inputs[0].shape[0], W.shape[0]
        out = torch.empty((M, N), dtype=W.dtype, device=W.device)
        # up to two segments per launch; more (never in FSW_conv) accumulate
        for i in range(0, len(inputs), 2):
            gemm_fused(list(inputs[i:i + 2]), segs[i:i + 2], bias=b if i == 0 else None, out=out, accumulate=(i > 0))
        ctx.save_for_backward(Wc, *inputs)
        ctx.bounds, ctx.has_bias = bounds, b is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        Wc, *inputs = ctx.saved_tensors
        bounds = ctx.bounds
        if not (g.stride(1) == 1 and g.stride(0) % 4 == 0):
            g = g.contiguous()
        M, N = g.shape
        dW = db = None
        if ctx.needs_input_grad[0]:
            dW = torch.zeros_like(Wc)
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = g.sum(dim=0)
        dins = []
        for i, x in enumerate(inputs):
            k0, k1 = bounds[i], bounds[i + 1]
            if ctx.needs_input_grad[2 + i]:
                dins.append(gemm(1, g, Wc[:, k0:k1], M, k1 - k0, N, g.stride(0), Wc.stride(0)))
            else:
                dins.append(None)
            if dW is not None:
                gemm(2, g, x, N, k1 - k0, M, g.stride(0), x.stride(0), out=dW[:, k0:k1], ldc=dW.stride(0), accumulate=True)
        return (dW, db) + tuple(dins)


def linear_cat(inputs, weight, bias):
    """nn.Linear over the column-wise concatenation of `inputs` (K1 kernels; fp32 CUDA tensors)."""
    return FSWLinearFunction.apply(weight, bias, *inputs)


def project(X, theta_part, ldp):
    """Xp[:, :K] = X . theta_part^T  (fsw_embedding.py:911).  X [N, d] contiguous; theta_part [K, d] view
    with row stride theta_part.stride(0).  Columns K..ldp-1 of the result are zero."""
    Nrows, d = X.shape
    K = theta_part.shape[0]
    if ldp == K:
        out = torch.empty((Nrows, ldp), dtype=X.dtype, device=X.device)
    else:
        out = torch.zeros((Nrows, ldp), dtype=X.dtype, device=X.device)
    if Nrows == 0 or K == 0:
        return out
    return gemm(0, X, theta_part, Nrows, K, d, X.stride(0), theta_part.stride(0), out=out, ldc=ldp)


def embed_forward(plan, Xp, ldp, Ep, freqs, out, ld_out, out_col0, bias, ranks=None, dxi_out=None):
    lib = _lib.load()
    K = freqs.numel()
    scratch = plan.scratch(K, False)
    _lib.call(plan.device, "fsw_embed_forward", dtype_code(plan.dtype), ptr(Xp), ldp, ptr(Ep), ptr(plan.rowptr), plan.n_fixed,
              ptr(plan.col), ptr(plan.W), ptr(plan.mass), ptr(plan.info), ptr(plan.order),
              plan.bucket_offsets, plan.S, K, ptr(freqs), plan.thresh, ptr(out), ld_out, out_col0,
              ptr(bias), plan.max_n_eff, ptr(scratch), 0 if scratch is None else scratch.numel(),
              ptr(ranks), 0 if ranks is None else ranks.stride(0),
              ptr(dxi_out), 0 if dxi_out is None else dxi_out.stride(0), stream_ptr(plan.device))


def embed_backward(plan, Xp, ldp, Ep, freqs, g, ld_g, g_col0, dXp, dEp, dfreqs_acc, ranks=None, dxi_from_forward=False,
                   transpose=None, nrows=0):
    lib = _lib.load()
    K = freqs.numel()
    scratch = plan.scratch(K, True)
    _lib.call(plan.device, "fsw_embed_backward", dtype_code(plan.dtype), ptr(Xp), ldp, ptr(Ep), ptr(plan.rowptr), plan.n_fixed,
              ptr(plan.col), ptr(plan.W), ptr(plan.mass), ptr(plan.info), ptr(plan.order),
              plan.bucket_offsets, plan.S, K, ptr(freqs), plan.thresh, ptr(g), ld_g, g_col0,
              ptr(dXp), ptr(dEp), ptr(dfreqs_acc), None, plan.max_n_eff, ptr(scratch),
              0 if scratch is None else scratch.numel(), ptr(ranks),
              0 if ranks is None else ranks.stride(0), 1 if dxi_from_forward else 0,
              ptr(transpose[0]) if transpose else None, ptr(transpose[1]) if transpose else None,
              ptr(transpose[2]) if transpose else None, ptr(transpose[3]) if transpose else None,
              int(nrows), stream_ptr(plan.device))


def total_mass_function(T, name):
    """fsw_embedding.py:857-865."""
    if name == "identity":
        return T
    if name == "sqrt":
        return 2 * (T / (torch.sqrt(T + 1) + 1))
    if name == "log":
        return torch.log1p(T)
    raise RuntimeError("This should not happen")


def total_mass_function_derivative(T, name):
    """d/dT of total_mass_function."""
    if name == "identity":
        return torch.ones_like(T)
    if name == "sqrt":   # 2 T / (sqrt(T+1) + 1) = 2 (sqrt(T+1) - 1)
        return 1.0 / torch.sqrt(T + 1)
    if name == "log":
        return 1.0 / (1 + T)
    raise RuntimeError("This should not happen")


# Rank saving (see include/fsw_embedding.h, fsw_embed_forward): on by default, bounded by free memory.
SAVE_RANKS = True
RANK_MEMORY_FRACTION = 0.35


CLOUD_MODE = True   # point-cloud mode for dense batches of low-dimensional points (include/fsw_embedding.h section 5b)


def cloud_eligible(X, plan, E_feat, K):
    """dense batch of unit-weight multisets of 33..1024 points, d <= 4, fp32, total mass >= pad threshold, one GPU"""
    return (CLOUD_MODE and X.dtype == torch.float32 and E_feat is None and plan.rowptr is None and plan.col is None
            and plan.W is None and 1 <= X.shape[1] <= 4 and 33 <= plan.n_fixed <= 1024 and plan.n_fixed >= plan.thresh
            and plan.S <= 65535 and K > 0 and getattr(plan, "exchange", None) is None)


class FSWCloudFunction(torch.autograd.Function):
    """FSWEmbedFunction for point clouds: no projected matrix, no projected gradient (csrc/fsw_cloud.cu).
    out[S, tm_dim + K] like FSWEmbedFunction ('plain' total-mass channel and bias included)."""

    @staticmethod
    def forward(ctx, X, projVecs, freqs, bias, tm_scale, plan, tm_function, grad_mode):
        lib = _lib.load()
        d = X.shape[1]
        K = projVecs.shape[0]
        tm_dim = 0 if tm_function is None else 1
        X = X.contiguous()
        freqs = freqs.contiguous()
        theta = projVecs if projVecs.stride(1) == 1 else projVecs.contiguous()
        out = torch.empty((plan.S, K + tm_dim), dtype=X.dtype, device=X.device)
        bias_core = None
        if bias is not None:
            bias = bias.contiguous()
            bias_core = bias[tm_dim:]
        needs_grad = bool(grad_mode) and any(ctx.needs_input_grad[:5])
        n = plan.n_fixed
        ranksT = dxi_out = None
        if needs_grad:
            ranksT = torch.empty(plan.S * K * n, dtype=torch.int16, device=X.device)
            if ctx.needs_input_grad[2]:
                dxi_out = torch.zeros((plan.S, K), dtype=X.dtype, device=X.device)
        scratch = plan.scratch(K, False)
        _lib.call(X.device, "fsw_embed_forward_cloud", dtype_code(X.dtype), ptr(X), d, ptr(theta), theta.stride(0), n, ptr(plan.mass),
                  ptr(plan.info), ptr(plan.order), plan.bucket_offsets, plan.S, K, ptr(freqs), plan.thresh, ptr(out), out.stride(0),
                  tm_dim, ptr(bias_core), ptr(scratch), 0 if scratch is None else scratch.numel(), ptr(ranksT), ptr(dxi_out),
                  0 if dxi_out is None else dxi_out.stride(0), stream_ptr(X.device))
        fT = None
        if tm_dim:
            fT = total_mass_function(plan.mass_as(X.dtype), tm_function)
            col0 = fT * tm_scale
            if bias is not None:
                col0 = col0 + bias[0]
            out[:, 0] = col0
        ctx.plan, ctx.tm_dim = plan, tm_dim
        ctx.has_bias, ctx.has_scale = bias is not None, tm_scale is not None
        ctx.ranksT, ctx.dxi_out = ranksT, dxi_out
        ctx.save_for_backward(X, theta, freqs, fT)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        X, theta, freqs, fT = ctx.saved_tensors
        plan, tm_dim = ctx.plan, ctx.tm_dim
        need_X, need_theta, need_xi, need_bias, need_scale = ctx.needs_input_grad[:5]
        if not (g.dim() == 2 and g.stride(1) == 1):
            g = g.contiguous()
        K, d, n = theta.shape[0], X.shape[1], plan.n_fixed
        dX = dtheta = dxi = dbias = dscale = None
        if need_bias and ctx.has_bias:
            dbias = g.sum(dim=0)
        if need_scale and ctx.has_scale and tm_dim:
            dscale = (g[:, 0] * fT).sum()
        if need_X:
            dX = torch.empty_like(X)
        if need_theta:
            dtheta = torch.zeros_like(theta)
        if (need_X or need_theta) and ctx.ranksT is None:
            raise RuntimeError("FSWCloudFunction.backward: the forward ran without gradient recording")
        if need_X or need_theta:
            _lib.call(X.device, "fsw_embed_backward_cloud", dtype_code(X.dtype), ptr(X), d, ptr(theta), theta.stride(0), n, plan.S, K,
                      ptr(freqs), ptr(g), g.stride(0), tm_dim, ptr(ctx.ranksT), ptr(dX), ptr(dtheta),
                      0 if dtheta is None else dtheta.stride(0), stream_ptr(X.device))
        if need_xi:
            dxi = (g[:, tm_dim:tm_dim + K] * ctx.dxi_out).sum(dim=0)
        return dX, dtheta, dxi, dbias, dscale, None, None, None


class FSWEmbedFunction(torch.autograd.Function):
    """out[S, d_out] = [ total-mass channel | (1+xi_k) sum_j p_(j) D_j ] + bias   (+ autograd).

    Forward: K1 projection(s) + K2 fused kernel.  Backward: K3 fused kernel + the three contractions
    dX = dXp.theta, dtheta = dXp^T.X, and the edge-feature analogues.
    Inputs that are None: E_feat, bias, tm_scale, W_vals.  `tm_function` None disables the total-mass channel
    ('plain' method only; the homogeneous variants are composed in torch by the caller).
    W_vals: the raw weights of the plan's elements (same values and order as plan.W) when their gradient is wanted,
    else None - the kernels read plan.W either way; the argument only makes the weights a differentiable input.
    """

    @staticmethod
    def forward(ctx, X, projVecs, freqs, bias, tm_scale, E_feat, W_vals, plan, tm_function, grad_mode=True):
        d = X.shape[1]
        K = projVecs.shape[0]
        tm_dim = 0 if tm_function is None else 1
        d_out = K + tm_dim
        X = X.contiguous()
        freqs = freqs.contiguous()
        # Multi-GPU (dist.RowExchange on the plan): X holds this rank's rows only.  The slices are cut into column
        # chunks; every chunk is projected locally, all-gathered asynchronously, and embedded as soon as its rows
        # have arrived - the exchange of chunk c+1 runs under the kernels of chunk c.  Single GPU: one chunk.
        exch = getattr(plan, "exchange", None)
        nchunk = exch.chunks if (exch is not None and E_feat is None) else 1
        bounds = column_chunks(K, nchunk)
        out = torch.empty((plan.S, d_out), dtype=X.dtype, device=X.device)
        bias_core = None
        if bias is not None:
            bias = bias.contiguous()
            bias_core = bias[tm_dim:]
        # under torch.no_grad() needs_input_grad is still True for parameters, and inside forward() grad mode is always
        # off: the caller's grad mode is sampled before apply() (fsw_embed) so that evaluation runs the inference
        # kernels and records nothing
        needs_grad = bool(grad_mode) and any(ctx.needs_input_grad[:7])
        Ep = None
        if E_feat is not None:
            E_feat = E_feat.contiguous()
        chunks = []
        for (k0, k1) in bounds:
            ldc = round_up(max(k1 - k0, 1), 8)
            xp = project(X, projVecs[k0:k1, :d], ldc)
            work = None
            if exch is not None:
                xp, work = exch.gather_async(xp)
            chunks.append([k0, k1, ldc, xp, work, None, None])
        for ch in chunks:
            k0, k1, ldc, xp, work = ch[:5]
            if work is not None:
                work.wait()   # the compute stream waits for this chunk's rows only
                ch[4] = None
            if E_feat is not None:
                Ep = project(E_feat, projVecs[k0:k1, d:], ldc)
            # training: record each element's sorted position per slice (uint16) so that the backward needs no sort.
            # 2 bytes per (element, slice); skipped when that would not fit comfortably (then the backward re-sorts).
            ranks = None
            if needs_grad and SAVE_RANKS and plan.E > 0 and k1 > k0:
                nbytes = plan.E * ldc * 2
                ok = nbytes < (256 << 20)  # small buffers never need the (slow) driver query
                if not ok:
                    free_b, _total = torch.cuda.mem_get_info(X.device)
                    # blocks cached by torch's allocator are free for us too
                    free_b += torch.cuda.memory_reserved(X.device) - torch.cuda.memory_allocated(X.device)
                    ok = nbytes < RANK_MEMORY_FRACTION * free_b
                if ok:
                    ranks = torch.empty((plan.E, ldc), dtype=torch.int16, device=X.device)
            # with learnable frequencies the forward also emits d out / d xi per (segment, slice)
            dxi_out = None
            if ranks is not None and ctx.needs_input_grad[2]:
                # the rank-recording forward writes every (segment, slice) of the uniform segments of up to 32768 elements;
                # only a plan with other segments needs the zero fill (1.9 GB per layer at configs[3])
                covered = X.dtype == torch.float32 and plan.uniform_fraction() == 1.0 and plan.max_n_eff <= RANKT_NMAX
                # (rows padded to the projected width: 32-byte aligned rows for the kernels that write and read it)
                dxi_out = (torch.empty if covered else torch.zeros)((plan.S, ldc), dtype=X.dtype, device=X.device)[:, :k1 - k0]
            embed_forward(plan, xp, ldc, Ep, freqs[k0:k1], out, d_out, tm_dim + k0,
                          None if bias_core is None else bias_core[k0:k1], ranks, dxi_out)
            ch[5], ch[6] = ranks, dxi_out
        ctx.chunks = [(k0, k1, ldc, xp, ranks, dxi) for (k0, k1, ldc, xp, _w, ranks, dxi) in chunks]
        fT = None
        if tm_dim:
            fT = total_mass_function(plan.mass_as(X.dtype), tm_function)
            col0 = fT * tm_scale
            if bias is not None:
                col0 = col0 + bias[0]
            out[:, 0] = col0
        ctx.plan, ctx.tm_dim = plan, tm_dim
        ctx.has_E, ctx.has_bias, ctx.has_scale = E_feat is not None, bias is not None, tm_scale is not None
        ctx.Ep = Ep   # only with edge features (then there is a single chunk)
        ctx.tm_function = tm_function
        ctx.save_for_backward(X, projVecs, freqs, E_feat, fT, tm_scale)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        X, projVecs, freqs, E_feat, fT, tm_scale = ctx.saved_tensors
        plan, tm_dim = ctx.plan, ctx.tm_dim
        need_X, need_theta, need_xi, need_bias, need_scale, need_E, need_W = ctx.needs_input_grad[:7]
        d = X.shape[1]
        K = projVecs.shape[0]
        if not (g.dim() == 2 and g.stride(1) == 1 and g.stride(0) >= g.shape[1]):
            g = g.contiguous()   # row-strided views (e.g. a column slice handed back by torch.cat) are read in place
        dX = dtheta = dxi = dbias = dscale = dE = dW = None
        if need_bias and ctx.has_bias:
            dbias = g.sum(dim=0)
        if need_scale and ctx.has_scale and tm_dim:
            dscale = (g[:, 0] * fT).sum()
        exch = getattr(plan, "exchange", None)
        # a rank that owns no destination rows still takes part in the exchange (its peers block in the collective)
        if (need_X or need_theta or need_xi or (need_E and ctx.has_E)) and K > 0 and (plan.S > 0 or exch is not None):
            n_local = X.shape[0]
            if need_xi:
                dxi = torch.empty(K, dtype=X.dtype, device=X.device)
            done = []
            for (k0, k1, ldc, Xp, ranks, dxi_out) in ctx.chunks:
                Nrows = Xp.shape[0]   # rows the segments index (all ranks' rows when the projected rows were exchanged)
                dxi_fwd = dxi_out is not None
                transpose = None
                if ranks is not None and X.dtype == torch.float32 and plan.col is not None and (dxi_fwd or not need_xi):
                    transpose = plan.transpose(Nrows)
                # graph mode accumulates dXp with atomics -> it must start from zero, UNLESS the source-major kernel runs: that
                # one writes every row (padding columns included) with plain stores before anything is added (1.9 GB of memset
                # per layer at configs[3] otherwise)
                if plan.col is not None and transpose is None:
                    dXp = torch.zeros((Nrows, ldc), dtype=X.dtype, device=X.device)
                else:
                    dXp = torch.empty((Nrows, ldc), dtype=X.dtype, device=X.device)
                if plan.col is None and ldc != k1 - k0:
                    dXp.zero_()
                dEp = torch.zeros((plan.E, ldc), dtype=X.dtype, device=X.device) if ctx.has_E else None
                dxi_acc = torch.zeros(k1 - k0, dtype=torch.float64, device=X.device) if need_xi else None
                embed_backward(plan, Xp, ldc, ctx.Ep, freqs[k0:k1], g, g.stride(0), tm_dim + k0, dXp, dEp, dxi_acc, ranks,
                               dxi_from_forward=(dxi_fwd or not need_xi), transpose=transpose, nrows=Nrows)
                if need_xi:
                    if dxi_fwd:   # + sum_s g[s, k] d out[s, k] / d xi_k of the segments the forward covered: one fused pass
                        gk = g[:, tm_dim + k0:tm_dim + k1]
                        _lib.call(X.device, "fsw_column_dot", dtype_code(X.dtype), ptr(gk), gk.stride(0), ptr(dxi_out), dxi_out.stride(0),
                                  plan.S, k1 - k0, ptr(dxi_acc), stream_ptr(X.device))
                    dxi[k0:k1] = dxi_acc.to(X.dtype)
                work = None
                if exch is not None:
                    # sum over ranks, keep this rank's rows: runs under the next chunk's kernels
                    dXp, work = exch.scatter_async(dXp, n_local)
                done.append((k0, k1, ldc, dXp, dEp, work))
            if need_theta:
                dtheta = torch.zeros_like(projVecs)
            for i, (k0, k1, ldc, dXp, dEp, work) in enumerate(done):
                if work is not None:
                    work.wait()
                    dXp = dXp[:n_local]
                kc = k1 - k0
                if need_X:
                    if i == 0:
                        dX = gemm(1, dXp, projVecs[k0:k1], n_local, d, kc, ldc, projVecs.stride(0))
                    else:
                        gemm(1, dXp, projVecs[k0:k1], n_local, d, kc, ldc, projVecs.stride(0), out=dX, ldc=dX.stride(0), accumulate=True)
                if need_theta:
                    gemm(2, dXp, X, kc, d, n_local, ldc, X.stride(0), out=dtheta[k0:k1], ldc=dtheta.stride(0), accumulate=True)
                if ctx.has_E:
                    de = E_feat.shape[1]
                    if need_E:
                        dE = gemm(1, dEp, projVecs[k0:k1, d:], plan.E, de, kc, ldc, projVecs.stride(0))
                    if need_theta:
                        gemm(2, dEp, E_feat, kc, de, plan.E, ldc, E_feat.stride(0), out=dtheta[k0:k1, d:], ldc=dtheta.stride(0),
                             accumulate=True)
        else:
            if need_X:
                dX = torch.zeros_like(X)
            if need_theta:
                dtheta = torch.zeros_like(projVecs)
            if need_xi:
                dxi = torch.zeros_like(freqs)
        if need_W:
            dW = weight_gradient(plan, ctx.chunks, ctx.Ep, freqs, g, tm_dim)
            if tm_dim:
                # total-mass channel f(T) * scale (fsw_embedding.py:857-868): d/dW_e = g[s, 0] * scale * f'(T_s)
                T = plan.mass_as(X.dtype)
                dchan = g[:, 0] * tm_scale * total_mass_function_derivative(T, ctx.tm_function)
                dW = dW + plan.expand_to_elements(dchan)
        return dX, dtheta, dxi, dbias, dscale, dE, dW, None, None, None


def weight_gradient(plan, chunks, Ep, freqs, g, tm_dim):
    """dL/dW (raw weights, [E]) of the core embedding: K3w accumulates the gradient w.r.t. the normalised weights over the
    column chunks, then one pass applies the normalisation chain (include/fsw_embedding.h section 6b)."""
    lib = _lib.load()
    dev, dtype = plan.device, plan.dtype
    dwn = torch.zeros(max(plan.E, 1), dtype=torch.float64, device=dev)
    dwn_pad = torch.zeros(max(plan.S, 1), dtype=torch.float64, device=dev)
    out = torch.zeros(plan.E, dtype=dtype, device=dev)
    if plan.E == 0 or plan.S == 0:
        return out
    max_n = plan.max_n_eff   # >= the largest element count (n_eff = n + pad)
    any_def = plan.any_deficient()
    scratch = _ws(lib.fsw_embed_weight_grad_scratch_bytes(dtype_code(dtype), max_n), dev)
    for (k0, k1, ldc, Xp, _ranks, _dxi) in chunks:
        if k1 <= k0:
            continue
        _lib.call(dev, "fsw_embed_backward_weights", dtype_code(dtype), ptr(Xp), ldc, ptr(Ep), ptr(plan.rowptr), plan.n_fixed,
                  ptr(plan.col), ptr(plan.W), ptr(plan.mass), plan.S, k1 - k0, ptr(freqs[k0:k1]), plan.thresh, 1 if any_def else 0,
                  ptr(g), g.stride(0), tm_dim + k0, ptr(dwn), ptr(dwn_pad), max_n, ptr(scratch), scratch.numel(), stream_ptr(dev))
    _lib.call(dev, "fsw_embed_weight_grad_finish", dtype_code(dtype), ptr(plan.rowptr), plan.n_fixed, ptr(plan.W), ptr(plan.mass),
              plan.S, plan.thresh, 1 if any_def else 0, ptr(dwn), ptr(dwn_pad), ptr(out), stream_ptr(dev))
    return out


def column_chunks(K, n):
    """[k0, k1) column ranges: n chunks of equal width rounded up to 8 columns (fewer when K is small)."""
    n = max(1, min(int(n), (K + 7) // 8 if K > 0 else 1))
    w = round_up(-(-K // n), 8) if K > 0 else 0
    out = []
    k0 = 0
    while k0 < K:
        out.append((k0, min(k0 + w, K)))
        k0 += w
    return out or [(0, 0)]


def fsw_embed(X, projVecs, freqs, bias, tm_scale, E_feat, plan, tm_function, W_vals=None):
    if W_vals is not None and not W_vals.requires_grad:
        W_vals = None
    if W_vals is None and cloud_eligible(X, plan, E_feat, projVecs.shape[0]):
        return FSWCloudFunction.apply(X, projVecs, freqs, bias, tm_scale, plan, tm_function, torch.is_grad_enabled())
    return FSWEmbedFunction.apply(X, projVecs, freqs, bias, tm_scale, E_feat, W_vals, plan, tm_function, torch.is_grad_enabled())


# ------------------------------------------------------------------------------------------------
# Segmented cumulative sum (public function of the reference, fsw_embedding.py:2795)
# ------------------------------------------------------------------------------------------------
def segcumsum_cuda(values, segment_ids, in_place=False):
    lib = _lib.load()
    _lib.require_cuda(values, "values")
    out = values if in_place else torch.empty_like(values, memory_format=torch.contiguous_format)
    vin = values if values.is_contiguous() else values.contiguous()
    n = vin.numel()
    if n == 0:
        return out
    ws = _ws(lib.fsw_segcumsum_workspace_bytes(n), values.device)
    _lib.call(values.device, "fsw_segcumsum", dtype_code(values.dtype), ptr(vin), ptr(out), ptr(segment_ids),
              segment_ids.element_size(), n, ptr(ws), ws.numel(), stream_ptr(values.device))
    return out
