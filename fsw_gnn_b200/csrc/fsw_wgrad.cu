// K3w: gradient of the embedding with respect to the WEIGHTS of the multisets.
//
// Reference path (fsw_embedding.py): the weights enter through the cumulative sums C_j of the sorted, normalised weights
// (ag.cumsum_sparse :2141-2172, backward = reverse cumulative sum :2160-2172), the normalisation W / max(T, thresh)
// (ag.div_sparse_dense :1656) and the deficit padding with its low clamp (:787-829, custom_lowclamp :1735-1744).
// With D_j = s(C_j) - s(C_{j-1}), s(C) = sin(2 pi xi C) / (pi xi) (the 'diff' form of :999-1003, equal to the product form
// :1047-1075), out_k = (1 + xi) sum_j p_(j) D_j gives
//     d out_k / d C_j = (1 + xi) 2 cos(2 pi xi C_j) (p_(j) - p_(j+1)),   p_(n_eff) := 0,
//     d out_k / d w_(j) = sum_{i >= j} d out_k / d C_i                   (reverse cumulative sum),
// summed over the slices with the upstream gradient, un-permuted to the element order, and finally pushed through the
// normalisation:
//     T <  thresh (padded, pad weight thresh - T, normaliser thresh):  dW_i = (dw_i - dw_pad) / thresh
//     T == thresh while some other segment is deficient (the reference then pads EVERY row, this one with weight 0 whose
//                 clamp is still 'active'):                             dW_i = dw_i / T - <dw, W> / T^2 - dw_pad / T
//     otherwise:                                                        dW_i = dw_i / T - <dw, W> / T^2
//
// This is the rare path (FSW_conv never differentiates its adjacency, fsw_conv.py:388-398): one CTA per tile
// [n_pad elements][32 slices] with the block-wide bitonic network, tiles in shared memory or, beyond 200 KB, in a global
// scratch slice per resident CTA.  Sums run in fp64.
#include "fsw_sortnet.cuh"

namespace {

const int kWgSmemBudget = 200 * 1024;
const int kWgGrid = 148 * 2;

template <typename T>
__host__ __device__ constexpr size_t wg_tile_bytes_per_row() {
    return 32 * (2 * sizeof(T) + sizeof(int));  // keys, aux (T) and element index (int) per lane
}

// pad_mode of a segment: 0 none, 1 deficit pad (weight thresh - T), 2 zero-weight pad that still carries a gradient
__device__ __forceinline__ int wg_pad_mode(double Ts, double thresh, int any_deficient) {
    if (Ts < thresh) return 1;
    if (any_deficient && Ts == thresh) return 2;
    return 0;
}

template <typename T>
__global__ void __launch_bounds__(256) fsw_wgrad_kernel(SegArgs<T> a, int64_t S, int nchunks, int64_t ntiles, int any_deficient,
                                                        const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                        double* __restrict__ dwn /*[E]*/, double* __restrict__ dwn_pad /*[S]*/,
                                                        int cap_rows, unsigned char* gscratch) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const size_t tile_elems = (size_t)cap_rows * 32;
    unsigned char* base = gscratch ? gscratch + (size_t)blockIdx.x * (size_t)cap_rows * wg_tile_bytes_per_row<T>() : fsw_smem_raw;
    T* keys = reinterpret_cast<T*>(base);
    T* aux = keys + tile_elems;
    int* idx = reinterpret_cast<int*>(aux + tile_elems);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int s = (int)(tile / nchunks);
        const int chunk = (int)(tile % nchunks);
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);
        if (n == 0) continue;  // block-uniform
        const double Ts = a.mass[s];
        const int pm = wg_pad_mode(Ts, a.thresh, any_deficient);
        const int n_eff = n + (pm != 0);
        if (n_eff > cap_rows) continue;  // cannot happen: the host sizes cap_rows from the largest segment
        const int n_pad = fsw_next_pow2(n_eff);
        const int k = chunk * 32 + lane;
        const bool act = k < a.K;
        const int kk = act ? k : a.K - 1;
        const double xid = (double)fsw_ldg(a.freqs + kk);
        const double gk = act ? (double)g[(int64_t)s * ld_g + g_col0 + k] : 0.0;
        const double G2 = 2.0 * gk * (1.0 + xid);
        const double invS = 1.0 / fmax(Ts, a.thresh);
        const double padw = (pm == 1) ? a.thresh - Ts : 0.0;

        // ---- gather (key, element index); the pad point sits at x = 0 ----
        for (int r = warp; r < n_pad; r += nw) {
            T v = Num<T>::big();
            if (r < n) {
                const int64_t row = a.col ? (int64_t)a.col[e0 + r] : e0 + r;
                v = fsw_ldg(a.Xp + row * a.ldp + kk);
                if (a.Ep) v += fsw_ldg(a.Ep + (e0 + r) * a.ldp + kk);
            } else if (r == n && pm != 0) {
                v = (T)0;
            }
            keys[r * 32 + lane] = v;
            idx[r * 32 + lane] = r;
        }
        __syncthreads();
        fsw_block_bitonic<T, int, true>(keys, idx, n_pad);

        // ---- cumulative normalised weights C_j (prefix over the warps' contiguous ranges), d/dC_j into aux ----
        const int rpw = (n_eff + nw - 1) / nw;
        const int r0 = min(warp * rpw, n_eff);
        const int r1 = min(r0 + rpw, n_eff);
        double part = 0.0;
        for (int r = r0; r < r1; ++r) {
            const int id = idx[r * 32 + lane];
            part += (id >= n) ? padw : (a.W ? (double)a.W[e0 + id] : 1.0);
        }
        red[warp][lane] = part;
        __syncthreads();
        double Craw = 0.0;
        for (int w2 = 0; w2 < warp; ++w2) Craw += red[w2][lane];
        __syncthreads();
        double tail = 0.0;  // sum of d/dC over this warp's range
        for (int r = r0; r < r1; ++r) {
            const int id = idx[r * 32 + lane];
            Craw += (id >= n) ? padw : (a.W ? (double)a.W[e0 + id] : 1.0);
            const double C = Craw * invS;
            const double ph = 2.0 * xid * C;                               // phase in units of pi
            const double c = cospi(ph - 2.0 * rint(0.5 * ph));
            const double p = (double)keys[r * 32 + lane];
            const double pn = (r + 1 < n_eff) ? (double)keys[(r + 1) * 32 + lane] : 0.0;
            const double dC = G2 * c * (p - pn);
            aux[r * 32 + lane] = (T)dC;
            tail += dC;
        }
        red[warp][lane] = tail;
        __syncthreads();
        // ---- reverse cumulative sum: d/dw_(j) = sum_{i >= j} d/dC_i; result un-permuted into keys[element][lane] ----
        double suffix = 0.0;
        for (int w2 = warp + 1; w2 < nw; ++w2) suffix += red[w2][lane];
        __syncthreads();  // everyone has read keys[r + 1] and red[]
        for (int r = r1 - 1; r >= r0; --r) {
            suffix += (double)aux[r * 32 + lane];
            const int id = idx[r * 32 + lane];
            keys[id * 32 + lane] = (T)suffix;
        }
        __syncthreads();
        // ---- sum over the 32 slices of the tile, one atomic per element ----
        for (int i = warp; i < n_eff; i += nw) {
            double v = act ? (double)keys[i * 32 + lane] : 0.0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FSW_FULL, v, m);
            if (lane == 0) {
                if (i < n)
                    atomicAdd(dwn + e0 + i, v);
                else
                    atomicAdd(dwn_pad + s, v);
            }
        }
        __syncthreads();
    }
}

// one warp per segment: normalisation chain (see the file header)
template <typename T>
__global__ void __launch_bounds__(256) fsw_wgrad_finish_kernel(const int32_t* __restrict__ rowptr, int64_t n_fixed,
                                                               const T* __restrict__ W, const double* __restrict__ mass, int64_t S,
                                                               double thresh, int any_deficient, const double* __restrict__ dwn,
                                                               const double* __restrict__ dwn_pad, T* __restrict__ dW) {
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= S) return;
    int64_t e0;
    int n;
    if (rowptr) {
        e0 = rowptr[s];
        n = rowptr[s + 1] - (int)e0;
    } else {
        e0 = s * n_fixed;
        n = (int)n_fixed;
    }
    if (n == 0) return;
    const double Ts = mass[s];
    const int pm = wg_pad_mode(Ts, thresh, any_deficient);
    if (pm == 1) {
        const double dp = dwn_pad[s];
        for (int i = lane; i < n; i += 32) dW[e0 + i] = (T)((dwn[e0 + i] - dp) / thresh);
        return;
    }
    double dot = 0.0;
    for (int i = lane; i < n; i += 32) dot += dwn[e0 + i] * (W ? (double)W[e0 + i] : 1.0);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) dot += __shfl_xor_sync(FSW_FULL, dot, m);
    const double invT = 1.0 / Ts;
    const double shift = dot * invT * invT + (pm == 2 ? dwn_pad[s] * invT : 0.0);
    for (int i = lane; i < n; i += 32) dW[e0 + i] = (T)(dwn[e0 + i] * invT - shift);
}

template <typename T>
int wgrad_t(const SegArgs<T>& a, int64_t S, int any_deficient, const T* g, int64_t ld_g, int64_t g_col0, double* dwn, double* dwn_pad,
            int64_t max_n, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int64_t ntiles = S * nchunks;
    int cap = 2;
    while (cap < max_n + 1) cap <<= 1;
    const size_t tb = (size_t)cap * wg_tile_bytes_per_row<T>();
    unsigned grid = (unsigned)(ntiles < kWgGrid ? ntiles : kWgGrid);
    unsigned char* gs = nullptr;
    size_t smem = tb;
    if (tb > (size_t)kWgSmemBudget) {
        if ((size_t)grid * tb > scratch_bytes) {
            grid = (unsigned)(scratch_bytes / tb);
            if (grid == 0) return fsw_fail(FSW_ERR_WORKSPACE, "weight-gradient scratch too small: need >= %zu bytes", tb);
        }
        gs = (unsigned char*)scratch;
        smem = 0;
    }
    auto kern = fsw_wgrad_kernel<T>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fsw_prof_begin(sizeof(T) == 4 ? "bwd_wgrad_f32" : "bwd_wgrad_f64", st);
    kern<<<grid, 256, smem, st>>>(a, S, nchunks, ntiles, any_deficient, g, ld_g, g_col0, dwn, dwn_pad, cap, gs);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_wgrad_kernel");
    return FSW_OK;
}

}  // namespace

extern "C" size_t fsw_embed_weight_grad_scratch_bytes(int dtype, int64_t max_n) {
    int64_t cap = 2;
    while (cap < max_n + 1) cap <<= 1;
    const size_t per_row = dtype == FSW_F64 ? wg_tile_bytes_per_row<double>() : wg_tile_bytes_per_row<float>();
    const size_t tb = (size_t)cap * per_row;
    if (tb <= (size_t)kWgSmemBudget) return 0;
    size_t want = (size_t)kWgGrid * tb;
    const size_t limit = (size_t)2 << 30;  // never more than 2 GiB: fewer resident CTAs instead
    if (want > limit) want = (limit / tb ? limit / tb : 1) * tb;
    return want;
}

extern "C" int fsw_embed_backward_weights(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr,
                                          int64_t n_fixed, const int32_t* col, const void* W, const double* mass, int64_t S,
                                          int64_t K, const void* freqs, double thresh, int any_deficient, const void* g,
                                          int64_t ld_g, int64_t g_col0, double* dwn_acc, double* dwn_pad_acc, int64_t max_n,
                                          void* scratch, size_t scratch_bytes, void* stream) {
    if (S == 0 || K == 0) return FSW_OK;
    if (!Xp || !mass || !freqs || !g || !dwn_acc || !dwn_pad_acc)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_weights: null argument");
    if (!rowptr && n_fixed <= 0) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_weights: rowptr == NULL needs n_fixed > 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FSW_F32) {
        SegArgs<float> a{(const float*)Xp, (const float*)Ep, rowptr, col, (const float*)W, mass, nullptr, nullptr, (const float*)freqs,
                         ldp, n_fixed, (int)K, thresh};
        return wgrad_t<float>(a, S, any_deficient, (const float*)g, ld_g, g_col0, dwn_acc, dwn_pad_acc, max_n, scratch, scratch_bytes, st);
    } else if (dtype == FSW_F64) {
        SegArgs<double> a{(const double*)Xp, (const double*)Ep, rowptr, col, (const double*)W, mass, nullptr, nullptr,
                          (const double*)freqs, ldp, n_fixed, (int)K, thresh};
        return wgrad_t<double>(a, S, any_deficient, (const double*)g, ld_g, g_col0, dwn_acc, dwn_pad_acc, max_n, scratch, scratch_bytes, st);
    }
    return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_weights: dtype %d", dtype);
}

extern "C" int fsw_embed_weight_grad_finish(int dtype, const int32_t* rowptr, int64_t n_fixed, const void* W, const double* mass,
                                            int64_t S, double thresh, int any_deficient, const double* dwn_acc,
                                            const double* dwn_pad_acc, void* dW, void* stream) {
    if (S == 0) return FSW_OK;
    if (!mass || !dwn_acc || !dwn_pad_acc || !dW) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_weight_grad_finish: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)fsw_cdiv(S, 8);
    if (dtype == FSW_F32)
        fsw_wgrad_finish_kernel<float><<<blocks, 256, 0, st>>>(rowptr, n_fixed, (const float*)W, mass, S, thresh, any_deficient, dwn_acc,
                                                               dwn_pad_acc, (float*)dW);
    else if (dtype == FSW_F64)
        fsw_wgrad_finish_kernel<double><<<blocks, 256, 0, st>>>(rowptr, n_fixed, (const double*)W, mass, S, thresh, any_deficient,
                                                                dwn_acc, dwn_pad_acc, (double*)dW);
    else
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_weight_grad_finish: dtype %d", dtype);
    FSW_CHECK_LAUNCH("fsw_wgrad_finish_kernel");
    return FSW_OK;
}

// ---------------------------------------------------------------------------------------------------
// dL/dxi of the segments whose d out / d xi the forward produced: acc[k] += sum_s g[s, k] dxi_out[s, k]  (float64 accumulators).
// One pass over the two [S, K] matrices (2 x 1.9 GB at configs[3]) instead of an elementwise product written back to memory and
// a column reduction over it.  A warp takes 64 rows (lanes = columns), fp32 partial sums of 64 terms, then the 8 warps of a block
// meet in shared memory and one float64 atomic per column and block lands in `acc`.
// ---------------------------------------------------------------------------------------------------
namespace {
template <typename T>
__global__ void __launch_bounds__(256) fsw_column_dot_kernel(const T* __restrict__ g, int64_t ld_g, const T* __restrict__ d, int64_t ld_d,
                                                             int64_t S, int K, double* __restrict__ acc) {
    constexpr int ROWS = 64;   // rows per warp
    constexpr int CB = 8;      // column chunks of 32 held per lane: a warp reads whole rows of up to 256 columns, contiguously
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * ROWS;
    const int nr = (int)((r0 + ROWS < S ? r0 + ROWS : S) - r0);   // may be <= 0 for the last warps (they still join the barriers)
    for (int kb = 0; kb < K; kb += 32 * CB) {   // block-uniform
        T part[CB];
#pragma unroll
        for (int i = 0; i < CB; ++i) part[i] = (T)0;
        const T* gp = g + r0 * ld_g + kb + lane;
        const T* dp = d + r0 * ld_d + kb + lane;
#pragma unroll 2
        for (int r = 0; r < nr; ++r) {
#pragma unroll
            for (int i = 0; i < CB; ++i)
                if (kb + 32 * i + lane < K) part[i] = fma(__ldg(gp + (int64_t)r * ld_g + 32 * i), __ldg(dp + (int64_t)r * ld_d + 32 * i), part[i]);
        }
#pragma unroll
        for (int i = 0; i < CB; ++i) {
            if (kb + 32 * i >= K) break;   // block-uniform
            red[warp][lane] = (double)part[i];
            __syncthreads();
            const int k = kb + 32 * i + lane;
            if (warp == 0 && k < K) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += red[w][lane];
                atomicAdd(acc + k, t);
            }
            __syncthreads();
        }
    }
}
}  // namespace

extern "C" int fsw_column_dot(int dtype, const void* g, int64_t ld_g, const void* d, int64_t ld_d, int64_t S, int64_t K, double* acc,
                              void* stream) {
    if (S == 0 || K == 0) return FSW_OK;
    if (!g || !d || !acc) return fsw_fail(FSW_ERR_INVALID, "fsw_column_dot: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)fsw_cdiv(S, 8 * 64);
    fsw_prof_begin("bwd_dxi_dot", st);
    if (dtype == FSW_F32)
        fsw_column_dot_kernel<float><<<blocks, 256, 0, st>>>((const float*)g, ld_g, (const float*)d, ld_d, S, (int)K, acc);
    else if (dtype == FSW_F64)
        fsw_column_dot_kernel<double><<<blocks, 256, 0, st>>>((const double*)g, ld_g, (const double*)d, ld_d, S, (int)K, acc);
    else
        return fsw_fail(FSW_ERR_INVALID, "fsw_column_dot: dtype %d", dtype);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_column_dot_kernel");
    return FSW_OK;
}
