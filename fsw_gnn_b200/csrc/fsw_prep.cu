// K0: graph preparation (edge list -> destination-major CSR) and the segment plan.
//
// Replaces FSW_conv.edge_index_to_adj (fsw_conv.py:384-447: sparse_coo_tensor + coalesce +
// get_slice_info + sum_sparseToDense), sp.get_slice_info (fsw_embedding.py:2586-2678) and the
// total-mass prologue of FSW_embedding.forward (fsw_embedding.py:778-829).
#include <cub/device/device_radix_sort.cuh>

#include "fsw_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// int32 exclusive scan (three small kernels; N+1 <= 2^31)
// ---------------------------------------------------------------------------------------------------
const int SCAN_BLOCK = 1024;  // elements per block (256 threads x 4)

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-level exclusive scan of 1024 ints held 4 per thread; returns the block total
__device__ __forceinline__ int block_excl_scan4(int (&x)[4], int* smem_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int tsum = x[0] + x[1] + x[2] + x[3];
    int incl = warp_incl_scan(tsum, lane);
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (int)(blockDim.x >> 5)) ? smem_warp[lane] : 0;
        int wi = warp_incl_scan(w, lane);
        smem_warp[lane] = wi - w;       // exclusive prefix of warps
        if (lane == 31) smem_warp[32] = wi;  // total
    }
    __syncthreads();
    int run = smem_warp[warp] + incl - tsum;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int t = x[i];
        x[i] = run;
        run += t;
    }
    return smem_warp[32];
}

__global__ void __launch_bounds__(256) scan_block_sums(const int* __restrict__ in, int64_t n, int* __restrict__ block_sums) {
    __shared__ int sw[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    int x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (base + i < n) ? in[base + i] : 0;
    int tot = block_excl_scan4(x, sw);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_of_block_sums(int* __restrict__ block_sums, int nblocks) {
    // single block, sequential over chunks of 1024
    __shared__ int sw[33];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < nblocks) ? block_sums[i] : 0;
        int incl = warp_incl_scan(v, lane);
        if (lane == 31) sw[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = sw[lane];
            int wi = warp_incl_scan(w, lane);
            sw[lane] = wi - w;
            if (lane == 31) sw[32] = wi;
        }
        __syncthreads();
        int excl = carry + sw[warp] + incl - v;
        if (i < nblocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) carry += sw[32];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) scan_apply(const int* __restrict__ in, int64_t n, const int* __restrict__ block_offs,
                                                  int* __restrict__ out) {
    __shared__ int sw[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    int x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (base + i < n) ? in[base + i] : 0;
    block_excl_scan4(x, sw);
    const int off = block_offs[blockIdx.x];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) out[base + i] = x[i] + off;
}

int exclusive_scan_i32(const int* in, int64_t n, int* out, int* block_tmp, cudaStream_t st) {
    const int nblocks = (int)fsw_cdiv(n, SCAN_BLOCK);
    scan_block_sums<<<nblocks, 256, 0, st>>>(in, n, block_tmp);
    FSW_CHECK_LAUNCH("scan_block_sums");
    scan_of_block_sums<<<1, 1024, 0, st>>>(block_tmp, nblocks);
    FSW_CHECK_LAUNCH("scan_of_block_sums");
    scan_apply<<<nblocks, 256, 0, st>>>(in, n, block_tmp, out);
    FSW_CHECK_LAUNCH("scan_apply");
    return FSW_OK;
}

// ---------------------------------------------------------------------------------------------------
// CSR from edge_index
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csr_count(const int64_t* __restrict__ dst, int64_t E, int64_t N, int self_loops,
                                                 int* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) {
        const int64_t d = dst[i];
        if (d >= 0 && d < N) atomicAdd(counts + d, 1);
    } else if (self_loops && i - E < N) {
        atomicAdd(counts + (i - E), 1);
    }
}

// Stable placement of the elements: the (destination, element id) pairs are radix-sorted by destination (cub, stable), so the
// elements of a segment keep the order of the edge list - the same graph always gives the same CSR, whatever the scheduling
// (a cursor scatter with returning atomics does not), and it is faster (1 GB of 8-byte pairs, 3 passes over 22 key bits).
__global__ void __launch_bounds__(256) csr_keys(const int64_t* __restrict__ dst, int64_t E, int64_t N, int self_loops,
                                                int32_t* __restrict__ keys, int32_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tot = E + (self_loops ? N : 0);
    if (i >= tot) return;
    int64_t d = (i < E) ? dst[i] : i - E;
    if (d < 0 || d >= N) d = N;   // out of range (callers validate): behind every real segment, outside rowptr[N]
    keys[i] = (int32_t)d;
    ids[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256) csr_gather_cols(const int64_t* __restrict__ src, int64_t E, int64_t cnt,
                                                       const int32_t* __restrict__ eid, int32_t* __restrict__ col) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt) return;
    const int32_t i = eid[p];
    col[p] = (i < E) ? (int32_t)src[i] : (int32_t)(i - E);
}

__global__ void __launch_bounds__(256) iota_kernel(int64_t n, int32_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = (int32_t)i;
}

int key_bits(int64_t max_key) {
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) <= max_key) ++b;
    return b;
}

size_t sort_temp_bytes(int64_t n) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (int)(n > 0 ? n : 1));
    return (tb + 255) & ~(size_t)255;
}

__global__ void __launch_bounds__(256) rowptr_from_rows(const int64_t* __restrict__ rows, int64_t nnz, int64_t S,
                                                        int32_t* __restrict__ rowptr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (nnz == 0) {
        if (i <= S) rowptr[i] = 0;
        return;
    }
    if (i < nnz) {
        const int64_t r = rows[i];
        const int64_t prev = (i > 0) ? rows[i - 1] : -1;
        for (int64_t x = prev + 1; x <= r; ++x) rowptr[x] = (int32_t)i;
        if (i == nnz - 1)
            for (int64_t x = r + 1; x <= S; ++x) rowptr[x] = (int32_t)nnz;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) edge_weights_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                           const int32_t* __restrict__ eid, int64_t N, int64_t E, int self_loops,
                                                           double slw, int gcn, T* __restrict__ deg, T* __restrict__ w,
                                                           int phase) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 0) {
        if (i < N) {
            const int cnt = rowptr[i + 1] - rowptr[i];
            const double d = self_loops ? (double)(cnt - 1) + slw : (double)cnt;
            deg[i] = (T)d;
        }
        return;
    }
    // phase 1: one thread per destination row (rows are short); writes w for its slots
    if (i < N && w) {
        const int lo = rowptr[i], hi = rowptr[i + 1];
        const double dd = (double)deg[i];
        for (int p = lo; p < hi; ++p) {
            const bool is_loop = self_loops && eid[p] >= E;
            double b = is_loop ? slw : 1.0;
            if (gcn) b = b / sqrt(dd) / sqrt((double)deg[col[p]]);
            w[p] = (T)b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Segment plan
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) plan_stats(const int32_t* __restrict__ rowptr, int64_t n_fixed, const T* __restrict__ W,
                                                  int64_t S, double thresh, double* __restrict__ mass, int32_t* __restrict__ info,
                                                  int* __restrict__ hist /*[BUCKETS+1] last = max n_eff*/,
                                                  unsigned long long* __restrict__ elems /*[BUCKETS] or null*/) {
    const int lane = threadIdx.x & 31;
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s < S) {
        int64_t e0;
        int n;
        if (rowptr) {
            e0 = rowptr[s];
            n = rowptr[s + 1] - (int)e0;
        } else {
            e0 = s * n_fixed;
            n = (int)n_fixed;
        }
        double m = 0.0;
        bool uni = true;
        if (W) {
            const T w_first = (n > 0) ? W[e0] : (T)0;
            for (int j = lane; j < n; j += 32) {
                const T w = W[e0 + j];
                m += (double)w;
                uni = uni && (w == w_first);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m += __shfl_xor_sync(0xffffffffu, m, d);
            uni = __all_sync(0xffffffffu, uni);
        } else {
            m = (double)n;
        }
        const bool deficient = m < thresh;
        const int n_eff = n + (deficient ? 1 : 0);
        const bool uniform = uni && !deficient && n > 0;
        if (lane == 0) {
            mass[s] = m;
            info[s] = n_eff | (uniform ? FSW_INFO_UNIFORM : 0);
            const int b = (uniform ? 0 : FSW_PLAN_BUCKETS_PER_KIND) + fsw_size_bucket(n_eff);
            atomicAdd(hist + b, 1);
            atomicMax(hist + FSW_PLAN_BUCKETS, n_eff);
            if (elems) atomicAdd(elems + b, (unsigned long long)n_eff);
        }
    }
}

__global__ void plan_scan(const int* __restrict__ hist, int32_t* __restrict__ offsets, int* __restrict__ cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < FSW_PLAN_BUCKETS; ++i) {
            offsets[i] = run;
            cursor[i] = run;
            run += hist[i];
        }
        offsets[FSW_PLAN_BUCKETS] = run;
        offsets[FSW_PLAN_BUCKETS + 1] = hist[FSW_PLAN_BUCKETS];  // max n_eff
    }
}

// Stable-enough scatter: a block reserves one contiguous range per bucket, then hands out slots in
// thread order, so the order inside a bucket follows the segment index up to block scheduling.
__global__ void __launch_bounds__(256) plan_scatter(const int32_t* __restrict__ info, int64_t S, int* __restrict__ cursor,
                                                    int32_t* __restrict__ order) {
    __shared__ int cnt[FSW_PLAN_BUCKETS];
    __shared__ int base[FSW_PLAN_BUCKETS];
    for (int i = threadIdx.x; i < FSW_PLAN_BUCKETS; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = -1, local = 0;
    if (s < S) {
        const int w = info[s];
        b = ((w & FSW_INFO_UNIFORM) ? 0 : FSW_PLAN_BUCKETS_PER_KIND) + fsw_size_bucket(w & FSW_INFO_NMASK);
        local = atomicAdd(&cnt[b], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FSW_PLAN_BUCKETS; i += blockDim.x)
        if (cnt[i]) base[i] = atomicAdd(cursor + i, cnt[i]);
    __syncthreads();
    if (b >= 0) order[base[b] + local] = (int32_t)s;
}

// ---------------------------------------------------------------------------------------------------
// Transpose of the segment structure: for every point row j the list of (segment, slot) pairs that
// reference it.  Used by the source-major backward (fsw_embed_small.cu) which replaces the scatter of
// atomics into dXp by one register accumulation + one plain store per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_count(const int32_t* __restrict__ col, int64_t E, int* __restrict__ counts) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) atomicAdd(counts + col[e], 1);
}

// segment of every element (elements are stored segment by segment) and its eligibility
__global__ void __launch_bounds__(256) transpose_segids(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ info, int64_t S,
                                                        int nmax, int32_t* __restrict__ segid, int32_t* __restrict__ elig_of_elem) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= S) return;
    const int lo = rowptr[s], hi = rowptr[s + 1];
    const int w = info[s];
    const int n = hi - lo;
    const int elig = ((w & FSW_INFO_UNIFORM) && n <= nmax) ? n : 0;
    for (int e = lo + lane; e < hi; e += 32) {
        segid[e] = (int32_t)s;
        elig_of_elem[e] = elig;
    }
}

// tslot (elements sorted by source row, stable: in element order) -> segment and eligible size of every pair
__global__ void __launch_bounds__(256) transpose_lookup(const int32_t* __restrict__ tslot, const int32_t* __restrict__ segid,
                                                        const int32_t* __restrict__ elig_of_elem, int64_t E, int32_t* __restrict__ tseg,
                                                        int32_t* __restrict__ tn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E) return;
    const int32_t e = tslot[t];
    tseg[t] = segid[e];
    tn[t] = elig_of_elem[e];
}

// ---------------------------------------------------------------------------------------------------
// Coalescing CSR (graphs with edge features, fsw_conv.py:397-398, :438-439): duplicate (dst, src) pairs become ONE element that
// carries the sum of their base weights; every input edge learns the slot it was merged into (the caller sums the edge
// features per slot).  Stable radix sort of 64-bit keys dst * N + src, head flags, scan.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) coal_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int64_t N,
                                                 int self_loops, unsigned long long* __restrict__ keys, int32_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tot = E + (self_loops ? N : 0);
    if (i >= tot) return;
    const int64_t d = (i < E) ? dst[i] : i - E, c = (i < E) ? src[i] : i - E;
    keys[i] = (d >= 0 && d < N && c >= 0 && c < N) ? (unsigned long long)d * (unsigned long long)N + (unsigned long long)c
                                                  : (unsigned long long)N * (unsigned long long)N;   // invalid: behind everything
    ids[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256) coal_heads(const unsigned long long* __restrict__ keys, int64_t tot, int64_t N, int* __restrict__ head) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= tot) return;
    const unsigned long long k = keys[p];
    const bool valid = k < (unsigned long long)N * (unsigned long long)N;
    head[p] = (valid && (p == 0 || keys[p - 1] != k)) ? 1 : 0;
}

// one thread per sorted position: the slot of its input edge; the head of a run writes the element (column, summed base weight,
// destination count).  Runs are short (duplicates are rare), so the head walks its run - a deterministic sum.
template <typename T>
__global__ void __launch_bounds__(256) coal_emit(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ ids,
                                                 const int* __restrict__ head, const int* __restrict__ excl, int64_t tot, int64_t E,
                                                 int64_t N, double slw, int32_t* __restrict__ col, T* __restrict__ W,
                                                 int32_t* __restrict__ slot_of_elem, int* __restrict__ counts, int32_t* __restrict__ nslots) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= tot) return;
    const unsigned long long k = keys[p];
    const bool valid = k < (unsigned long long)N * (unsigned long long)N;
    const int slot = excl[p] + head[p] - 1;   // exclusive scan of the heads + own head - 1 = index of the run this position is in
    slot_of_elem[ids[p]] = valid ? slot : -1;
    if (head[p]) {
        double w = 0.0;
        for (int64_t q = p; q < tot && keys[q] == k; ++q) w += (ids[q] < E) ? 1.0 : slw;
        col[slot] = (int32_t)(k % (unsigned long long)N);
        W[slot] = (T)w;
        atomicAdd(counts + (int64_t)(k / (unsigned long long)N), 1);
    }
    if (p == tot - 1) *nslots = excl[p] + head[p];
}

// in-degrees (sum of the element weights of a row) and, with gcn weighting, w / sqrt(deg[dst]) / sqrt(deg[src])
template <typename T>
__global__ void __launch_bounds__(256) coal_weights(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t N,
                                                    T* __restrict__ deg, T* __restrict__ W, int phase) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    const int lo = rowptr[v], hi = rowptr[v + 1];
    if (phase == 0) {
        double d = 0.0;
        for (int p = lo; p < hi; ++p) d += (double)W[p];
        deg[v] = (T)d;
    } else {
        const double dv = (double)deg[v];
        for (int p = lo; p < hi; ++p) W[p] = (T)((double)W[p] / sqrt(dv) / sqrt((double)deg[col[p]]));
    }
}

size_t sort64_temp_bytes(int64_t n) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)(n > 0 ? n : 1));
    return (tb + 255) & ~(size_t)255;
}

}  // namespace

extern "C" size_t fsw_transpose_workspace_bytes(int64_t Nrows, int64_t E) {
    // counts [Nrows+1] + block sums | ids, sorted keys, segment ids, eligibility [E] each | cub temporary storage
    const size_t head = ((size_t)((Nrows + 1) + fsw_cdiv(Nrows + 1, SCAN_BLOCK) + 64) * sizeof(int) + 255) & ~(size_t)255;
    const size_t arr = ((size_t)(E > 0 ? E : 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    return head + 4 * arr + sort_temp_bytes(E);
}

extern "C" int fsw_csr_transpose(const int32_t* rowptr, const int32_t* col, const int32_t* info, int64_t S, int64_t Nrows,
                                 int64_t E, int nmax_eligible, int32_t* tptr, int32_t* tseg, int32_t* tslot, int32_t* tn,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    if (!rowptr || !col || !info || !tptr) return fsw_fail(FSW_ERR_INVALID, "fsw_csr_transpose: null argument");
    if (workspace_bytes < fsw_transpose_workspace_bytes(Nrows, E)) return fsw_fail(FSW_ERR_WORKSPACE, "fsw_csr_transpose: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t head = ((size_t)((Nrows + 1) + fsw_cdiv(Nrows + 1, SCAN_BLOCK) + 64) * sizeof(int) + 255) & ~(size_t)255;
    const size_t arr = ((size_t)(E > 0 ? E : 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    unsigned char* wsb = (unsigned char*)workspace;
    int* counts = (int*)wsb;
    int* btmp = counts + (Nrows + 1);
    int32_t* ids = (int32_t*)(wsb + head);
    int32_t* keys_out = (int32_t*)(wsb + head + arr);
    int32_t* segid = (int32_t*)(wsb + head + 2 * arr);
    int32_t* elig = (int32_t*)(wsb + head + 3 * arr);
    void* temp = wsb + head + 4 * arr;
    FSW_CUDA(cudaMemsetAsync(counts, 0, (size_t)(Nrows + 1) * sizeof(int), st));
    if (E > 0) {
        transpose_count<<<(unsigned)fsw_cdiv(E, 256), 256, 0, st>>>(col, E, counts);
        FSW_CHECK_LAUNCH("transpose_count");
    }
    int rc = exclusive_scan_i32(counts, Nrows + 1, tptr, btmp, st);
    if (rc) return rc;
    if (E > 0 && S > 0) {
        // pairs of a source row in element order (stable sort of the element ids by source row): the source-major backward then
        // adds them in the same order in every run - bitwise reproducible gradients - and the scatter needs no atomics
        iota_kernel<<<(unsigned)fsw_cdiv(E, 256), 256, 0, st>>>(E, ids);
        FSW_CHECK_LAUNCH("iota_kernel");
        size_t tb = sort_temp_bytes(E);
        FSW_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, col, keys_out, ids, tslot, (int)E, 0, key_bits(Nrows), st));
        fsw_count_launch(3);
        transpose_segids<<<(unsigned)fsw_cdiv(S * 32, 256), 256, 0, st>>>(rowptr, info, S, nmax_eligible, segid, elig);
        FSW_CHECK_LAUNCH("transpose_segids");
        transpose_lookup<<<(unsigned)fsw_cdiv(E, 256), 256, 0, st>>>(tslot, segid, elig, E, tseg, tn);
        FSW_CHECK_LAUNCH("transpose_lookup");
    }
    return FSW_OK;
}

extern "C" size_t fsw_csr_coalesce_workspace_bytes(int64_t N, int64_t E) {
    // counts [N+1] + block sums | keys, sorted keys (8 B) | ids, sorted ids, heads, scanned heads (4 B) [E + N] each | cub temp
    const int64_t tot = E + N;
    const size_t head = ((size_t)((N + 1) + fsw_cdiv(tot + 1, SCAN_BLOCK) + fsw_cdiv(N + 1, SCAN_BLOCK) + 128) * sizeof(int) + 255) & ~(size_t)255;
    const size_t a4 = ((size_t)(tot > 0 ? tot : 1) * 4 + 255) & ~(size_t)255;
    const size_t a8 = ((size_t)(tot > 0 ? tot : 1) * 8 + 255) & ~(size_t)255;
    return head + 2 * a8 + 4 * a4 + sort64_temp_bytes(tot);
}

// edge_index [2, E] int64 -> coalesced destination-major CSR: rowptr [N+1], col [cap], W [cap] (`dtype`; cap >= E (+ N with self
// loops): the first *nslots entries are written), slot_of_elem [E (+ N)] = CSR slot of every input edge (then of every self
// loop), deg [N] in-degrees, nslots (device int32).  gcn != 0: W = base / sqrt(deg[dst]) / sqrt(deg[src]).
extern "C" int fsw_csr_coalesce(int dtype, const int64_t* edge_index, int64_t E, int64_t N, int self_loops, double self_loop_weight,
                                int gcn, int32_t* rowptr, int32_t* col, void* W, int32_t* slot_of_elem, void* deg, int32_t* nslots,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (N < 0 || E < 0) return fsw_fail(FSW_ERR_INVALID, "fsw_csr_coalesce: negative size");
    if (E + N >= (int64_t)INT32_MAX) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_csr_coalesce: more than 2^31 elements");
    if (N >= ((int64_t)1 << 31)) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_csr_coalesce: more than 2^31 vertices");
    if (!rowptr || !col || !W || !slot_of_elem || !deg || !nslots) return fsw_fail(FSW_ERR_INVALID, "fsw_csr_coalesce: null argument");
    if (dtype != FSW_F32 && dtype != FSW_F64) return fsw_fail(FSW_ERR_INVALID, "fsw_csr_coalesce: dtype %d", dtype);
    if (workspace_bytes < fsw_csr_coalesce_workspace_bytes(N, E)) return fsw_fail(FSW_ERR_WORKSPACE, "fsw_csr_coalesce: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t totmax = E + N, tot = E + (self_loops ? N : 0);
    const size_t head_b = ((size_t)((N + 1) + fsw_cdiv(totmax + 1, SCAN_BLOCK) + fsw_cdiv(N + 1, SCAN_BLOCK) + 128) * sizeof(int) + 255) & ~(size_t)255;
    const size_t a4 = ((size_t)(totmax > 0 ? totmax : 1) * 4 + 255) & ~(size_t)255;
    const size_t a8 = ((size_t)(totmax > 0 ? totmax : 1) * 8 + 255) & ~(size_t)255;
    unsigned char* wsb = (unsigned char*)workspace;
    int* counts = (int*)wsb;
    int* btmp = counts + (N + 1);
    unsigned long long* keys = (unsigned long long*)(wsb + head_b);
    unsigned long long* keys_s = (unsigned long long*)(wsb + head_b + a8);
    int32_t* ids = (int32_t*)(wsb + head_b + 2 * a8);
    int32_t* ids_s = (int32_t*)(wsb + head_b + 2 * a8 + a4);
    int* heads = (int*)(wsb + head_b + 2 * a8 + 2 * a4);
    int* excl = (int*)(wsb + head_b + 2 * a8 + 3 * a4);
    void* temp = wsb + head_b + 2 * a8 + 4 * a4;
    FSW_CUDA(cudaMemsetAsync(counts, 0, (size_t)(N + 1) * sizeof(int), st));
    FSW_CUDA(cudaMemsetAsync(nslots, 0, sizeof(int32_t), st));
    if (tot > 0) {
        const unsigned blocks = (unsigned)fsw_cdiv(tot, 256);
        coal_keys<<<blocks, 256, 0, st>>>(edge_index, edge_index + E, E, N, self_loops, keys, ids);
        FSW_CHECK_LAUNCH("coal_keys");
        size_t tb = sort64_temp_bytes(totmax);
        int bits = 2 * key_bits(N) + 1;
        if (bits > 64) bits = 64;
        FSW_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_s, ids, ids_s, (int)tot, 0, bits, st));
        fsw_count_launch(4);
        coal_heads<<<blocks, 256, 0, st>>>(keys_s, tot, N, heads);
        FSW_CHECK_LAUNCH("coal_heads");
        int rc = exclusive_scan_i32(heads, tot, excl, btmp, st);
        if (rc) return rc;
        if (dtype == FSW_F32)
            coal_emit<float><<<blocks, 256, 0, st>>>(keys_s, ids_s, heads, excl, tot, E, N, self_loop_weight, col, (float*)W, slot_of_elem, counts, nslots);
        else
            coal_emit<double><<<blocks, 256, 0, st>>>(keys_s, ids_s, heads, excl, tot, E, N, self_loop_weight, col, (double*)W, slot_of_elem, counts, nslots);
        FSW_CHECK_LAUNCH("coal_emit");
    }
    int rc = exclusive_scan_i32(counts, N + 1, rowptr, btmp, st);
    if (rc) return rc;
    if (N > 0) {
        const unsigned nb = (unsigned)fsw_cdiv(N, 256);
        for (int phase = 0; phase < (gcn ? 2 : 1); ++phase) {
            if (dtype == FSW_F32) coal_weights<float><<<nb, 256, 0, st>>>(rowptr, col, N, (float*)deg, (float*)W, phase);
            else coal_weights<double><<<nb, 256, 0, st>>>(rowptr, col, N, (double*)deg, (double*)W, phase);
            FSW_CHECK_LAUNCH("coal_weights");
        }
    }
    return FSW_OK;
}

extern "C" size_t fsw_csr_workspace_bytes(int64_t N, int64_t E) {
    // counts [N+1] + block sums | keys, ids, sorted keys, sorted ids [E + N] each | cub temporary storage
    const int64_t tot = E + N;
    const size_t head = ((size_t)((N + 1) + fsw_cdiv(N + 1, SCAN_BLOCK) + 64) * sizeof(int) + 255) & ~(size_t)255;
    const size_t arr = ((size_t)(tot > 0 ? tot : 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    return head + 4 * arr + sort_temp_bytes(tot);
}

extern "C" int fsw_csr_from_edge_index(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int32_t* rowptr,
                                       int32_t* col, int32_t* eid, void* workspace, size_t workspace_bytes, void* stream) {
    if (N < 0 || E < 0) return fsw_fail(FSW_ERR_INVALID, "fsw_csr_from_edge_index: negative size");
    if (E + (self_loops ? N : 0) >= (int64_t)INT32_MAX) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_csr_from_edge_index: more than 2^31 elements");
    if (workspace_bytes < fsw_csr_workspace_bytes(N, E)) return fsw_fail(FSW_ERR_WORKSPACE, "fsw_csr_from_edge_index: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tot_max = E + N;
    const size_t head = ((size_t)((N + 1) + fsw_cdiv(N + 1, SCAN_BLOCK) + 64) * sizeof(int) + 255) & ~(size_t)255;
    const size_t arr = ((size_t)(tot_max > 0 ? tot_max : 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    unsigned char* wsb = (unsigned char*)workspace;
    int* counts = (int*)wsb;
    int* btmp = counts + (N + 1);
    int32_t* keys = (int32_t*)(wsb + head);
    int32_t* ids = (int32_t*)(wsb + head + arr);
    int32_t* keys_out = (int32_t*)(wsb + head + 2 * arr);
    int32_t* ids_out = eid ? eid : (int32_t*)(wsb + head + 3 * arr);
    void* temp = wsb + head + 4 * arr;
    FSW_CUDA(cudaMemsetAsync(counts, 0, (size_t)(N + 1) * sizeof(int), st));
    const int64_t tot = E + (self_loops ? N : 0);
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    if (tot > 0) {
        csr_count<<<(unsigned)fsw_cdiv(tot, 256), 256, 0, st>>>(dst, E, N, self_loops, counts);
        FSW_CHECK_LAUNCH("csr_count");
    }
    int rc = exclusive_scan_i32(counts, N + 1, rowptr, btmp, st);
    if (rc) return rc;
    if (tot > 0) {
        csr_keys<<<(unsigned)fsw_cdiv(tot, 256), 256, 0, st>>>(dst, E, N, self_loops, keys, ids);
        FSW_CHECK_LAUNCH("csr_keys");
        size_t tb = sort_temp_bytes(tot_max);
        FSW_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, keys, keys_out, ids, ids_out, (int)tot, 0, key_bits(N), st));
        fsw_count_launch(3);
        // elements with an out-of-range destination (key N) sort behind every segment: the gather fills their slots too, but they
        // lie beyond rowptr[N] only if col has room - callers size col / eid with E (+ N), so every slot is inside
        csr_gather_cols<<<(unsigned)fsw_cdiv(tot, 256), 256, 0, st>>>(src, E, tot, ids_out, col);
        FSW_CHECK_LAUNCH("csr_gather_cols");
    }
    return FSW_OK;
}

extern "C" int fsw_rowptr_from_sorted_rows(const int64_t* rows, int64_t nnz, int64_t S, int32_t* rowptr, void* stream) {
    if (nnz >= (int64_t)INT32_MAX) return fsw_fail(FSW_ERR_UNSUPPORTED, "more than 2^31 nonzeros");
    const int64_t work = nnz > 0 ? nnz : S + 1;
    rowptr_from_rows<<<(unsigned)fsw_cdiv(work, 256), 256, 0, (cudaStream_t)stream>>>(rows, nnz, S, rowptr);
    FSW_CHECK_LAUNCH("rowptr_from_rows");
    return FSW_OK;
}

extern "C" int fsw_edge_weights(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* eid, int64_t N,
                                int64_t E, int self_loops, double self_loop_weight, int gcn, void* deg_out, void* w_out,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!deg_out) return fsw_fail(FSW_ERR_INVALID, "fsw_edge_weights: deg_out is NULL");
    if ((gcn || self_loops) && !w_out) return fsw_fail(FSW_ERR_INVALID, "fsw_edge_weights: w_out is NULL");
    if (self_loops && !eid) return fsw_fail(FSW_ERR_INVALID, "fsw_edge_weights: eid needed with self loops");
    if (N == 0) return FSW_OK;
    const unsigned grid = (unsigned)fsw_cdiv(N, 256);
    for (int phase = 0; phase < 2; ++phase) {
        if (dtype == FSW_F32)
            edge_weights_kernel<float><<<grid, 256, 0, st>>>(rowptr, col, eid, N, E, self_loops, self_loop_weight, gcn, (float*)deg_out, (float*)w_out, phase);
        else if (dtype == FSW_F64)
            edge_weights_kernel<double><<<grid, 256, 0, st>>>(rowptr, col, eid, N, E, self_loops, self_loop_weight, gcn, (double*)deg_out, (double*)w_out, phase);
        else
            return fsw_fail(FSW_ERR_INVALID, "fsw_edge_weights: dtype %d", dtype);
        FSW_CHECK_LAUNCH("edge_weights_kernel");
    }
    return FSW_OK;
}

extern "C" size_t fsw_plan_workspace_bytes(int64_t S) {
    (void)S;
    return (size_t)(2 * (FSW_PLAN_BUCKETS + 8)) * sizeof(int);
}

extern "C" int fsw_segment_plan(int dtype, const int32_t* rowptr, int64_t n_fixed, const void* W, int64_t S, double thresh,
                                double* mass, int32_t* info, int32_t* order, int32_t* bucket_offsets,
                                int64_t* bucket_elems, void* workspace, size_t workspace_bytes, void* stream) {
    if (workspace_bytes < fsw_plan_workspace_bytes(S)) return fsw_fail(FSW_ERR_WORKSPACE, "fsw_segment_plan: workspace too small");
    if (!(thresh > 0)) return fsw_fail(FSW_ERR_INVALID, "fsw_segment_plan: thresh must be positive");
    if (!rowptr && n_fixed <= 0) return fsw_fail(FSW_ERR_INVALID, "fsw_segment_plan: rowptr == NULL needs n_fixed > 0");
    cudaStream_t st = (cudaStream_t)stream;
    int* hist = (int*)workspace;
    int* cursor = hist + (FSW_PLAN_BUCKETS + 8);
    FSW_CUDA(cudaMemsetAsync(hist, 0, (size_t)(2 * (FSW_PLAN_BUCKETS + 8)) * sizeof(int), st));
    if (bucket_elems) FSW_CUDA(cudaMemsetAsync(bucket_elems, 0, (size_t)FSW_PLAN_BUCKETS * sizeof(int64_t), st));
    if (S > 0) {
        const unsigned grid = (unsigned)fsw_cdiv(S * 32, 256);
        if (dtype == FSW_F32)
            plan_stats<float><<<grid, 256, 0, st>>>(rowptr, n_fixed, (const float*)W, S, thresh, mass, info, hist, (unsigned long long*)bucket_elems);
        else if (dtype == FSW_F64)
            plan_stats<double><<<grid, 256, 0, st>>>(rowptr, n_fixed, (const double*)W, S, thresh, mass, info, hist, (unsigned long long*)bucket_elems);
        else
            return fsw_fail(FSW_ERR_INVALID, "fsw_segment_plan: dtype %d", dtype);
        FSW_CHECK_LAUNCH("plan_stats");
    }
    plan_scan<<<1, 32, 0, st>>>(hist, bucket_offsets, cursor);
    FSW_CHECK_LAUNCH("plan_scan");
    if (S > 0) {
        plan_scatter<<<(unsigned)fsw_cdiv(S, 256), 256, 0, st>>>(info, S, cursor, order);
        FSW_CHECK_LAUNCH("plan_scatter");
    }
    return FSW_OK;
}
