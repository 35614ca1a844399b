// Kseg: segmented inclusive cumulative sum.
//
//  * fsw_segcumsum: single-pass scan with decoupled look-back across tiles (one read and one write of
//    the values, one read of the ids).  Replaces the multi-level hierarchy that the reference builds
//    in Python around its two kernels (fsw_embedding.py:2878-3012, fsw_embedding.cu:35-117).
//  * the reference's legacy entry points (segcumsum_wrapper, add_block_sums_wrapper, launch_*,
//    get_max_threads_per_block) with their per-call semantics, so that the reference's own
//    segcumsum_cuda keeps working against this library.
#include "fsw_common.cuh"

namespace {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

// tile status words
constexpr int ST_INVALID = 0, ST_AGG = 1, ST_PREFIX = 2;

template <typename T>
__device__ __forceinline__ void seg_combine(T& v, int& f, T pv, int pf) {
    // (v, f) = (pv, pf) (+) (v, f)   : left operand precedes
    if (!f) v += pv;
    f |= pf;
}

template <typename T, typename IdT>
__global__ void __launch_bounds__(SC_THREADS) fsw_segcumsum_kernel(const T* in, T* out,
                                                                   const IdT* __restrict__ ids, int64_t n, int* ticket,
                                                                   volatile int* status, volatile int* heads,
                                                                   volatile T* aggs, volatile T* prefixes) {
    __shared__ T sv[SC_TILE];
    __shared__ unsigned char sh[SC_TILE];
    __shared__ T warp_v[SC_THREADS / 32];
    __shared__ int warp_f[SC_THREADS / 32];
    __shared__ int s_tile;
    __shared__ T s_carry;

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const int64_t base = (int64_t)tile * SC_TILE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // coalesced load (striped) into shared memory
#pragma unroll
    for (int it = 0; it < SC_ITEMS; ++it) {
        const int loc = it * SC_THREADS + tid;
        const int64_t i = base + loc;
        T v = (T)0;
        unsigned char h = 1;
        if (i < n) {
            v = in[i];
            h = (i == 0) ? 1 : (ids[i] != ids[i - 1]);
        }
        sv[loc] = v;
        sh[loc] = h;
    }
    __syncthreads();

    // blocked: thread owns SC_ITEMS consecutive items
    T x[SC_ITEMS];
    int hd[SC_ITEMS];
    T run = (T)0;
    int any = 0;
#pragma unroll
    for (int it = 0; it < SC_ITEMS; ++it) {
        const int loc = tid * SC_ITEMS + it;
        const T v = sv[loc];
        hd[it] = sh[loc];
        if (hd[it]) {
            run = v;
            any = 1;
        } else {
            run += v;
        }
        x[it] = run;
    }
    // inclusive segmented scan of (run, any) across the warp
    T wv = run;
    int wf = any;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T pv = __shfl_up_sync(0xffffffffu, wv, d);
        int pf = __shfl_up_sync(0xffffffffu, wf, d);
        if (lane >= d) seg_combine(wv, wf, pv, pf);
    }
    if (lane == 31) {
        warp_v[warp] = wv;
        warp_f[warp] = wf;
    }
    __syncthreads();
    // exclusive carry of this thread inside the tile: warps before + lanes before
    T cv = (T)0;
    int cf = 0;
    for (int w = 0; w < warp; ++w) {  // combine in order: carry = carry (+) warp_w
        T v2 = warp_v[w];
        int f2 = warp_f[w];
        seg_combine(v2, f2, cv, cf);
        cv = v2;
        cf = f2;
    }
    {
        T pv = __shfl_up_sync(0xffffffffu, wv, 1);
        int pf = __shfl_up_sync(0xffffffffu, wf, 1);
        if (lane > 0) {
            seg_combine(pv, pf, cv, cf);
            cv = pv;
            cf = pf;
        }
    }
    // tile aggregate (computed by the last thread)
    if (tid == SC_THREADS - 1) {
        // combine all warps
        T av = (T)0;
        int af = 0;
        for (int w = 0; w < SC_THREADS / 32; ++w) {
            T v2 = warp_v[w];
            int f2 = warp_f[w];
            seg_combine(v2, f2, av, af);
            av = v2;
            af = f2;
        }
        // publish aggregate, then look back.  `aggs` entries are written once; `prefixes` entries are
        // written once, before the status turns ST_PREFIX, so a reader never sees a half-updated pair.
        T carry = (T)0;
        aggs[tile] = av;
        heads[tile] = af;
        if (tile == 0 || af) {
            // a tile that contains a head (or tile 0) already knows the running sum at its end
            prefixes[tile] = av;
            __threadfence();
            status[tile] = ST_PREFIX;
        } else {
            __threadfence();
            status[tile] = ST_AGG;
        }
        if (tile > 0) {
            // carry-in for the elements before the first head of this tile
            int p = tile - 1;
            while (true) {
                int st;
                while ((st = status[p]) == ST_INVALID) {
                }
                __threadfence();
                if (st == ST_PREFIX) {
                    carry += prefixes[p];
                    break;
                }
                carry += aggs[p];  // ST_AGG: tile p has no head, keep walking back
                --p;
            }
            if (!af) {
                prefixes[tile] = av + carry;
                __threadfence();
                status[tile] = ST_PREFIX;
            }
        }
        s_carry = carry;
    }
    __syncthreads();
    const T tile_carry = s_carry;
    // final values: items before the first head seen (in thread, then in tile) get the carries
    T add = cf ? cv : cv + tile_carry;  // carry entering this thread
    // if no head precedes inside the tile (cf == 0) the tile carry applies as well
    bool open = true;
#pragma unroll
    for (int it = 0; it < SC_ITEMS; ++it) {
        if (hd[it]) open = false;
        sv[tid * SC_ITEMS + it] = open ? x[it] + add : x[it];
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SC_ITEMS; ++it) {
        const int loc = it * SC_THREADS + tid;
        const int64_t i = base + loc;
        if (i < n) out[i] = sv[loc];
    }
}

// ---------------------------------------------------------------------------------------------------
// Legacy kernels (per-call semantics of fsw_embedding.cu:35-98 and :103-117)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void legacy_segcumsum_kernel(T* values, const int64_t* segment_ids, int64_t size, T* block_sums_out,
                                        int64_t* block_last_ids_out, bool return_next_level) {
    __shared__ T sv[1024];
    __shared__ unsigned char sf[1024];
    const int tid = threadIdx.x;
    const int64_t index = (int64_t)blockIdx.x * blockDim.x + tid;
    const bool ok = index < size;
    int64_t id_curr = 0;
    T v = (T)0;
    unsigned char f = 1;
    if (ok) {
        v = values[index];
        id_curr = segment_ids[index];
        f = (tid == 0) ? 1 : (segment_ids[index - 1] != id_curr);
    }
    sv[tid] = v;
    sf[tid] = f;
    __syncthreads();
    for (int stride = 1; stride < (int)blockDim.x; stride <<= 1) {
        T pv = (T)0;
        unsigned char pf = 0;
        const bool take = tid >= stride;
        if (take) {
            pv = sv[tid - stride];
            pf = sf[tid - stride];
        }
        __syncthreads();
        if (take) {
            if (!sf[tid]) sv[tid] += pv;
            sf[tid] |= pf;
        }
        __syncthreads();
    }
    if (ok) values[index] = sv[tid];
    if (return_next_level && tid == (int)blockDim.x - 1 && ok) {
        block_sums_out[blockIdx.x] = sv[tid];
        block_last_ids_out[blockIdx.x] = id_curr;
    }
}

template <typename T>
__global__ void legacy_add_block_sums_kernel(T* output, const T* block_sums, const int64_t* segment_ids,
                                             const int64_t* block_last_id, int64_t size) {
    const int64_t index = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (index < size && blockIdx.x >= 1) {
        if (block_last_id[blockIdx.x - 1] == segment_ids[index]) output[index] += block_sums[blockIdx.x - 1];
    }
}

void legacy_check(cudaError_t e, const char* what) {
    // The reference printf()s and exit(1)s here (fsw_embedding.cu:20-27); we record the error instead.
    if (e != cudaSuccess) fsw_fail(FSW_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

}  // namespace

extern "C" size_t fsw_segcumsum_workspace_bytes(int64_t n) {
    const int64_t tiles = fsw_cdiv(n > 0 ? n : 1, SC_TILE);
    return (size_t)(64 + tiles * (2 * sizeof(int) + 2 * sizeof(double)) + 64);
}

template <typename T, typename IdT>
static int segcumsum_launch(const void* in, void* out, const void* ids, int64_t n, void* ws, cudaStream_t st) {
    const int64_t tiles = fsw_cdiv(n, SC_TILE);
    unsigned char* p = (unsigned char*)ws;
    int* ticket = (int*)p;
    double* aggs = (double*)(p + 64);
    double* prefixes = aggs + tiles;
    int* status = (int*)(prefixes + tiles);
    int* heads = status + tiles;
    cudaError_t e = cudaMemsetAsync(ws, 0, fsw_segcumsum_workspace_bytes(n), st);
    if (e != cudaSuccess) return fsw_fail(FSW_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    fsw_segcumsum_kernel<T, IdT><<<(unsigned)tiles, SC_THREADS, 0, st>>>((const T*)in, (T*)out, (const IdT*)ids, n, ticket, status,
                                                                        heads, (volatile T*)aggs, (volatile T*)prefixes);
    FSW_CHECK_LAUNCH("fsw_segcumsum_kernel");
    return FSW_OK;
}

extern "C" int fsw_segcumsum(int dtype, const void* values_in, void* values_out, const void* segment_ids, int id_bytes,
                             int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
    if (n == 0) return FSW_OK;
    if (!values_in || !values_out || !segment_ids || !workspace) return fsw_fail(FSW_ERR_INVALID, "fsw_segcumsum: null argument");
    if (workspace_bytes < fsw_segcumsum_workspace_bytes(n)) return fsw_fail(FSW_ERR_WORKSPACE, "fsw_segcumsum: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FSW_F32 && id_bytes == 8) return segcumsum_launch<float, int64_t>(values_in, values_out, segment_ids, n, workspace, st);
    if (dtype == FSW_F32 && id_bytes == 4) return segcumsum_launch<float, int32_t>(values_in, values_out, segment_ids, n, workspace, st);
    if (dtype == FSW_F64 && id_bytes == 8) return segcumsum_launch<double, int64_t>(values_in, values_out, segment_ids, n, workspace, st);
    if (dtype == FSW_F64 && id_bytes == 4) return segcumsum_launch<double, int32_t>(values_in, values_out, segment_ids, n, workspace, st);
    return fsw_fail(FSW_ERR_INVALID, "fsw_segcumsum: dtype %d / id_bytes %d", dtype, id_bytes);
}

// ---- legacy C ABI -----------------------------------------------------------------------------------
extern "C" void launch_segcumsum_kernel_float(float* values, const int64_t* segment_ids, int64_t size, int64_t max_seg_size,
                                              float* block_sums_out, int64_t* block_last_ids_out, bool return_next_level,
                                              int64_t num_blocks, int64_t threads_per_block, int64_t shared_memory_size) {
    (void)max_seg_size;
    (void)shared_memory_size;
    if (num_blocks <= 0 || threads_per_block <= 0 || threads_per_block > 1024) {
        fsw_fail(FSW_ERR_INVALID, "launch_segcumsum_kernel_float: bad launch shape");
        return;
    }
    legacy_segcumsum_kernel<float><<<(unsigned)num_blocks, (unsigned)threads_per_block>>>(values, segment_ids, size, block_sums_out,
                                                                                         block_last_ids_out, return_next_level);
    legacy_check(cudaGetLastError(), "legacy_segcumsum_kernel<float>");
    fsw_count_launch();
}

extern "C" void launch_segcumsum_kernel_double(double* values, const int64_t* segment_ids, int64_t size, int64_t max_seg_size,
                                               double* block_sums_out, int64_t* block_last_ids_out, bool return_next_level,
                                               int64_t num_blocks, int64_t threads_per_block, int64_t shared_memory_size) {
    (void)max_seg_size;
    (void)shared_memory_size;
    if (num_blocks <= 0 || threads_per_block <= 0 || threads_per_block > 1024) {
        fsw_fail(FSW_ERR_INVALID, "launch_segcumsum_kernel_double: bad launch shape");
        return;
    }
    legacy_segcumsum_kernel<double><<<(unsigned)num_blocks, (unsigned)threads_per_block>>>(values, segment_ids, size, block_sums_out,
                                                                                          block_last_ids_out, return_next_level);
    legacy_check(cudaGetLastError(), "legacy_segcumsum_kernel<double>");
    fsw_count_launch();
}

extern "C" void launch_add_block_sums_kernel_float(float* output, const float* block_sums, const int64_t* segment_ids,
                                                   const int64_t* block_last_id, int64_t size, int64_t num_blocks,
                                                   int64_t threads_per_block) {
    if (num_blocks <= 0 || threads_per_block <= 0) return;
    legacy_add_block_sums_kernel<float><<<(unsigned)num_blocks, (unsigned)threads_per_block>>>(output, block_sums, segment_ids, block_last_id, size);
    legacy_check(cudaGetLastError(), "legacy_add_block_sums_kernel<float>");
    fsw_count_launch();
}

extern "C" void launch_add_block_sums_kernel_double(double* output, const double* block_sums, const int64_t* segment_ids,
                                                    const int64_t* block_last_id, int64_t size, int64_t num_blocks,
                                                    int64_t threads_per_block) {
    if (num_blocks <= 0 || threads_per_block <= 0) return;
    legacy_add_block_sums_kernel<double><<<(unsigned)num_blocks, (unsigned)threads_per_block>>>(output, block_sums, segment_ids, block_last_id, size);
    legacy_check(cudaGetLastError(), "legacy_add_block_sums_kernel<double>");
    fsw_count_launch();
}

extern "C" void segcumsum_wrapper(int64_t dtype, void* values, const int64_t* segment_ids, int64_t size, int64_t max_seg_size,
                                  void* block_sums_out, int64_t* block_last_ids_out, bool return_next_level, int64_t num_blocks,
                                  int64_t threads_per_block, size_t shared_memory_size) {
    legacy_check(cudaDeviceSynchronize(), "segcumsum_wrapper: sync before");
    switch ((int)dtype) {
        case FSW_F32:
            launch_segcumsum_kernel_float((float*)values, segment_ids, size, max_seg_size, (float*)block_sums_out, block_last_ids_out,
                                          return_next_level, num_blocks, threads_per_block, (int64_t)shared_memory_size);
            break;
        case FSW_F64:
            launch_segcumsum_kernel_double((double*)values, segment_ids, size, max_seg_size, (double*)block_sums_out, block_last_ids_out,
                                           return_next_level, num_blocks, threads_per_block, (int64_t)shared_memory_size);
            break;
        default:
            fsw_fail(FSW_ERR_INVALID, "segcumsum_wrapper: dtype %lld", (long long)dtype);
    }
    legacy_check(cudaDeviceSynchronize(), "segcumsum_wrapper: sync after");
}

extern "C" void add_block_sums_wrapper(int64_t dtype, void* output, const void* block_sums, const int64_t* segment_ids,
                                       const int64_t* block_last_id, int64_t size, int64_t num_blocks, int64_t threads_per_block) {
    legacy_check(cudaDeviceSynchronize(), "add_block_sums_wrapper: sync before");
    switch ((int)dtype) {
        case FSW_F32:
            launch_add_block_sums_kernel_float((float*)output, (const float*)block_sums, segment_ids, block_last_id, size, num_blocks, threads_per_block);
            break;
        case FSW_F64:
            launch_add_block_sums_kernel_double((double*)output, (const double*)block_sums, segment_ids, block_last_id, size, num_blocks, threads_per_block);
            break;
        default:
            fsw_fail(FSW_ERR_INVALID, "add_block_sums_wrapper: dtype %lld", (long long)dtype);
    }
    legacy_check(cudaDeviceSynchronize(), "add_block_sums_wrapper: sync after");
}

extern "C" int get_max_threads_per_block(int device_index) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerBlock, device_index) != cudaSuccess) return 0;
    return v;
}
