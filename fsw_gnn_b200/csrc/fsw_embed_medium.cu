// K2 / K3 medium path: segments of 65 .. 512 elements with uniform weights (fp32).
//
// One CTA owns a tile [n][32 slices] of one (segment, slice-chunk); lanes are slices, so every shared
// memory access of a warp hits 32 different banks whatever the rows are (bank = lane).
//   1. each warp loads runs of 32 rows straight into registers (coalesced 128-byte gathers), sorts them
//      with the compile-time network and stores the sorted run to shared memory;
//   2. merge passes double the run length: every 32 output rows are one work unit; a unit finds its
//      start with a merge-path binary search (fixed trip count -> no divergence) and then runs a
//      branch-free sequential 2-way merge; units are dealt round-robin to the warps;
//   3. the LAST pass never stores its output: the merged stream is consumed on the fly -
//        forward : acc += p_(j) * cos(pi xi (2j+1)/n), coefficients from a per-(n, slice) table kept in
//                  shared memory and rebuilt only when n changes (segments arrive sorted by n);
//        backward: dL/dp_(j) goes to row idx_j of the free ping-pong buffer (un-permutation), then the
//                  tile is scattered to dXp in ORIGINAL element order, one coalesced row per element.
// Tiles up to 512 rows live in shared memory.  Larger segments (up to the biggest hub) run the SAME code
// with the two ping-pong buffers in an L2-resident global scratch slice per CTA (GLOBAL = true; the
// coefficient table is then replaced by direct evaluation) on a persistent grid.
// Non-uniform weights and fp64 use the generic path (fsw_embed.cu).
#include "fsw_sortnet.cuh"

namespace {

constexpr int UNIT = 32;  // output rows per merge work unit

// merge-path split: number of A elements among the first `oo` outputs of merge(A, B), ties -> A first
template <typename T>
__device__ __forceinline__ int fsw_merge_path(const T* A, int La, const T* B, int Lb, int oo, int iters, int lane) {
    int lo = max(0, oo - Lb), hi = min(oo, La);
    for (int it = 0; it < iters; ++it) {
        if (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const T av = A[mid * 32 + lane];
            const T bv = B[(oo - mid - 1) * 32 + lane];
            if (av <= bv)
                lo = mid + 1;
            else
                hi = mid;
        }
    }
    return lo;
}

template <typename T>
__device__ __forceinline__ T fsw_inf();
template <>
__device__ __forceinline__ float fsw_inf<float>() { return __int_as_float(0x7f800000); }

// ---------------------------------------------------------------------------------------------------
template <typename T, typename IdxT, int W, bool BWD, bool NEED_DXI, bool GLOBAL>
__global__ void __launch_bounds__(W * 32) fsw_medium_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks, int cap,
                                                            int64_t nwork, T* __restrict__ out, int64_t ld_out,
                                                            int64_t out_col0, const T* __restrict__ bias,
                                                            const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                            T* __restrict__ dXp, T* __restrict__ dEp,
                                                            double* __restrict__ dfreqs, unsigned char* gscratch) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[W][32];
    __shared__ double red2[W][32];
    constexpr bool USE_TABLE = !BWD && !GLOBAL;
    const size_t tile_bytes = (size_t)cap * 32 * (BWD ? 2 * (sizeof(T) + sizeof(IdxT)) : (USE_TABLE ? 3 : 2) * sizeof(T));
    unsigned char* basep = GLOBAL ? gscratch + (size_t)blockIdx.x * tile_bytes : fsw_smem_raw;
    T* bufA = reinterpret_cast<T*>(basep);
    T* bufB = bufA + (size_t)cap * 32;
    T* table = bufB + (size_t)cap * 32;                                     // forward, shared-memory tiles only
    IdxT* idxA = reinterpret_cast<IdxT*>(bufB + (size_t)cap * 32);          // backward only
    IdxT* idxB = idxA + (size_t)cap * 32;
    (void)table;
    (void)idxA;
    (void)idxB;
    (void)red2;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int n_prev = -1;
    T A0 = (T)0, A0p = (T)0;
    (void)A0p;
    const T INF = fsw_inf<T>();

  for (int64_t work = blockIdx.x; work < nwork; work += gridDim.x) {   // persistent over (item, chunk) work groups
    const int item = (int)(work / nchunks);
    const int chunk = (int)(work - (int64_t)item * nchunks);
    const int64_t first = (int64_t)seg_lo + (int64_t)item * G;
    if (first >= seg_hi) continue;
    const int last = (int)((first + G < seg_hi) ? first + G : seg_hi);
    n_prev = -1;
    const int k = chunk * 32 + lane;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const T xi = fsw_ldg(a.freqs + kk);
    const double xid = (double)xi;
    const T bk = (!BWD && bias != nullptr) ? fsw_ldg(bias + kk) : (T)0;
    double dxi_acc = 0.0;
    (void)dxi_acc;

    for (int q = (int)first; q < last; ++q) {
        const int s = a.order ? a.order[q] : q;
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);
        const int R = (n + 31) >> 5;
        const int Ntot = R << 5;
        const double u = xid / (double)n;
        const T wn = (T)(1.0 / (double)n);

        if (n != n_prev) {
            fsw_amplitude<T, NEED_DXI>(u, wn, xi, A0, A0p);
            if (USE_TABLE) {
                // coefficient table for this (n, slice): first read happens after the barrier below
                for (int r = warp; r < Ntot; r += W)
                    table[r * 32 + lane] = (r < n) ? Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * r + 1))) : (T)0;
            }
            n_prev = n;
        }

        // ---- 1. runs of 32 rows: gather -> registers -> network sort -> shared ----
        for (int run = warp; run < R; run += W) {
            const int base = run << 5;
            const int cnt = min(32, n - base);
            int c0, c1;
            T key[32];
            int idx[32];
            (void)idx;
            fsw_gather_keys<T, 32>(a, e0 + base, cnt, kk, lane, key, c0, c1);
            if (BWD) {
#pragma unroll
                for (int j = 0; j < 32; ++j) idx[j] = base + j;
            }
            if (BWD) {
                fsw_sort_network<32>([&](int i, int l) {
                    T x = key[i], y = key[l];
                    int px = idx[i], py = idx[l];
                    bool sw = x > y;
                    key[i] = sw ? y : x;
                    key[l] = sw ? x : y;
                    idx[i] = sw ? py : px;
                    idx[l] = sw ? px : py;
                });
            } else {
                fsw_sort_network<32>([&](int i, int l) {
                    T x = key[i], y = key[l];
                    key[i] = fmin(x, y);
                    key[l] = fmax(x, y);
                });
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                bufA[(base + j) * 32 + lane] = key[j];
                if (BWD) idxA[(base + j) * 32 + lane] = (IdxT)idx[j];
            }
        }
        __syncthreads();

        // ---- 2./3. merge passes; the last one is consumed instead of stored ----
        T* src = bufA;
        T* dst = bufB;
        IdxT* isrc = idxA;
        IdxT* idst = idxB;
        const int nunits = Ntot / UNIT;
        const T gk = (BWD && act) ? g[(int64_t)s * ld_g + g_col0 + k] : (T)0;
        const T GA = gk * ((T)1 + xi) * A0;
        double acc = 0.0, Sc = 0.0, Ss = 0.0;
        (void)GA;
        (void)Sc;
        (void)Ss;
        for (int len = 32;; len <<= 1) {
            const bool final_pass = (len << 1) >= Ntot;
            const int iters = 33 - __clz(len);
            for (int un = warp; un < nunits; un += W) {
                const int o = un * UNIT;
                const int pair = o / (2 * len);
                const int p0 = pair * 2 * len;
                const int La = min(len, Ntot - p0);
                const int Lb = min(len, max(0, Ntot - p0 - len));
                const int oo = o - p0;
                const T* Aq = src + (size_t)p0 * 32;
                const T* Bq = Aq + (size_t)La * 32;
                int ia = (oo > 0) ? fsw_merge_path(Aq, La, Bq, Lb, oo, iters, lane) : 0;
                int ib = oo - ia;
                T av = (ia < La) ? Aq[ia * 32 + lane] : INF;
                T bv = (ib < Lb) ? Bq[ib * 32 + lane] : INF;
#pragma unroll 4
                for (int t = 0; t < UNIT; ++t) {
                    const bool takeA = av <= bv;
                    const T v = takeA ? av : bv;
                    const int pos = o + t;
                    int src_row;  // row (inside the pair) the value came from
                    if (takeA) {
                        src_row = ia;
                        ++ia;
                    } else {
                        src_row = La + ib;
                        ++ib;
                    }
                    if (!final_pass) {
                        dst[pos * 32 + lane] = v;
                        if (BWD) idst[pos * 32 + lane] = isrc[(p0 + src_row) * 32 + lane];
                    } else if (!BWD) {
                        if (USE_TABLE)
                            acc += (double)(v * table[pos * 32 + lane]);
                        else if (pos < n)
                            acc += (double)(v * Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * pos + 1))));
                    } else if (pos < n) {
                        const T r = Num<T>::reduce(u * (double)(2 * pos + 1));
                        const T c = Num<T>::cospi_(r);
                        const int id = isrc[(p0 + src_row) * 32 + lane];
                        dst[id * 32 + lane] = GA * c;
                        if (NEED_DXI) {
                            Sc += (double)(v * c);
                            Ss += (double)(v * ((T)M_PI * wn * (T)(2 * pos + 1) * Num<T>::sinpi_(r)));
                        }
                    }
                    // refill the head that was consumed
                    const int nxt = takeA ? ia : ib;
                    const int lim = takeA ? La : Lb;
                    const T* base_ptr = takeA ? Aq : Bq;
                    const T nv = (nxt < lim) ? base_ptr[nxt * 32 + lane] : INF;
                    if (takeA)
                        av = nv;
                    else
                        bv = nv;
                }
            }
            __syncthreads();
            if (final_pass) break;
            T* tk = src;
            src = dst;
            dst = tk;
            IdxT* ti = isrc;
            isrc = idst;
            idst = ti;
        }

        if (!BWD) {
            red[warp][lane] = acc;
            __syncthreads();
            if (warp == 0 && act) {
                double tot = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < W; ++w2) tot += red[w2][lane];
                out[(int64_t)s * ld_out + out_col0 + k] = ((T)1 + xi) * A0 * (T)tot + bk;
            }
        } else {
            // dst now holds dL/dp in ORIGINAL row order
            for (int r = warp; r < n; r += W) {
                if (act) {
                    const T v = dst[r * 32 + lane];
                    if (a.col)
                        atomicAdd(dXp + (int64_t)a.col[e0 + r] * a.ldp + k, v);
                    else
                        dXp[(e0 + r) * a.ldp + k] = v;
                    if (dEp) dEp[(e0 + r) * a.ldp + k] = v;
                }
            }
            if (NEED_DXI) {
                red[warp][lane] = Sc;
                red2[warp][lane] = Ss;
                __syncthreads();
                if (warp == 0) {
                    double tc = 0.0, ts = 0.0;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) {
                        tc += red[w2][lane];
                        ts += red2[w2][lane];
                    }
                    dxi_acc += (double)gk * ((double)A0 * tc + (1.0 + xid) * ((double)A0p * tc - (double)A0 * ts));
                }
            }
        }
        __syncthreads();
    }
    if (BWD && NEED_DXI && warp == 0 && act) atomicAdd(dfreqs + k, dxi_acc);
  }
}

const int kPersistentGrid = 148 * 2;

template <typename T, typename IdxT, int W, bool BWD, bool NEED_DXI, bool GLOBAL>
int launch_medium(const SegArgs<T>& a, int lo, int hi, int cap, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                  const T* g, int64_t ld_g, int64_t g_col0, T* dXp, T* dEp, double* dfreqs, void* scratch,
                  size_t scratch_bytes, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int64_t cnt = hi - lo;
    int64_t G = GLOBAL ? 1 : cnt * nchunks / (148 * 8);
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    const int64_t nwork = fsw_cdiv(cnt, G) * nchunks;
    constexpr bool USE_TABLE = !BWD && !GLOBAL;
    const size_t tile = (size_t)cap * 32 * (BWD ? 2 * (sizeof(T) + sizeof(IdxT)) : (USE_TABLE ? 3 : 2) * sizeof(T));
    auto kern = fsw_medium_kernel<T, IdxT, W, BWD, NEED_DXI, GLOBAL>;
    int64_t blocks = nwork;
    size_t smem = tile;
    if (GLOBAL) {
        smem = 0;
        if (blocks > kPersistentGrid) blocks = kPersistentGrid;
        if ((size_t)blocks * tile > scratch_bytes) blocks = (int64_t)(scratch_bytes / tile);
        if (blocks < 1) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small: need >= %zu bytes", tile);
    } else {
        FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const std::string label = std::string(BWD ? "bwd_medium_u" : "fwd_medium_u") + std::to_string(cap) + "_f32";
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, W * 32, smem, st>>>(a, lo, hi, (int)G, nchunks, cap, nwork, out, ld_out, out_col0, bias, g, ld_g, g_col0,
                                                  dXp, dEp, dfreqs, (unsigned char*)scratch);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_medium_kernel");
    return FSW_OK;
}

template <bool BWD, bool NEED_DXI>
int dispatch_medium(const SegArgs<float>& a, int lo, int hi, int cap, float* out, int64_t ld_out, int64_t out_col0,
                    const float* bias, const float* g, int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs,
                    void* scratch, size_t scratch_bytes, cudaStream_t st) {
#define FSW_MED(IDX, W, GLOBAL) \
    launch_medium<float, IDX, W, BWD, NEED_DXI, GLOBAL>(a, lo, hi, cap, out, ld_out, out_col0, bias, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch, scratch_bytes, st)
    if (cap <= 128) return FSW_MED(unsigned short, 4, false);
    if (cap <= 256) return FSW_MED(unsigned short, 8, false);
    if (cap <= 512) return FSW_MED(unsigned short, 16, false);
    if (cap <= 32768) return FSW_MED(unsigned short, 16, true);
    return FSW_MED(int, 16, true);
#undef FSW_MED
}

}  // namespace

size_t fsw_medium_tile_bytes(int cap, bool backward) {
    if (cap <= 512) return 0;  // shared-memory tiles
    const size_t idx = cap <= 32768 ? 2 : 4;
    return (size_t)cap * 32 * (backward ? 2 * (4 + idx) : 2 * 4);
}

int fsw_medium_grid() { return kPersistentGrid; }

// uniform-weight fp32 segments order[lo, hi) whose size class is `cap` (>= 128)
int fsw_medium_forward_f32(const SegArgs<float>& a, int lo, int hi, int cap, float* out, int64_t ld_out, int64_t out_col0,
                           const float* bias, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    return dispatch_medium<false, false>(a, lo, hi, cap, out, ld_out, out_col0, bias, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                         scratch, scratch_bytes, st);
}

int fsw_medium_backward_f32(const SegArgs<float>& a, int lo, int hi, int cap, const float* g, int64_t ld_g, int64_t g_col0,
                            float* dXp, float* dEp, double* dfreqs, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (dfreqs != nullptr)
        return dispatch_medium<true, true>(a, lo, hi, cap, nullptr, 0, 0, nullptr, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch,
                                           scratch_bytes, st);
    return dispatch_medium<true, false>(a, lo, hi, cap, nullptr, 0, 0, nullptr, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch,
                                        scratch_bytes, st);
}
