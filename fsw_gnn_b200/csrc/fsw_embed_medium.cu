// K2 / K3 medium / large path: uniform-weight fp32 segments beyond the register-sort kernels (more than 512 elements;
// 65 .. 512 only when the coefficient tables of the packed-key kernels do not fit the scratch).
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
// TEAM = true (hubs: more than FSW_TEAM_MIN_CAP elements): the whole grid, launched cooperatively with one CTA per SM, works
// on ONE tile at a time - runs and merge units are dealt to the warps of all CTAs, the phases are separated by grid-wide
// barriers, partial sums meet in a few global double accumulators.  A 100 000-element hub (configs[4]) otherwise keeps one
// CTA busy for 12 merge passes over an L2-resident tile while 147 SMs idle.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "fsw_sortnet.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int UNIT = 32;  // output rows per merge work unit
constexpr int FSW_TEAM_MIN_CAP = 8192;

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
// One merge work unit: UNIT outputs of merge(A, B) starting after (ia, ib) elements were consumed.
// A and B are runs in the same buffer `src` (element r of the buffer at src[r * 32 + lane]); rows are
// addressed with 32-bit indices ga / gb, limits ea / eb.  Branch-free: every lane runs the same UNIT steps.
// emit(t, value, payload) is called once per output in sorted order.
// ---------------------------------------------------------------------------------------------------
template <typename T, typename IdxT, bool PAY, typename F>
__device__ __forceinline__ void fsw_merge_unit(const T* __restrict__ src, const IdxT* __restrict__ isrc, int ga, int ea, int gb,
                                               int eb, int lane, T INF, F&& emit) {
    T av = (ga < ea) ? src[ga * 32 + lane] : INF;
    T bv = (gb < eb) ? src[gb * 32 + lane] : INF;
    int ai = 0, bi = 0;
    if (PAY) {
        ai = (ga < ea) ? (int)isrc[ga * 32 + lane] : 0;
        bi = (gb < eb) ? (int)isrc[gb * 32 + lane] : 0;
    }
#pragma unroll 8
    for (int t = 0; t < UNIT; ++t) {
        const bool ta = av <= bv;
        emit(t, ta ? av : bv, ta ? ai : bi);
        if (ta) ++ga; else ++gb;
        const int gn = ta ? ga : gb;
        const bool ok = gn < (ta ? ea : eb);
        T nv = INF;
        int ni = 0;
        if (ok) {
            nv = src[gn * 32 + lane];
            if (PAY) ni = (int)isrc[gn * 32 + lane];
        }
        av = ta ? nv : av;
        bv = ta ? bv : nv;
        if (PAY) {
            ai = ta ? ni : ai;
            bi = ta ? bi : ni;
        }
    }
}

// MODE 0: forward, keys only.  MODE 1: forward that also records the sorted position of every element
// (uint16 ranks, consumed by the rank-based backward).  MODE 2: backward that re-sorts (no saved ranks).
template <int MODE, bool USE_TABLE, typename T, typename IdxT>
__host__ __device__ constexpr size_t fsw_medium_tile_bytes_t(int cap) {
    return (size_t)cap * 32 * (2 * sizeof(T) + (MODE >= 1 ? 2 * sizeof(IdxT) : 0) + (USE_TABLE ? sizeof(T) : 0));
}

template <typename T, typename IdxT, int W, int MODE, bool NEED_DXI, bool GLOBAL, bool USE_TABLE, bool TEAM>
__global__ void __launch_bounds__(W * 32) fsw_medium_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks, int cap,
                                                            int64_t nwork, T* __restrict__ out, int64_t ld_out,
                                                            int64_t out_col0, const T* __restrict__ bias,
                                                            const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                            T* __restrict__ dXp, T* __restrict__ dEp,
                                                            double* __restrict__ dfreqs, unsigned char* gscratch,
                                                            unsigned short* __restrict__ ranks, int64_t ldr,
                                                            T* __restrict__ dxi_out, int64_t ld_dxi,
                                                            const T* __restrict__ gtab_c, const T* __restrict__ gtab_t) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[W][32];
    __shared__ double red2[W][32];
    constexpr bool BWD = MODE == 2;
    constexpr bool PAY = MODE >= 1;
    const size_t tile_bytes = fsw_medium_tile_bytes_t<MODE, USE_TABLE, T, IdxT>(cap);
    static_assert(!TEAM || GLOBAL, "team mode works on a tile in global scratch");
    unsigned char* basep = GLOBAL ? gscratch + (TEAM ? (size_t)0 : (size_t)blockIdx.x * tile_bytes) : fsw_smem_raw;
    double* gacc = reinterpret_cast<double*>(gscratch + tile_bytes);   // TEAM: [3][32] accumulators behind the tile, zero on entry
    (void)gacc;
    // work distribution inside a tile: the warps of this CTA, or of the whole grid
    const int gw = TEAM ? (int)blockIdx.x * W + (int)(threadIdx.x >> 5) : (int)(threadIdx.x >> 5);
    const int GW = TEAM ? (int)gridDim.x * W : W;
    auto phase_sync = [&]() {
        if constexpr (TEAM) cg::this_grid().sync();
        else __syncthreads();
    };
    T* bufA = reinterpret_cast<T*>(basep);
    T* bufB = bufA + (size_t)cap * 32;
    IdxT* idxA = reinterpret_cast<IdxT*>(bufB + (size_t)cap * 32);          // payload (MODE >= 1)
    IdxT* idxB = idxA + (size_t)cap * 32;
    T* table = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(bufB + (size_t)cap * 32) +
                                    (PAY ? 2 * (size_t)cap * 32 * sizeof(IdxT) : 0));  // forward, when it fits
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
    const int ldb = (int)(a.ldp * (int64_t)sizeof(T));

  for (int64_t work = TEAM ? 0 : blockIdx.x; work < nwork; work += TEAM ? 1 : gridDim.x) {   // persistent over (item, chunk) work groups
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
    const char* xp_bytes = reinterpret_cast<const char*>(a.Xp + kk);
    const char* ep_bytes = a.Ep ? reinterpret_cast<const char*>(a.Ep + kk) : nullptr;
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
            fsw_amplitude<T, (NEED_DXI || MODE == 1)>(u, wn, xi, A0, A0p);
            if (USE_TABLE) {
                // coefficient table for this (n, slice): first read happens after the barrier below
                for (int r = warp; r < Ntot; r += W)
                    table[r * 32 + lane] = (r < n) ? Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * r + 1))) : (T)0;
            }
            n_prev = n;
        }

        // ---- 1. runs of 32 rows: gather -> registers -> network sort -> shared ----
        for (int run = gw; run < R; run += GW) {
            const int base = run << 5;
            const int cnt = min(32, n - base);
            int c0, c1;
            T key[32];
            int idx[32];
            (void)idx;
            if (a.col) {
                fsw_load_cols<32, true>(a.col, e0 + base, cnt, lane, c0, c1);
                fsw_gather_lean<T, 32, true>(xp_bytes, ldb, ep_bytes, e0 + base, cnt, c0, c1, key);
            } else {
                fsw_gather_lean<T, 32, false>(xp_bytes, ldb, ep_bytes, e0 + base, cnt, 0, 0, key);
            }
            if (PAY) {
#pragma unroll
                for (int j = 0; j < 32; ++j) idx[j] = base + j;
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
                if (PAY) idxA[(base + j) * 32 + lane] = (IdxT)idx[j];
            }
        }
        phase_sync();

        // ---- 2. merge passes that store their output ----
        T* src = bufA;
        T* dst = bufB;
        IdxT* isrc = idxA;
        IdxT* idst = idxB;
        const int nunits = Ntot / UNIT;
        int len = 32;
        while ((len << 1) < Ntot) {
            const int iters = 33 - __clz(len);
            for (int un = gw; un < nunits; un += GW) {
                const int o = un * UNIT;
                const int p0 = (o / (2 * len)) * 2 * len;
                const int La = min(len, Ntot - p0);
                const int Lb = min(len, max(0, Ntot - p0 - len));
                const int oo = o - p0;
                const int ia = (oo > 0) ? fsw_merge_path(src + (size_t)p0 * 32, La, src + (size_t)(p0 + La) * 32, Lb, oo, iters, lane) : 0;
                T* dk = dst + o * 32 + lane;
                IdxT* di = idst + o * 32 + lane;
                fsw_merge_unit<T, IdxT, PAY>(src, isrc, p0 + ia, p0 + La, p0 + La + (oo - ia), p0 + La + Lb, lane, INF,
                                             [&](int t, T v, int id) {
                                                 dk[t * 32] = v;
                                                 if (PAY) di[t * 32] = (IdxT)id;
                                             });
            }
            phase_sync();
            T* tk = src;
            src = dst;
            dst = tk;
            IdxT* ti = isrc;
            isrc = idst;
            idst = ti;
            len <<= 1;
        }

        // ---- 3. last pass: the merged stream is consumed, not stored ----
        const bool want_dxi = (MODE == 1) && (dxi_out != nullptr);
        const bool use_gtab = (MODE == 1) && (gtab_c != nullptr) && (n <= 256);  // (legacy row-major tables; callers pass NULL)
        const T* gtc = use_gtab ? gtab_c + ((int64_t)n * (n - 1) / 2) * a.ldp + kk : nullptr;
        const T* gtt = use_gtab ? gtab_t + ((int64_t)n * (n - 1) / 2) * a.ldp + kk : nullptr;
        (void)want_dxi;
        (void)gtc;
        (void)gtt;
        const T gk = (BWD && act) ? g[fsw_rowoff(s, ld_g) + g_col0 + k] : (T)0;
        const T GA = gk * ((T)1 + xi) * A0;
        (void)GA;
        double acc = 0.0, Sc = 0.0, Ss = 0.0;
        (void)Sc;
        (void)Ss;
        {
            const int iters = 33 - __clz(len);
            const int La = min(len, Ntot);
            const int Lb = Ntot - La;
            for (int un = gw; un < nunits; un += GW) {
                const int o = un * UNIT;
                const int ia = (o > 0) ? fsw_merge_path(src, La, src + (size_t)La * 32, Lb, o, iters, lane) : 0;
                T facc = (T)0, fSc = (T)0, fSs = (T)0;
                (void)fSc;
                (void)fSs;
                fsw_merge_unit<T, IdxT, PAY>(src, isrc, ia, La, La + (o - ia), Ntot, lane, INF, [&](int t, T v, int id) {
                    const int pos = o + t;
                    if (MODE != 2) {
                        T c;
                        if (MODE == 1 && use_gtab) {
                            // coefficient tables in global memory (L1/L2 resident, shared by every CTA)
                            c = (T)0;
                            if (pos < n) {
                                c = __ldg(gtc + fsw_rowoff(pos, a.ldp));
                                if (want_dxi) fSs = fma(v, __ldg(gtt + fsw_rowoff(pos, a.ldp)), fSs);
                            }
                        } else if (MODE == 1 && want_dxi) {
                            // training with learnable frequencies: d out / d xi needs the sine as well
                            T sn = (T)0;
                            c = (T)0;
                            if (pos < n) {
                                const T r = Num<T>::reduce(u * (double)(2 * pos + 1));
                                c = Num<T>::cospi_(r);
                                sn = Num<T>::sinpi_(r);
                            }
                            fSs = fma(v, (T)M_PI * wn * (T)(2 * pos + 1) * sn, fSs);
                        } else if (USE_TABLE) {
                            c = table[pos * 32 + lane];
                        } else {
                            c = (pos < n) ? Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * pos + 1))) : (T)0;
                        }
                        facc = fma(v, c, facc);
                        // un-permute the position through the free ping-pong buffer (rows >= n are padding, never read back)
                        if (MODE == 1) dst[id * 32 + lane] = __int_as_float(pos);
                    } else {
                        T c = (T)0, sn = (T)0;
                        if (pos < n) {
                            const T r = Num<T>::reduce(u * (double)(2 * pos + 1));
                            c = Num<T>::cospi_(r);
                            if (NEED_DXI) sn = Num<T>::sinpi_(r);
                        }
                        dst[id * 32 + lane] = GA * c;
                        if (NEED_DXI) {
                            fSc = fma(v, c, fSc);
                            fSs = fma(v, (T)M_PI * wn * (T)(2 * pos + 1) * sn, fSs);
                        }
                    }
                });
                acc += (double)facc;
                if (NEED_DXI) {
                    Sc += (double)fSc;
                    Ss += (double)fSs;
                }
                if (MODE == 1) Ss += (double)fSs;
            }
        }
        if constexpr (TEAM) {
            // partial sums of all CTAs meet in global double accumulators (zero on entry, re-zeroed by their reader)
            if (!BWD) {
                atomicAdd(gacc + lane, acc);
                if (MODE == 1) atomicAdd(gacc + 32 + lane, Ss);
            } else if (NEED_DXI) {
                atomicAdd(gacc + lane, Sc);
                atomicAdd(gacc + 32 + lane, Ss);
            }
        }
        phase_sync();

        if (MODE == 1) {
            for (int r = gw; r < n; r += GW)
                if (act) ranks[fsw_rowoff(e0 + r, ldr) + k] = (unsigned short)__float_as_int(dst[r * 32 + lane]);
        }
        if (!BWD) {
            if constexpr (TEAM) {
                if (blockIdx.x == 0 && warp == 0) {
                    const double tot = gacc[lane], tots = (MODE == 1) ? gacc[32 + lane] : 0.0;
                    gacc[lane] = 0.0;
                    if (MODE == 1) gacc[32 + lane] = 0.0;
                    if (act) {
                        out[fsw_rowoff(s, ld_out) + out_col0 + k] = ((T)1 + xi) * A0 * (T)tot + bk;
                        if (MODE == 1 && want_dxi)
                            dxi_out[fsw_rowoff(s, ld_dxi) + k] = (T)((double)A0 * tot + (1.0 + xid) * ((double)A0p * tot - (double)A0 * tots));
                    }
                }
            } else {
                red[warp][lane] = acc;
                if (MODE == 1) red2[warp][lane] = Ss;
                __syncthreads();
                if (warp == 0 && act) {
                    double tot = 0.0, tots = 0.0;
#pragma unroll
                    for (int w2 = 0; w2 < W; ++w2) {
                        tot += red[w2][lane];
                        if (MODE == 1) tots += red2[w2][lane];
                    }
                    out[fsw_rowoff(s, ld_out) + out_col0 + k] = ((T)1 + xi) * A0 * (T)tot + bk;
                    if (MODE == 1 && want_dxi)
                        dxi_out[fsw_rowoff(s, ld_dxi) + k] = (T)((double)A0 * tot + (1.0 + xid) * ((double)A0p * tot - (double)A0 * tots));
                }
            }
        } else {
            // dst now holds dL/dp in ORIGINAL row order
            for (int r = gw; r < n; r += GW) {
                if (act) {
                    const T v = dst[r * 32 + lane];
                    if (a.col)
                        atomicAdd(dXp + fsw_rowoff(a.col[e0 + r], a.ldp) + k, v);
                    else
                        dXp[fsw_rowoff(e0 + r, a.ldp) + k] = v;
                    if (dEp) dEp[fsw_rowoff(e0 + r, a.ldp) + k] = v;
                }
            }
            if (NEED_DXI) {
                if constexpr (TEAM) {
                    if (blockIdx.x == 0 && warp == 0) {
                        const double tc = gacc[lane], ts = gacc[32 + lane];
                        gacc[lane] = 0.0;
                        gacc[32 + lane] = 0.0;
                        dxi_acc += (double)gk * ((double)A0 * tc + (1.0 + xid) * ((double)A0p * tc - (double)A0 * ts));
                    }
                } else {
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
        }
        phase_sync();
    }
    if (BWD && NEED_DXI && warp == 0 && act && (!TEAM || blockIdx.x == 0)) atomicAdd(dfreqs + k, dxi_acc);
  }
}

const int kPersistentGrid = 148 * 2;

template <typename T, typename IdxT, int W, int MODE, bool NEED_DXI, bool GLOBAL, bool USE_TABLE, bool TEAM = false>
int launch_medium(const SegArgs<T>& a, int lo, int hi, int cap, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                  const T* g, int64_t ld_g, int64_t g_col0, T* dXp, T* dEp, double* dfreqs, void* scratch,
                  size_t scratch_bytes, unsigned short* ranks, int64_t ldr, T* dxi_out, int64_t ld_dxi, const T* gtab_c,
                  const T* gtab_t, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int64_t cnt = hi - lo;
    int64_t G = GLOBAL ? 1 : cnt * nchunks / (148 * 8);
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    const int64_t nwork = fsw_cdiv(cnt, G) * nchunks;
    const size_t tile = fsw_medium_tile_bytes_t<MODE, USE_TABLE, T, IdxT>(cap);
    auto kern = fsw_medium_kernel<T, IdxT, W, MODE, NEED_DXI, GLOBAL, USE_TABLE, TEAM>;
    int64_t blocks = nwork;
    size_t smem = tile;
    static const char* names[3] = {"fwd_medium_u", "fwdr_medium_u", "bwd_medium_u"};
    const std::string label = std::string(names[MODE]) + std::to_string(cap) + (TEAM ? "_team_f32" : "_f32");
    if constexpr (TEAM) {
        // one tile at a time, the whole machine on it: cooperative launch, every CTA resident (grid-wide barriers)
        const size_t need = tile + 3 * 32 * sizeof(double);
        if (need > scratch_bytes) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small: need >= %zu bytes", need);
        int dev = 0, sms = 0, per_sm = 0;
        FSW_CUDA(cudaGetDevice(&dev));
        FSW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        FSW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, 0));
        if (per_sm < 1) return fsw_fail(FSW_ERR_CUDA, "fsw_medium_kernel (team): no resident CTA");
        const int grid = sms * (per_sm > 2 ? 2 : per_sm);
        FSW_CUDA(cudaMemsetAsync((unsigned char*)scratch + tile, 0, 3 * 32 * sizeof(double), st));
        int seg_lo = lo, seg_hi = hi, Gi = (int)G, nch = nchunks, capi = cap;
        int64_t nw = nwork;
        SegArgs<T> aa = a;
        unsigned char* gs = (unsigned char*)scratch;
        void* args[] = {&aa, &seg_lo, &seg_hi, &Gi, &nch, &capi, &nw, &out, &ld_out, &out_col0, &bias, &g, &ld_g, &g_col0,
                        &dXp, &dEp, &dfreqs, &gs, &ranks, &ldr, &dxi_out, &ld_dxi, &gtab_c, &gtab_t};
        fsw_prof_begin(label.c_str(), st);
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(W * 32), args, 0, st);
        fsw_prof_end(st);
        if (e != cudaSuccess) return fsw_fail(FSW_ERR_CUDA, "fsw_medium_kernel (team): cooperative launch failed: %s", cudaGetErrorString(e));
        fsw_count_launch();
        return FSW_OK;
    }
    if (GLOBAL) {
        smem = 0;
        if (blocks > kPersistentGrid) blocks = kPersistentGrid;
        if ((size_t)blocks * tile > scratch_bytes) blocks = (int64_t)(scratch_bytes / tile);
        if (blocks < 1) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small: need >= %zu bytes", tile);
    } else {
        FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, W * 32, smem, st>>>(a, lo, hi, (int)G, nchunks, cap, nwork, out, ld_out, out_col0, bias, g, ld_g, g_col0,
                                                  dXp, dEp, dfreqs, (unsigned char*)scratch, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_medium_kernel");
    return FSW_OK;
}

template <int MODE, bool NEED_DXI>
int dispatch_medium(const SegArgs<float>& a, int lo, int hi, int cap, float* out, int64_t ld_out, int64_t out_col0,
                    const float* bias, const float* g, int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs,
                    void* scratch, size_t scratch_bytes, unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi,
                    const float* gtab_c, const float* gtab_t, cudaStream_t st) {
#define FSW_MED(IDX, W, GLOBAL, TABLE) \
    launch_medium<float, IDX, W, MODE, NEED_DXI, GLOBAL, TABLE>(a, lo, hi, cap, out, ld_out, out_col0, bias, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch, scratch_bytes, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st)
    constexpr bool TAB = MODE == 0;  // keys-only forward: coefficient table in shared memory; training: global tables
    if (cap <= 128) return FSW_MED(unsigned short, 4, false, TAB);
    if (cap <= 256) return FSW_MED(unsigned short, 8, false, TAB);
    if (cap <= 512) return FSW_MED(unsigned short, 16, false, TAB);
#define FSW_MED_TEAM(IDX) \
    launch_medium<float, IDX, 16, MODE, NEED_DXI, true, false, true>(a, lo, hi, cap, out, ld_out, out_col0, bias, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch, scratch_bytes, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st)
    // hubs: few tiles, each too long for one CTA -> the whole grid works on one tile at a time
    static const bool team_ok = getenv("FSW_NO_TEAM") == nullptr;
    const int64_t tiles = (int64_t)(hi - lo) * ((a.K + 31) / 32);
    if (team_ok && cap >= FSW_TEAM_MIN_CAP && tiles < 2 * 148) {
        if (cap <= 32768) return FSW_MED_TEAM(unsigned short);
        return FSW_MED_TEAM(int);
    }
    if (cap <= 32768) return FSW_MED(unsigned short, 16, true, false);
    return FSW_MED(int, 16, true, false);
#undef FSW_MED_TEAM
#undef FSW_MED
}

}  // namespace

// global scratch per CTA (0: shared-memory tile).  mode: 0 forward, 1 forward + ranks, 2 backward (re-sort)
size_t fsw_medium_tile_bytes(int cap, int mode) {
    if (cap <= 512) return 0;
    const size_t idx = cap <= 32768 ? 2 : 4;
    return (size_t)cap * 32 * (2 * 4 + (mode >= 1 ? 2 * idx : 0));
}

int fsw_medium_grid() { return kPersistentGrid; }

// uniform-weight fp32 segments order[lo, hi) whose size class is `cap` (>= 128); ranks != NULL records positions
int fsw_medium_forward_f32(const SegArgs<float>& a, int lo, int hi, int cap, float* out, int64_t ld_out, int64_t out_col0,
                           const float* bias, void* scratch, size_t scratch_bytes, unsigned short* ranks, int64_t ldr,
                           float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t, cudaStream_t st) {
    if (ranks != nullptr && cap <= 32768)
        return dispatch_medium<1, false>(a, lo, hi, cap, out, ld_out, out_col0, bias, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                         scratch, scratch_bytes, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st);
    return dispatch_medium<0, false>(a, lo, hi, cap, out, ld_out, out_col0, bias, nullptr, 0, 0, nullptr, nullptr, nullptr,
                                     scratch, scratch_bytes, nullptr, 0, nullptr, 0, nullptr, nullptr, st);
}

int fsw_medium_backward_f32(const SegArgs<float>& a, int lo, int hi, int cap, const float* g, int64_t ld_g, int64_t g_col0,
                            float* dXp, float* dEp, double* dfreqs, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (dfreqs != nullptr)
        return dispatch_medium<2, true>(a, lo, hi, cap, nullptr, 0, 0, nullptr, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch,
                                        scratch_bytes, nullptr, 0, nullptr, 0, nullptr, nullptr, st);
    return dispatch_medium<2, false>(a, lo, hi, cap, nullptr, 0, 0, nullptr, g, ld_g, g_col0, dXp, dEp, dfreqs, scratch,
                                     scratch_bytes, nullptr, 0, nullptr, 0, nullptr, nullptr, st);
}
