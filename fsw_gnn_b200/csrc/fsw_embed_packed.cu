// K2, uniform-weight fp32 forward for segments of 33..256 elements: the whole sort in registers, packed keys,
// L cooperating lanes per slice.
//
//   fsw_coop_fwd_kernel<R, L, HAS_COL, SAVE_RANK>     (R, L) = (12|16, 4) n <= 48|64, (12|16, 8) n <= 96|128,
//                                                      (12|16, 16) n <= 192|256, (24|32, 16) n <= 384|512;
//                                                      dense batches also (32, 32) n <= 1024.  Run lengths that are
//                                                      not powers of two (12, 24) re-sort the lane's bitonic run with
//                                                      the sorting network instead of half-cleaners: ~20 % more
//                                                      comparators per slot, 25 % fewer slots to carry.
//     A (segment, slice) is sorted by L lanes of one warp holding R elements each; a warp works on SW = 32 / L
//     consecutive slices of one segment.
//     * gather, row-wise: the SW slices of a source row are one aligned 16/32-byte piece, read by one or two
//       lanes with a vector load (32 or 16 rows per instruction) and parked in shared memory as [element][slice].
//     * packed keys: the network sorts ONE 32-bit word per element - the order-preserving integer image of the
//       key with its low log2(R L) bits replaced by the element index.  A comparator is two integer min/max
//       instructions instead of a compare and four selects, and no index array travels with the keys.
//     * each lane sorts its R words with a merge-exchange network, then the lanes are merged with bitonic
//       merges whose cross-lane steps are shuffles (static register indices).  The straight-line code of one
//       segment stays below the 32 KB L1.5 instruction cache (R = 16: ~1500 instructions) - a monolithic
//       128-element network per thread, and R = 32 variants, stalled most cycles on instruction fetch
//       (profiles/r1/README.md).
//     * the dropped low bits are restored exactly: adjacent sorted positions whose truncated keys coincide
//       (min over XORs per block of 8 positions, voted across the warp) are bubble-ordered by their full keys
//       from shared memory.  Equal full keys keep element order, as a stable sort would.  Padding slots get
//       huge finite keys with pairwise different upper bits, so they sort last and never look like ties.
//     * Fourier coefficients cos(pi xi (2j+1)/n) and their d/dxi companions come from global tables blocked as
//       [n][position/4][slice][4] (fsw_build_fwd_tables): a lane reads its 16 positions as four float4s and the
//       SW lanes of a run share cache lines.  Entries beyond n are zero, which also cancels the padding keys.
//       The L partial sums of a slice are combined with shuffles.
//     * SAVE_RANK: the slot of each consumed key is overwritten by its sorted position; the ranks leave row-wise,
//       one vector store of SW uint16 per element, for the sort-free backward (fsw_embed_small.cu).
//   Bound (ncu, profiles/r1): ALU pipe 66-74 % and L1 72-82 % busy at once; DRAM < 20 %.
#include <stdlib.h>

#include "fsw_sortnet.cuh"

namespace {

constexpr int FSW_PK_BLOCK = 8;  // register positions per near-tie detection block

// compile-time loop: f(std::integral_constant<int, I>) for I = 0..N-1 (indices stay constants whatever the size)
template <typename F, int... Is>
__device__ __forceinline__ void fsw_static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void fsw_static_for(F&& f) {
    fsw_static_for_impl(f, std::make_integer_sequence<int, N>{});
}

constexpr int fsw_clog2(int x) {
    int b = 0;
    while ((1 << b) < x) ++b;
    return b;
}

struct PkMeta {
    int s;
    int n;
    int64_t e0;
};

__device__ __forceinline__ int fsw_pk_order(const SegArgs<float>& a, int q, int last) {
    const int qq = (q < last) ? q : last - 1;
    return a.order ? __ldg(a.order + qq) : qq;
}

__device__ __forceinline__ void fsw_pk_range(const SegArgs<float>& a, int s, int64_t& e0, int& n) {
    if (a.rowptr) {
        const int lo = __ldg(a.rowptr + s);
        const int hi = __ldg(a.rowptr + s + 1);
        e0 = lo;
        n = hi - lo;
    } else {
        e0 = (int64_t)s * a.n_fixed;
        n = (int)a.n_fixed;
    }
}

template <int NC, bool HAS_COL>
__device__ __forceinline__ void fsw_pk_cols(const int32_t* __restrict__ col, int64_t ebase, int cnt, int lane, int (&c)[NC]) {
#pragma unroll
    for (int m = 0; m < NC; ++m) {
        c[m] = 0;
        if (HAS_COL) c[m] = __ldg(col + ebase + min(lane + 32 * m, cnt - 1));
    }
}

// SW consecutive 32-bit words (SW = 1, 2, 4, 8, 16) at a 4 SW-byte aligned address
template <int SW>
__device__ __forceinline__ void fsw_ldg_words(const float* p, float (&v)[SW]) {
    if constexpr (SW == 1) {
        v[0] = __ldg(p);
    } else if constexpr (SW == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < SW / 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
            v[4 * q] = t.x;
            v[4 * q + 1] = t.y;
            v[4 * q + 2] = t.z;
            v[4 * q + 3] = t.w;
        }
    }
}

template <int SW>
__device__ __forceinline__ void fsw_sts_words(float* p, const float (&v)[SW]) {
    if constexpr (SW == 1) {
        *p = v[0];
    } else if constexpr (SW == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < SW / 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// SW sorted positions (ints in shared memory) -> SW uint16 ranks, one vector store
template <int SW>
__device__ __forceinline__ void fsw_store_ranks(unsigned short* dst, const int* src) {
    if constexpr (SW == 1) {
        *dst = (unsigned short)src[0];
    } else if constexpr (SW == 2) {
        const int2 t = *reinterpret_cast<const int2*>(src);
        *reinterpret_cast<unsigned*>(dst) = (unsigned)t.x | ((unsigned)t.y << 16);
    } else if constexpr (SW == 4) {
        const int4 t = *reinterpret_cast<const int4*>(src);
        *reinterpret_cast<uint2*>(dst) = make_uint2((unsigned)t.x | ((unsigned)t.y << 16), (unsigned)t.z | ((unsigned)t.w << 16));
    } else {
#pragma unroll
        for (int q = 0; q < SW / 8; ++q) {
            const int4 t = reinterpret_cast<const int4*>(src)[2 * q];
            const int4 u = reinterpret_cast<const int4*>(src)[2 * q + 1];
            reinterpret_cast<uint4*>(dst)[q] = make_uint4((unsigned)t.x | ((unsigned)t.y << 16), (unsigned)t.z | ((unsigned)t.w << 16),
                                                          (unsigned)u.x | ((unsigned)u.y << 16), (unsigned)u.z | ((unsigned)u.w << 16));
        }
    }
}

// compare-exchange of two packed words held by this lane
#define FSW_PK_CMPX(x, y)        \
    {                            \
        const int lo_ = min(x, y); \
        y = max(x, y);           \
        x = lo_;                 \
    }

// asynchronous copy global -> shared (LDGSTS), BYTES = 4, 8 or 16 at a BYTES-aligned address; completes before cp.async.wait_all returns
template <int BYTES>
__device__ __forceinline__ void fsw_cp_async(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void fsw_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// a bitonic run s[R0 .. R0 + LEN) in ascending order: half-cleaners while the length is even (valid for every even length,
// not only powers of two), a sorting network on what remains (24 comparators for 12 elements instead of the 41 of a full
// network, 60 instead of 127 for 24)
template <int R0, int LEN, typename CE>
__device__ __forceinline__ void fsw_bitonic_merge(CE&& ce) {
    if constexpr (LEN <= 1) {
    } else if constexpr (LEN % 2 == 0) {
        fsw_static_for<LEN / 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            ce(R0 + i, R0 + i + LEN / 2);
        });
        fsw_bitonic_merge<R0, LEN / 2>(ce);
        fsw_bitonic_merge<R0 + LEN / 2, LEN / 2>(ce);
    } else {
        fsw_sort_network<LEN>([&](int i, int l) { ce(R0 + i, R0 + l); });
    }
}

// Per-class policy, from the measurements in profiles/r2/README.md (the kernels are bound by the shared-memory data pipe -
// shuffles cost two wavefronts each - and by the integer ALU; DRAM is < 25 % busy):
//   gather: synchronous (all loads issued before the first use) except R = 16, where the asynchronous copy (LDGSTS) of the
//           next segment into a second key buffer is as fast and leaves registers for the sort;
//   coefficients: staged per lane in shared memory (reloaded when n changes) for R = 12 and for R >= 24 with up to 8 lanes per
//           slice (-5..11 %); through L1 for R = 16 (one CTA less per SM otherwise) and for 16 / 32 lanes per slice (+16 %);
//   registers: 64 (8 CTAs) at R = 12, 72 (7 CTAs) at R = 16, 96 (5 CTAs) at R = 24, 128 (4 CTAs) at R = 32 - R = 24 at 80
//           registers and R = 32 at 96 spill ~150 bytes and lose 5-15 %.
#ifndef FSW_COOP_ASYNC
#define FSW_COOP_ASYNC(R) ((R) == 16)
#endif
#ifndef FSW_COOP_TABS
#define FSW_COOP_TABS(R, L) ((R) == 12 || ((R) >= 24 && (L) <= 8))
#endif
#ifndef FSW_COOP_MINB
#define FSW_COOP_MINB(R) ((R) <= 12 ? 8 : ((R) <= 16 ? 7 : ((R) <= 24 ? 5 : 4)))
#endif

// shared memory of one warp, in floats: key buffers [NBUF][R L][32 / L] + the lane-private coefficient rows [TABF][R / 4][32] float4
template <int R, int L, bool SAVE_RANK, bool CLOUD>
struct FswCoopSmem {
    static constexpr bool ASYNC = FSW_COOP_ASYNC(R) && !CLOUD;
    static constexpr int NBUF = ASYNC ? 2 : 1;                             // graphs: the gather of the next segment lands while this one is sorted
    static constexpr int TABF = (FSW_COOP_TABS(R, L) && !CLOUD) ? (SAVE_RANK ? 2 : 1) : 0;  // tables staged in shared memory (reloaded when n changes)
    static constexpr int KEYS = R * 32;
    static constexpr int PER_WARP = (NBUF + TABF) * KEYS;
};

template <int R, int L, bool HAS_COL, bool SAVE_RANK, bool CLOUD>
__global__ void __launch_bounds__(128, FSW_COOP_MINB(R)) fsw_coop_fwd_kernel(
    SegArgs<float> a, int seg_lo, int seg_hi, int G, int nchunks, float* __restrict__ out, int64_t ld_out, int64_t out_col0,
    const float* __restrict__ bias, unsigned short* __restrict__ ranks, int64_t ldr, float* __restrict__ dxi_out, int64_t ld_dxi,
    const float* __restrict__ gtab_c, const float* __restrict__ gtab_t, int tab_n0, int tab_ld4) {
    static_assert(R % 4 == 0 && R >= 4, "a lane reads its table positions as float4s");
    static_assert((L & (L - 1)) == 0 && L >= 4 && L <= 32, "lanes per slice: power of two; 32 / L <= 8 slices tile the padded width");
    using SM = FswCoopSmem<R, L, SAVE_RANK, CLOUD>;
    constexpr int SW = 32 / L;            // slices per warp
    constexpr int NS = R * L;             // element slots per (segment, slice)
    constexpr int NC = (NS + 31) / 32;    // column-id registers per lane (element 32 m + lane)
    constexpr bool PREFETCH_COLS = NC <= 8;
    constexpr int VW = SW < 4 ? SW : 4;   // gather: floats per lane,
    constexpr int LPR = SW / VW;          //         lanes per row,
    constexpr int RPI = 32 / LPR;         //         rows per load instruction
    constexpr int IDXB = fsw_clog2(NS);   // low bits that carry the element index
    constexpr int IMASK = (1 << IDXB) - 1;
    constexpr int LOGL = fsw_clog2(L);
    constexpr bool TABS = SM::TABF > 0;
    constexpr bool ASYNC = SM::ASYNC;
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / SW;   // which run of the slice this lane holds
    const int sl = lane % SW;  // slice within the warp
    float* kbase = reinterpret_cast<float*>(fsw_smem_raw) + (size_t)warp * SM::PER_WARP;  // key buffers [NBUF][NS][SW] of this warp
    float4* tabC = reinterpret_cast<float4*>(kbase + SM::NBUF * SM::KEYS) + lane;           // this lane's coefficients: tabC[32 (i4 / 4)]
    float4* tabT = tabC + (R / 4) * 32;

    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first64 = (int64_t)seg_lo + item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k0 = chunk * SW;  // the warp's slices k0 .. k0+SW-1 (inside the padded width: SW divides 8 | ldp, or ldp % SW == 0)
    const int k = k0 + sl;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const float xi = __ldg(a.freqs + kk);
    const double xid = (double)xi;
    const float bk = (bias != nullptr) ? __ldg(bias + kk) : 0.f;
    const int ldp = (int)a.ldp;
    const float* xp0 = a.Xp + k0;
    const float* ep0 = a.Ep ? a.Ep + k0 : nullptr;
    const bool want_dxi = SAVE_RANK && dxi_out != nullptr;
    // this lane's table entries: [n][position / 4][kk][position % 4], positions g R .. g R + R - 1
    const int tstride = a.K * 4;  // floats between consecutive position blocks
    const float* tck = gtab_c + (int64_t)(g * (R / 4)) * tstride + kk * 4;
    const float* ttk = gtab_t + (int64_t)(g * (R / 4)) * tstride + kk * 4;

    // point-cloud mode: the slices this lane projects onto (theta rows k0 + part .. + VW - 1), kept in registers
    // (a separate template instance: the registers and branches of this mode cost the graph kernels 12 % when they shared one)
    float th[CLOUD ? VW : 1][4];
#pragma unroll
    for (int j = 0; j < (CLOUD ? VW : 1); ++j)
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) th[j][dd] = 0.f;
    if constexpr (CLOUD) {
        const int part0 = (lane % LPR) * VW;
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            const int ks = min(k0 + part0 + j, a.K - 1);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd)
                if (dd < a.proj_d) th[j][dd] = __ldg(a.projTheta + fsw_rowoff(ks, a.proj_ldt) + dd);
        }
    }

    // ---- gather of one segment into a key buffer [element][slice], row-wise: LPR adjacent lanes take the SW slices of one row
    //      (one sector-aligned piece each), 32/LPR rows per instruction.  Graphs: asynchronous copies (LDGSTS) - no register
    //      staging, nothing waits here; the consumer canonicalises -0 to +0 when it packs (zeros tie, torch.sort compares
    //      values).  Slots beyond n get a huge finite key whose upper bits differ per slot (their own tie groups).
    auto fill = [&](float* buf, const PkMeta& m, const int (&cc)[NC]) {
        float vv[(!CLOUD && !ASYNC) ? NS / RPI : 1][VW];
#pragma unroll
        for (int t = 0; t < NS / RPI; ++t) {
            const int e = t * RPI + lane / LPR;
            const int part = (lane % LPR) * VW;
            float* dst = buf + e * SW + part;
            int row = 0;
            if (HAS_COL) row = (LPR == 1) ? cc[t] : __shfl_sync(FSW_FULL, cc[(t * RPI) >> 5], ((t * RPI) & 31) + lane / LPR);
            if constexpr (CLOUD) {
                // keys on the fly: <x_e, theta_k>, same FMA order as the projection kernel (fsw_project_small_kernel), so the
                // keys are bit-identical to a materialised projection; the point rows are read contiguously (clamped row:
                // the loads carry no predicate and are issued back to back)
                const float* xr = a.projX + fsw_rowoff(m.e0 + min(e, m.n - 1), a.proj_d);
                float x[4];
#pragma unroll
                for (int dd = 0; dd < 4; ++dd) x[dd] = dd < a.proj_d ? __ldg(xr + dd) : 0.f;
                float v[VW];
#pragma unroll
                for (int j = 0; j < VW; ++j) {
                    float acc = 0.f;
#pragma unroll
                    for (int dd = 0; dd < 4; ++dd)
                        if (dd < a.proj_d) acc = fmaf(x[dd], th[CLOUD ? j : 0][dd], acc);
                    v[j] = (e < m.n) ? acc : __int_as_float(0x7f000000 | (e << IDXB));
                }
                fsw_sts_words<VW>(dst, v);
            } else if constexpr (ASYNC) {
                if (e < m.n) {
                    const int64_t r64 = HAS_COL ? (int64_t)row : m.e0 + e;
                    const float* src = xp0 + fsw_rowoff(r64, ldp) + part;
                    if (ep0 == nullptr) {
                        fsw_cp_async<VW * 4>(dst, src);
                    } else {  // edge features: per-slot additive projection (rare path, synchronous)
                        float v[VW], w[VW];
                        fsw_ldg_words<VW>(src, v);
                        fsw_ldg_words<VW>(ep0 + fsw_rowoff(m.e0 + e, ldp) + part, w);
#pragma unroll
                        for (int j = 0; j < VW; ++j) v[j] += w[j];
                        fsw_sts_words<VW>(dst, v);
                    }
                } else {
                    float v[VW];
                    const float pad = __int_as_float(0x7f000000 | (e << IDXB));
#pragma unroll
                    for (int j = 0; j < VW; ++j) v[j] = pad;
                    fsw_sts_words<VW>(dst, v);
                }
            } else {
                // synchronous: the loads carry no predicate (column ids are clamped, so every lane holds a valid row) and are
                // all issued before the first one is consumed
                const int64_t r64 = HAS_COL ? (int64_t)row : m.e0 + min(e, m.n - 1);
                fsw_ldg_words<VW>(xp0 + fsw_rowoff(r64, ldp) + part, vv[t]);
            }
        }
        if constexpr (!CLOUD && !ASYNC) {
#pragma unroll
            for (int t = 0; t < NS / RPI; ++t) {
                const int e = t * RPI + lane / LPR;
                const int part = (lane % LPR) * VW;
                if (ep0 != nullptr && e < m.n) {  // edge features: per-slot additive projection (rare path)
                    float w[VW];
                    fsw_ldg_words<VW>(ep0 + fsw_rowoff(m.e0 + e, ldp) + part, w);
#pragma unroll
                    for (int j = 0; j < VW; ++j) vv[t][j] += w[j];
                }
                const float pad = __int_as_float(0x7f000000 | (e << IDXB));
#pragma unroll
                for (int j = 0; j < VW; ++j) vv[t][j] = (e < m.n) ? vv[t][j] : pad;
                fsw_sts_words<VW>(buf + e * SW + part, vv[t]);
            }
        }
    };

    // software pipeline over segments: order four ahead, row range three ahead, column ids two ahead, keys one ahead
    PkMeta cur, nx1, nx2;
    int s3;
    int c[NC] = {};
    cur.s = fsw_pk_order(a, first, last);
    fsw_pk_range(a, cur.s, cur.e0, cur.n);
    nx1.s = fsw_pk_order(a, first + 1, last);
    fsw_pk_range(a, nx1.s, nx1.e0, nx1.n);
    nx2.s = fsw_pk_order(a, first + 2, last);
    fsw_pk_range(a, nx2.s, nx2.e0, nx2.n);
    s3 = fsw_pk_order(a, first + 3, last);
    if constexpr (!CLOUD) fsw_pk_cols<NC, HAS_COL>(a.col, cur.e0, cur.n, lane, c);
    if constexpr (ASYNC) {
        fill(kbase, cur, c);
        fsw_pk_cols<NC, HAS_COL>(a.col, nx1.e0, nx1.n, lane, c);
    }

    int n_prev = -1;
    int pb = 0;  // key buffer of the current segment
    float A = 0.f, A0 = 0.f, A0p = 0.f;
    const float* tc = tck;
    const float* tt = ttk;

#pragma unroll 1
    for (int q = first; q < last; ++q) {
        const int n = cur.n;
        float* fkw = kbase + pb * SM::KEYS;       // full keys [NS][SW] of this segment
        float* fks = fkw + sl;                    // column of this lane's slice
        const float* fkl = fkw + lane;            // element i*L+g of this lane: fkl[i*32]
        int* fksi = reinterpret_cast<int*>(fks);
        int cn[PREFETCH_COLS ? NC : 1];
        if constexpr (ASYNC) {
            fsw_cp_async_wait_all();
            __syncwarp();
            // the next segment's rows start to arrive while this one is sorted
            if (q + 1 < last) fill(kbase + (pb ^ 1) * SM::KEYS, nx1, c);
            if constexpr (PREFETCH_COLS) fsw_pk_cols<NC, HAS_COL>(a.col, nx2.e0, nx2.n, lane, cn);
        } else {
            fill(fkw, cur, c);
            if constexpr (!CLOUD && PREFETCH_COLS) fsw_pk_cols<NC, HAS_COL>(a.col, nx1.e0, nx1.n, lane, cn);
        }
        PkMeta nx3;
        nx3.s = s3;
        fsw_pk_range(a, s3, nx3.e0, nx3.n);
        const int s4 = fsw_pk_order(a, q + 4, last);

        if (n != n_prev) {
            const double u = xid / (double)n;
            const float wn = (float)(1.0 / (double)n);
            fsw_amplitude<float, SAVE_RANK>(u, wn, xi, A0, A0p);
            A = (1.f + xi) * A0;
            tc = tck + (int64_t)(n - tab_n0) * tab_ld4 * tstride;
            tt = ttk + (int64_t)(n - tab_n0) * tab_ld4 * tstride;
            n_prev = n;
            if constexpr (TABS) {
                // this lane's R coefficients (and d/dxi companions) of size class n: shared memory, lane-private rows
#pragma unroll
                for (int i4 = 0; i4 < R; i4 += 4) {
                    float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), tv = cv;
                    const bool live = g * R + i4 < n;
                    if (live) cv = __ldg(reinterpret_cast<const float4*>(tc + fsw_rowoff(i4 / 4, tstride)));
                    if (want_dxi && live) tv = __ldg(reinterpret_cast<const float4*>(tt + fsw_rowoff(i4 / 4, tstride)));
                    tabC[(i4 / 4) * 32] = cv;
                    if (SAVE_RANK) tabT[(i4 / 4) * 32] = tv;
                }
            }
        }
        if constexpr (!ASYNC) __syncwarp();

        // ---- pack: lane (g, sl) takes elements e = i L + g of its slice: sortable image | element index ----
        int s[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int b = __float_as_int(fkl[i * 32] + 0.0f);  // = fkw[(i L + g) SW + sl]; -0 becomes +0
            const int t = b ^ ((b >> 31) & 0x7fffffff);
            s[i] = ((t & ~IMASK) | g) + i * L;
        }
        // ---- sort the lane's run, then merge the runs of the L lanes ----
        fsw_sort_network<R>([&](int i, int l) { FSW_PK_CMPX(s[i], s[l]); });
        if constexpr (L <= 16 && R <= 16) {
            fsw_static_for<LOGL>([&](auto lc) {
                constexpr int lv = decltype(lc)::value + 1;  // merge groups of 2^(lv-1) lanes into groups of 2^lv
                {
                    // lane g meets lane g ^ (2^lv - 1), position i against R-1-i of the partner
                    const bool upper = (g >> (lv - 1)) & 1;
                    constexpr int xm = ((1 << lv) - 1) * SW;
                    fsw_static_for<R / 2>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        const int ya = __shfl_xor_sync(FSW_FULL, s[R - 1 - i], xm);
                        const int yb = __shfl_xor_sync(FSW_FULL, s[i], xm);
                        s[i] = upper ? max(s[i], ya) : min(s[i], ya);
                        s[R - 1 - i] = upper ? max(s[R - 1 - i], yb) : min(s[R - 1 - i], yb);
                    });
                }
                fsw_static_for<lv - 1>([&](auto dc) {
                    constexpr int d = 1 << (lv - 2 - decltype(dc)::value);  // lane distance 2^(lv-2) .. 1, same position
                    const bool upper = (g & d) != 0;
    #pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int y = __shfl_xor_sync(FSW_FULL, s[i], d * SW);
                        s[i] = upper ? max(s[i], y) : min(s[i], y);
                    }
                });
                // the lane now holds a bitonic run
                fsw_bitonic_merge<0, R>([&](int i, int l) { FSW_PK_CMPX(s[i], s[l]); });
            });
        } else {
            // many lanes per slice or long runs: the same steps as loops over the merge level and the lane distance
            // (runtime shuffle masks), so that the code stays inside the instruction cache
#pragma unroll 1
            for (int lv = 1; lv <= LOGL; ++lv) {
                {
                    const bool upper = (g >> (lv - 1)) & 1;
                    const int xm = ((1 << lv) - 1) * SW;
                    fsw_static_for<R / 2>([&](auto ic) {
                        constexpr int i = decltype(ic)::value;
                        const int ya = __shfl_xor_sync(FSW_FULL, s[R - 1 - i], xm);
                        const int yb = __shfl_xor_sync(FSW_FULL, s[i], xm);
                        s[i] = upper ? max(s[i], ya) : min(s[i], ya);
                        s[R - 1 - i] = upper ? max(s[R - 1 - i], yb) : min(s[R - 1 - i], yb);
                    });
                }
#pragma unroll 1
                for (int d = (lv >= 2) ? (1 << (lv - 2)) : 0; d >= 1; d >>= 1) {
                    const bool upper = (g & d) != 0;
                    const int xm = d * SW;
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        const int y = __shfl_xor_sync(FSW_FULL, s[i], xm);
                        s[i] = upper ? max(s[i], y) : min(s[i], y);
                    }
                }
                fsw_bitonic_merge<0, R>([&](int i, int l) { FSW_PK_CMPX(s[i], s[l]); });
            }
        }
        // sorted position of s[i] in lane g: p = g R + i

        // ---- exact order inside groups of equal truncated keys ----
        constexpr int NB = (R + FSW_PK_BLOCK - 1) / FSW_PK_BLOCK;
        unsigned wflags = 0;  // warp-uniform: block b holds an adjacent pair with equal truncated keys in some lane
        int ynext = 0;        // first word of the next lane (position R of this lane)
        if (L > 1) {
            ynext = __shfl_down_sync(FSW_FULL, s[0], SW);
            if (g == L - 1) ynext = s[R - 1] ^ (IMASK + 1);  // no successor: never a tie
        }
        fsw_static_for<NB>([&](auto bc) {
            constexpr int b = decltype(bc)::value;
            constexpr int j0 = b * FSW_PK_BLOCK;
            constexpr int cnt = (R - j0 < FSW_PK_BLOCK) ? R - j0 : FSW_PK_BLOCK;
            unsigned mn = 0xffffffffu;
            fsw_static_for<cnt>([&](auto jc) {
                constexpr int j = j0 + decltype(jc)::value;
                if constexpr (j < R - 1)
                    mn = min(mn, (unsigned)(s[j] ^ s[j + 1]));
                else if constexpr (L > 1)
                    mn = min(mn, (unsigned)(s[j] ^ ynext));
            });
            if (__any_sync(FSW_FULL, mn <= (unsigned)IMASK)) wflags |= 1u << b;
        });
        while (wflags != 0) {
            bool swapped = false;
            fsw_static_for<NB>([&](auto bc) {
                constexpr int b = decltype(bc)::value;
                constexpr int j0 = b * FSW_PK_BLOCK;
                constexpr int cnt = (R - j0 < FSW_PK_BLOCK) ? R - j0 : FSW_PK_BLOCK;
                if (wflags & (1u << b)) {
                    fsw_static_for<cnt>([&](auto jc) {
                        constexpr int j = j0 + decltype(jc)::value;
                        if constexpr (j < R - 1) {
                            const int x = s[j], y = s[j + 1];
                            if ((unsigned)(x ^ y) <= (unsigned)IMASK) {
                                const float ka = fks[(x & IMASK) * SW];
                                const float kb = fks[(y & IMASK) * SW];
                                if (ka > kb) {
                                    s[j] = y;
                                    s[j + 1] = x;
                                    swapped = true;
                                }
                            }
                        } else if constexpr (L > 1) {
                            // pair across the lane boundary: both lanes evaluate the same exchange
                            const int yn = __shfl_down_sync(FSW_FULL, s[0], SW);      // successor of my last word
                            const int xp = __shfl_up_sync(FSW_FULL, s[R - 1], SW);    // predecessor of my first word
                            int new_last = s[R - 1], new_first = s[0];
                            if (g < L - 1 && (unsigned)(s[R - 1] ^ yn) <= (unsigned)IMASK) {
                                const float ka = fks[(s[R - 1] & IMASK) * SW];
                                const float kb = fks[(yn & IMASK) * SW];
                                if (ka > kb) {
                                    new_last = yn;
                                    swapped = true;
                                }
                            }
                            if (g > 0 && (unsigned)(xp ^ s[0]) <= (unsigned)IMASK) {
                                const float ka = fks[(xp & IMASK) * SW];
                                const float kb = fks[(s[0] & IMASK) * SW];
                                if (ka > kb) {
                                    new_first = xp;
                                    swapped = true;
                                }
                            }
                            s[R - 1] = new_last;
                            s[0] = new_first;
                        }
                    });
                }
            });
            if (!__any_sync(FSW_FULL, swapped)) break;
        }

        // ---- Fourier sums over sorted positions; ranks replace the consumed keys ----
        // Table entries are zero beyond n, so the huge keys of padding slots (positions >= n) contribute exactly 0.
        float acc = 0.f, acc2 = 0.f;
        const int p0 = g * R;
#pragma unroll
        for (int i4 = 0; i4 < R; i4 += 4) {
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), tv = cv;
            if constexpr (TABS) {
                cv = tabC[(i4 / 4) * 32];
                if (want_dxi) tv = tabT[(i4 / 4) * 32];
            } else {
                const bool live = p0 + i4 < n;
                if (live) cv = __ldg(reinterpret_cast<const float4*>(tc + fsw_rowoff(i4 / 4, tstride)));
                if (want_dxi && live) tv = __ldg(reinterpret_cast<const float4*>(tt + fsw_rowoff(i4 / 4, tstride)));
            }
            const float cq[4] = {cv.x, cv.y, cv.z, cv.w};
            const float tq[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = s[i4 + j] & IMASK;
                const float key = fks[idx * SW];
                acc = fmaf(key, cq[j], acc);
                acc2 = fmaf(key, tq[j], acc2);
                if (SAVE_RANK) fksi[idx * SW] = p0 + i4 + j;  // the consumed key's slot now holds the element's rank
            }
        }
#pragma unroll
        for (int m = L / 2; m >= 1; m >>= 1) {
            acc += __shfl_xor_sync(FSW_FULL, acc, m * SW);
            if (want_dxi) acc2 += __shfl_xor_sync(FSW_FULL, acc2, m * SW);
        }
        if (act && g == 0) {
            out[fsw_rowoff(cur.s, ld_out) + out_col0 + k] = A * acc + bk;
            if (want_dxi) dxi_out[fsw_rowoff(cur.s, ld_dxi) + k] = A0 * acc + (1.f + xi) * (A0p * acc - A0 * acc2);
        }
        __syncwarp();
        if constexpr (SAVE_RANK) {
            if (CLOUD) {
                // slice-major ranks [S][K][n]: consecutive lanes store consecutive elements of one slice
#pragma unroll
                for (int j = 0; j < SW; ++j) {
                    if (k0 + j < a.K) {
                        unsigned short* rt = ranks + ((int64_t)cur.s * a.K + (k0 + j)) * n;
#pragma unroll
                        for (int t = 0; t < NC; ++t) {
                            const int e = t * 32 + lane;
                            if (e < n) rt[e] = (unsigned short)reinterpret_cast<const int*>(fkw)[e * SW + j];
                        }
                    }
                }
            } else {
                // row-wise again: lane l writes the SW ranks of elements 32 t + l with one vector store; one 64-bit base
                // per segment, 32-bit offsets per row
                char* rbase = reinterpret_cast<char*>(ranks + k0) + 2 * fsw_rowoff(cur.e0, ldr);
                const unsigned rstep = 2u * (unsigned)ldr;
#pragma unroll
                for (int t = 0; t < NC; ++t) {
                    const int e = t * 32 + lane;
                    if (e < n) fsw_store_ranks<SW>(reinterpret_cast<unsigned short*>(rbase + (unsigned)e * rstep), reinterpret_cast<const int*>(fkw) + e * SW);
                }
            }
            __syncwarp();  // the gather after next overwrites the slots
        }
        // rotate the pipeline
        cur = nx1;
        nx1 = nx2;
        nx2 = nx3;
        s3 = s4;
        if constexpr (!CLOUD) {
            if constexpr (PREFETCH_COLS) {
#pragma unroll
                for (int m = 0; m < NC; ++m) c[m] = cn[m];
            } else {
                fsw_pk_cols<NC, HAS_COL>(a.col, ASYNC ? nx1.e0 : cur.e0, ASYNC ? nx1.n : cur.n, lane, c);
            }
            if constexpr (ASYNC) pb ^= 1;
        }
    }
}

// forward tables, blocked by 4 positions with the slice index inside:
//   tab[((n (LD/4) + j/4) K + k) 4 + j%4] = cos(pi xi_k (2j+1)/n) (and its d/dxi companion) for j < n, zero for j >= n.
// The 32/L lanes that hold the same positions of adjacent slices read adjacent float4s: one cache line per lane group.
__global__ void __launch_bounds__(256) fsw_build_fwd_tables_kernel(const float* __restrict__ freqs, int K, int n0, int ld4,
                                                                   float* __restrict__ tab_c, float* __restrict__ tab_t) {
    const int n = n0 + blockIdx.x;
    const int jb = blockIdx.y;
    if (n < 1 || jb * 4 >= n) return;   // position blocks beyond n are never read (the consumers test p < n per block of 4)
    for (int idx = threadIdx.x; idx < K * 4; idx += blockDim.x) {
        const int k = idx >> 2, j = jb * 4 + (idx & 3);
        float c = 0.f, t = 0.f;
        if (j < n) {
            const double u = (double)freqs[k] / (double)n;
            const float wn = (float)(1.0 / (double)n);
            const float rr = Num<float>::reduce(u * (double)(2 * j + 1));
            c = cospif(rr);
            t = (float)M_PI * wn * (float)(2 * j + 1) * sinpif(rr);
        }
        const int64_t at = ((int64_t)(n - n0) * ld4 + jb) * K * 4 + idx;
        tab_c[at] = c;
        tab_t[at] = t;
    }
}

template <int R, int L, bool HAS_COL, bool SAVE_RANK, bool CLOUD = false>
int launch_coop_fwd(const SegArgs<float>& a, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0, const float* bias,
                    unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t,
                    int tab_n0, int tab_ld4, cudaStream_t st) {
    constexpr int SW = 32 / L;
    const int nchunks = (a.K + SW - 1) / SW;
    int64_t G = (int64_t)(hi - lo) * nchunks / (148 * 64);
    if (G < 1) G = 1;
    if (G > 32) G = 32;
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    constexpr int WPB = 4;
    const int64_t blocks = fsw_cdiv(warps, WPB);
    const size_t smem = (size_t)WPB * FswCoopSmem<R, L, SAVE_RANK, CLOUD>::PER_WARP * sizeof(float);
    auto kern = fsw_coop_fwd_kernel<R, L, HAS_COL, SAVE_RANK, CLOUD>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const std::string label = std::string(SAVE_RANK ? "fwdr_" : "fwd_") + (CLOUD ? "cloud_u" : "coop_u") + std::to_string(R * L) + "_f32";  // R x L slots
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, WPB * 32, smem, st>>>(a, lo, hi, (int)G, nchunks, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_coop_fwd_kernel");
    return FSW_OK;
}

template <int R, int L>
int launch_coop(const SegArgs<float>& a, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0, const float* bias,
                unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t,
                int tab_n0, int tab_ld4, cudaStream_t st) {
    const bool has_col = a.col != nullptr;
    if (a.projX != nullptr) {   // point-cloud mode: dense batches only, keys formed on the fly
        if (has_col) return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: point-cloud mode needs a dense batch");
        if constexpr (R >= 24 && L <= 8) return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: no point-cloud instance of %d x %d", R, L);
        else return ranks ? launch_coop_fwd<R, L, false, true, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st)
                     : launch_coop_fwd<R, L, false, false, true>(a, lo, hi, out, ld_out, out_col0, bias, nullptr, 0, nullptr, 0, gtab_c, gtab_t, tab_n0, tab_ld4, st);
    }
    {
        if (ranks) {
            return has_col ? launch_coop_fwd<R, L, true, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st)
                           : launch_coop_fwd<R, L, false, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
        }
        return has_col ? launch_coop_fwd<R, L, true, false>(a, lo, hi, out, ld_out, out_col0, bias, nullptr, 0, nullptr, 0, gtab_c, gtab_t, tab_n0, tab_ld4, st)
                       : launch_coop_fwd<R, L, false, false>(a, lo, hi, out, ld_out, out_col0, bias, nullptr, 0, nullptr, 0, gtab_c, gtab_t, tab_n0, tab_ld4, st);
    }
}

}  // namespace

int fsw_build_fwd_tables(const float* freqs, int K, int n_lo, int n_hi, int ld4, float* tab_c, float* tab_t, cudaStream_t st) {
    fsw_prof_begin("coef_tables", st);
    fsw_build_fwd_tables_kernel<<<dim3(n_hi - n_lo + 1, ld4), 256, 0, st>>>(freqs, K, n_lo, ld4, tab_c, tab_t);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_build_fwd_tables_kernel");
    return FSW_OK;
}

// uniform-weight fp32 segments order[lo, hi) with n <= np, np in {48, 64, 96, 128, 192, 256, 384, 512} (graphs and dense
// batches) or 1024 (dense batches); gtab_c / gtab_t: tables of fsw_build_fwd_tables covering n >= tab_n0 with tab_ld4 position blocks per n
int fsw_packed_forward_u(const SegArgs<float>& a, int np, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0,
                         const float* bias, unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c,
                         const float* gtab_t, int tab_n0, int tab_ld4, cudaStream_t st) {
    if (gtab_c == nullptr) return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: coefficient table missing");
#define FSW_COOP_CASE(NP_, R_, L_) \
    case NP_: return launch_coop<R_, L_>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
    // run-length / lanes-per-slice choice per class: fewer lanes per slice = fewer cross-lane (shuffle) merge steps per element
    static const bool wide = [] { const char* e = getenv("FSW_COOP_WIDE"); return e == nullptr || atoi(e) != 0; }();
    if (wide && a.projX == nullptr) {
        switch (np) {
            FSW_COOP_CASE(96, 24, 4)
            FSW_COOP_CASE(128, 32, 4)
            FSW_COOP_CASE(192, 24, 8)
            FSW_COOP_CASE(256, 32, 8)
        }
    }
    switch (np) {
        FSW_COOP_CASE(48, 12, 4)
        FSW_COOP_CASE(64, 16, 4)
        FSW_COOP_CASE(96, 12, 8)
        FSW_COOP_CASE(128, 16, 8)
        FSW_COOP_CASE(192, 12, 16)
        FSW_COOP_CASE(256, 16, 16)
        FSW_COOP_CASE(384, 24, 16)
        FSW_COOP_CASE(512, 32, 16)
        FSW_COOP_CASE(1024, 32, 32)
    }
#undef FSW_COOP_CASE
    return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: class %d", np);
}
