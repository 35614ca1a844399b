// K2 / K3, uniform-weight fast path for small segments (n <= 32; up to 64 when the packed-key kernels are not
// available, <= 32 fp64) and the rank-based backward used for every uniform-weight segment.
//
//   fsw_small_fwd_kernel<T, NP, HAS_COL, SAVE_RANK>
//     one THREAD per (segment, slice), lanes = 32 consecutive slices, keys in registers, merge-exchange
//     network of exactly NP in {4, 8, 12, 16, 24, 32, 48, 64} slots.  Per element the gather costs
//     SHFL + IMAD.WIDE + LDG + select; segment metadata (order -> rowptr -> col) is prefetched two
//     segments ahead so that only the key gather itself is an exposed memory round trip.  The Fourier
//     coefficients cos(pi xi (2j+1)/n) live in a per-warp shared-memory table that is rebuilt only when
//     n changes (segments arrive sorted by n).
//     SAVE_RANK: the network carries the element index; the sorted position of every ORIGINAL element
//     is un-permuted through shared memory and written as uint16 rank[(e0+i), k] - the analogue of the
//     compressed permutation the reference saves for its backward (fsw_embedding.py:2041-2050).
//   fsw_rank_bwd_kernel<T, HAS_COL, NEED_DXI>
//     backward without any sorting: dL/dp_i = g (1+xi) A0(n) cos(pi xi (2 r_i + 1)/n) with r_i read back;
//     streaming, one coalesced 128-byte row per element (red.add for graphs, plain store for dense
//     batches); d/dxi from the same table pass.
#include <stdlib.h>

#include "fsw_sortnet.cuh"

namespace {

struct SegMeta {
    int s;
    int n;
    int64_t e0;
};

template <typename T>
__device__ __forceinline__ int fsw_ld_order(const SegArgs<T>& a, int q, int last) {
    const int qq = (q < last) ? q : last - 1;
    return a.order ? __ldg(a.order + qq) : qq;
}

template <typename T>
__device__ __forceinline__ void fsw_ld_range(const SegArgs<T>& a, int s, int64_t& e0, int& n) {
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

// ---------------------------------------------------------------------------------------------------
template <typename T, int NP, bool HAS_COL, bool SAVE_RANK>
__global__ void __launch_bounds__(128) fsw_small_fwd_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks,
                                                            T* __restrict__ out, int64_t ld_out, int64_t out_col0,
                                                            const T* __restrict__ bias, unsigned short* __restrict__ ranks,
                                                            int64_t ldr, T* __restrict__ dxi_out, int64_t ld_dxi) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // per warp: coefficient table [NP][32] T; when SAVE_RANK (training) also the d/dxi table [NP][32] T and the
    // rank scratch [NP][32] int
    constexpr size_t kWarpBytes = (size_t)NP * 32 * (sizeof(T) + (SAVE_RANK ? sizeof(T) + sizeof(int) : 0));
    T* tab = reinterpret_cast<T*>(fsw_smem_raw + warp * kWarpBytes);
    T* tabt = tab + NP * 32;
    int* srank = reinterpret_cast<int*>(tabt + NP * 32);
    (void)srank;
    (void)tabt;

    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first64 = (int64_t)seg_lo + item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k = chunk * 32 + lane;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const T xi = __ldg(a.freqs + kk);
    const double xid = (double)xi;
    const T bk = (bias != nullptr) ? __ldg(bias + kk) : (T)0;
    const int ldb = (int)(a.ldp * (int64_t)sizeof(T));
    const char* xp_bytes = reinterpret_cast<const char*>(a.Xp + kk);
    const char* ep_bytes = a.Ep ? reinterpret_cast<const char*>(a.Ep + kk) : nullptr;

    // ---- software pipeline over segments: order two ahead, row range one ahead, column ids one ahead ----
    SegMeta cur, nx1;
    int s2;
    int c0, c1;
    cur.s = fsw_ld_order(a, first, last);
    fsw_ld_range(a, cur.s, cur.e0, cur.n);
    fsw_load_cols<NP, HAS_COL>(a.col, cur.e0, cur.n, lane, c0, c1);
    nx1.s = fsw_ld_order(a, first + 1, last);
    fsw_ld_range(a, nx1.s, nx1.e0, nx1.n);
    s2 = fsw_ld_order(a, first + 2, last);

    int n_prev = -1;
    T A = (T)0, A0 = (T)0, A0p = (T)0;
    (void)A0;
    (void)A0p;

    for (int q = first; q < last; ++q) {
        const int n = cur.n;
        T key[NP];
        fsw_gather_lean<T, NP, HAS_COL>(xp_bytes, ldb, ep_bytes, cur.e0, n, c0, c1, key);
        // prefetches for the following segments (their addresses were loaded one iteration ago)
        int c0n, c1n;
        fsw_load_cols<NP, HAS_COL>(a.col, nx1.e0, nx1.n, lane, c0n, c1n);
        SegMeta nx2;
        nx2.s = s2;
        fsw_ld_range(a, s2, nx2.e0, nx2.n);
        const int s3 = fsw_ld_order(a, q + 3, last);

        if (n != n_prev) {
            const double u = xid / (double)n;
            const T wn = (T)(1.0 / (double)n);
#pragma unroll 1
            for (int j = 0; j < NP; ++j) {
                const T rr = Num<T>::reduce(u * (double)(2 * j + 1));
                tab[j * 32 + lane] = (j < n) ? Num<T>::cospi_(rr) : (T)0;
                if (SAVE_RANK) tabt[j * 32 + lane] = (j < n) ? (T)M_PI * wn * (T)(2 * j + 1) * Num<T>::sinpi_(rr) : (T)0;
            }
            fsw_amplitude<T, SAVE_RANK>(u, wn, xi, A0, A0p);
            A = ((T)1 + xi) * A0;
            n_prev = n;
        }

        T acc = (T)0, acc2 = (T)0;
        if constexpr (SAVE_RANK && sizeof(T) == 4) {
            // fp32 training path: PACKED keys (see fsw_embed_packed.cu) - the network sorts one 32-bit word per element,
            // the order-preserving integer image of the key with its low 5 (6) bits replaced by the element index: two
            // integer min/max per comparator instead of a compare and four selects.  The full keys wait in shared
            // memory (the rank scratch); adjacent positions with equal truncated keys are bubble-ordered on them, so
            // the order is exact and ties keep element order.
            constexpr int IB = (NP > 32) ? 6 : 5, IM = (1 << IB) - 1;
            static_assert(NP <= (1 << IB), "element index must fit the packed low bits");
            float* fk = reinterpret_cast<float*>(srank) + lane;
            int p[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float v = (j < n) ? key[j] + 0.0f : __int_as_float(0x7f000000 | (j << IB));  // -0 -> +0; padding: own groups
                fk[j * 32] = v;
                const int b = __float_as_int(v);
                const int t = b ^ ((b >> 31) & 0x7fffffff);
                p[j] = (t & ~IM) | j;
            }
            fsw_sort_network<NP>([&](int i, int l) {
                const int x = p[i], y = p[l];
                p[i] = min(x, y);
                p[l] = max(x, y);
            });
            if constexpr (NP > 1) {
                unsigned mn = 0xffffffffu;
#pragma unroll
                for (int j = 0; j + 1 < NP; ++j) mn = min(mn, (unsigned)(p[j] ^ p[j + 1]));
                if (__any_sync(FSW_FULL, mn <= (unsigned)IM)) {
                    bool again = true;
                    while (again) {
                        bool swapped = false;
#pragma unroll
                        for (int j = 0; j + 1 < NP; ++j) {
                            const int x = p[j], y = p[j + 1];
                            if ((unsigned)(x ^ y) <= (unsigned)IM) {
                                if (fk[(x & IM) * 32] > fk[(y & IM) * 32]) {
                                    p[j] = y;
                                    p[j + 1] = x;
                                    swapped = true;
                                }
                            }
                        }
                        again = __any_sync(FSW_FULL, swapped);
                    }
                }
            }
            const bool want_dxi = dxi_out != nullptr;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int ix = (p[j] & IM) * 32;
                const float kv = fk[ix];
                acc = fmaf(kv, tab[j * 32 + lane], acc);               // table rows >= n are zero: padding keys cancel
                if (want_dxi) acc2 = fmaf(kv, tabt[j * 32 + lane], acc2);
                reinterpret_cast<int*>(fk)[ix] = j;                    // the consumed key's slot now holds the element's rank
            }
            if (act) out[fsw_rowoff(cur.s, ld_out) + out_col0 + k] = A * acc + bk;
            if (want_dxi && act) dxi_out[fsw_rowoff(cur.s, ld_dxi) + k] = A0 * acc + ((T)1 + xi) * (A0p * acc - A0 * acc2);
        } else {
            if constexpr (SAVE_RANK) {
                int idx[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) idx[j] = j;
                fsw_sort_network<NP>([&](int i, int l) {
                    const T x = key[i], y = key[l];
                    const int px = idx[i], py = idx[l];
                    const bool sw = x > y;
                    key[i] = sw ? y : x;
                    key[l] = sw ? x : y;
                    idx[i] = sw ? py : px;
                    idx[l] = sw ? px : py;
                });
#pragma unroll
                for (int j = 0; j < NP; ++j) srank[idx[j] * 32 + lane] = j;
            } else {
                fsw_sort_network<NP>([&](int i, int l) {
                    const T x = key[i], y = key[l];
                    key[i] = fmin(x, y);
                    key[l] = fmax(x, y);
                });
            }
#pragma unroll
            for (int j = 0; j < NP; ++j) acc = fma(key[j], tab[j * 32 + lane], acc);
            if (act) out[fsw_rowoff(cur.s, ld_out) + out_col0 + k] = A * acc + bk;
            if constexpr (SAVE_RANK) {
                if (dxi_out != nullptr) {  // d out / d xi of this (segment, slice), summed against g in the backward
#pragma unroll
                    for (int j = 0; j < NP; ++j) acc2 = fma(key[j], tabt[j * 32 + lane], acc2);
                    if (act) dxi_out[fsw_rowoff(cur.s, ld_dxi) + k] = A0 * acc + ((T)1 + xi) * (A0p * acc - A0 * acc2);
                }
            }
        }
        if constexpr (SAVE_RANK) {
            // one base address per segment, then a 32-bit byte offset per element: the 64 x 64-bit row arithmetic per store
            // that `rp[i * ldr]` compiled to was 31 % of all instructions of this kernel (profiles/r2/README.md)
            char* rb = reinterpret_cast<char*>(ranks + fsw_rowoff(cur.e0, ldr) + k);
            const unsigned ldrb = (unsigned)ldr * 2u;
            const int nst = act ? n : 0;
#pragma unroll
            for (int i = 0; i < NP; ++i)
                if (i < nst) *reinterpret_cast<unsigned short*>(rb + (size_t)(i * ldrb)) = (unsigned short)srank[i * 32 + lane];
        }
        // rotate the pipeline
        cur = nx1;
        c0 = c0n;
        c1 = c1n;
        nx1 = nx2;
        s2 = s3;
    }
}

// ---------------------------------------------------------------------------------------------------
// Rank-based backward.  CTA = W warps working on the same (segment, slice chunk): the warps split the
// elements of the segment; the coefficient tables are shared by the CTA.
// ---------------------------------------------------------------------------------------------------
template <typename T, int W, bool HAS_COL, bool NEED_DXI>
__global__ void __launch_bounds__(W * 32) fsw_rank_bwd_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks, int cap,
                                                              const unsigned short* __restrict__ ranks, int64_t ldr,
                                                              const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                              T* __restrict__ dXp, T* __restrict__ dEp,
                                                              double* __restrict__ dfreqs) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[W][32];
    T* tab_c = reinterpret_cast<T*>(fsw_smem_raw);
    T* tab_t = tab_c + (size_t)cap * 32;  // only when NEED_DXI
    (void)tab_t;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - item * nchunks;
    const int64_t first64 = (int64_t)seg_lo + (int64_t)item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k = chunk * 32 + lane;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const T xi = __ldg(a.freqs + kk);
    const double xid = (double)xi;
    const int ldb = (int)(a.ldp * (int64_t)sizeof(T));
    const char* xp_bytes = reinterpret_cast<const char*>(a.Xp + kk);

    int n_prev = -1;
    T A0 = (T)0, A0p = (T)0;
    double dxi_acc = 0.0;
    (void)A0p;

    for (int q = first; q < last; ++q) {
        const int s = fsw_ld_order(a, q, last);
        int64_t e0;
        int n;
        fsw_ld_range(a, s, e0, n);
        const T wn = (T)(1.0 / (double)n);
        const double u = xid / (double)n;
        if (n != n_prev) {
            __syncthreads();  // everyone is done with the previous tables
            fsw_amplitude<T, NEED_DXI>(u, wn, xi, A0, A0p);
            for (int r = warp; r < n; r += W) {
                const T rr = Num<T>::reduce(u * (double)(2 * r + 1));
                tab_c[r * 32 + lane] = Num<T>::cospi_(rr);
                if (NEED_DXI) tab_t[r * 32 + lane] = (T)M_PI * wn * (T)(2 * r + 1) * Num<T>::sinpi_(rr);
            }
            __syncthreads();
            n_prev = n;
        }
        const T gk = act ? g[fsw_rowoff(s, ld_g) + g_col0 + k] : (T)0;
        const T GA = gk * ((T)1 + xi) * A0;
        T Sc = (T)0, Ss = (T)0;
        const unsigned short* rp = ranks + fsw_rowoff(e0, ldr) + kk;
        for (int i = warp; i < n; i += W) {
            const int r = rp[fsw_rowoff(i, ldr)];
            const T c = tab_c[r * 32 + lane];
            const T v = GA * c;
            int64_t row = e0 + i;
            if (HAS_COL) row = __ldg(a.col + e0 + i);
            if (NEED_DXI) {
                const T p = __ldg(reinterpret_cast<const T*>(xp_bytes + row * ldb)) +
                            (a.Ep ? __ldg(a.Ep + fsw_rowoff(e0 + i, a.ldp) + kk) : (T)0);
                Sc = fma(p, c, Sc);
                Ss = fma(p, tab_t[r * 32 + lane], Ss);
            }
            if (act) {
                if (HAS_COL)
                    atomicAdd(dXp + fsw_rowoff(row, a.ldp) + k, v);
                else
                    dXp[fsw_rowoff(row, a.ldp) + k] = v;
                if (dEp) dEp[fsw_rowoff(e0 + i, a.ldp) + k] = v;
            }
        }
        if (NEED_DXI) dxi_acc += (double)gk * ((double)A0 * (double)Sc + (1.0 + xid) * ((double)A0p * (double)Sc - (double)A0 * (double)Ss));
    }
    if (NEED_DXI) {
        red[warp][lane] = dxi_acc;
        __syncthreads();
        if (warp == 0 && act) {
            double tot = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < W; ++w2) tot += red[w2][lane];
            atomicAdd(dfreqs + k, tot);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// fp32 rank-based backward, V consecutive slices per thread (a warp covers 32*V slices): the scatter is
// one red.global.add.v4.f32 (REDG.E.ADD.F32x4) per element and thread instead of four scalar atomics -
// the scalar version was bound by the L2 atomic rate (~1.8e11 lane-atomics/s measured, profiles/r1).
// Tables are laid out [rank][q][lane] so that every shared-memory access is conflict free.
// ---------------------------------------------------------------------------------------------------
template <int V>
struct FswVec;
template <>
struct FswVec<1> {
    using rank_t = unsigned short;
    using val_t = float;
};
template <>
struct FswVec<2> {
    using rank_t = unsigned int;
    using val_t = float2;
};
template <>
struct FswVec<4> {
    using rank_t = uint2;
    using val_t = float4;
};

template <int V>
__device__ __forceinline__ void fsw_unpack_ranks(const unsigned short* p, int (&r)[V]) {
    if constexpr (V == 1) {
        r[0] = *p;
    } else if constexpr (V == 2) {
        const unsigned int w = *reinterpret_cast<const unsigned int*>(p);
        r[0] = w & 0xffff;
        r[1] = w >> 16;
    } else {
        const uint2 w = *reinterpret_cast<const uint2*>(p);
        r[0] = w.x & 0xffff;
        r[1] = w.x >> 16;
        r[2] = w.y & 0xffff;
        r[3] = w.y >> 16;
    }
}

template <int V>
__device__ __forceinline__ void fsw_red_add(float* p, const float (&v)[V]) {
    if constexpr (V == 1) {
        atomicAdd(p, v[0]);
    } else if constexpr (V == 2) {
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
}

template <int V>
__device__ __forceinline__ void fsw_store_vec(float* p, const float (&v)[V]) {
    if constexpr (V == 1) {
        *p = v[0];
    } else if constexpr (V == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <int V>
__device__ __forceinline__ void fsw_load_vec(const float* p, float (&v)[V]) {
    if constexpr (V == 1) {
        v[0] = __ldg(p);
    } else if constexpr (V == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        v[0] = t.x;
        v[1] = t.y;
    } else {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x;
        v[1] = t.y;
        v[2] = t.z;
        v[3] = t.w;
    }
}

template <int V, int W, bool HAS_COL, bool NEED_DXI>
__global__ void __launch_bounds__(W * 32) fsw_rank_bwdv_kernel(SegArgs<float> a, int seg_lo, int seg_hi, int G, int nchunks, int cap,
                                                               const unsigned short* __restrict__ ranks, int64_t ldr,
                                                               const float* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                               float* __restrict__ dXp, float* __restrict__ dEp,
                                                               double* __restrict__ dfreqs) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[W][32];
    float* tab_c = reinterpret_cast<float*>(fsw_smem_raw);
    float* tab_t = tab_c + (size_t)cap * V * 32;  // only when NEED_DXI
    (void)tab_t;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x / nchunks;
    const int chunk = blockIdx.x - item * nchunks;
    const int64_t first64 = (int64_t)seg_lo + (int64_t)item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k0 = (chunk * 32 + lane) * V;
    const bool lane_in_row = k0 < a.ldp;  // rows are padded to a multiple of 8 >= K: a vector never straddles a row end
    float xi[V];
    double xid[V];
    bool act[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        act[q] = k0 + q < a.K;
        xi[q] = __ldg(a.freqs + (act[q] ? k0 + q : a.K - 1));
        xid[q] = (double)xi[q];
    }
    int n_prev = -1;
    float A0[V], A0p[V];
    double dxi_acc[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        A0[q] = 0.f;
        A0p[q] = 0.f;
        dxi_acc[q] = 0.0;
    }

    for (int qs = first; qs < last; ++qs) {
        const int s = fsw_ld_order(a, qs, last);
        int64_t e0;
        int n;
        fsw_ld_range(a, s, e0, n);
        const float wn = (float)(1.0 / (double)n);
        if (n != n_prev) {
            __syncthreads();  // everyone is done with the previous tables
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const double u = xid[q] / (double)n;
                fsw_amplitude<float, NEED_DXI>(u, wn, xi[q], A0[q], A0p[q]);
                for (int r = warp; r < n; r += W) {
                    const float rr = Num<float>::reduce(u * (double)(2 * r + 1));
                    tab_c[(r * V + q) * 32 + lane] = cospif(rr);
                    if (NEED_DXI) tab_t[(r * V + q) * 32 + lane] = (float)M_PI * wn * (float)(2 * r + 1) * sinpif(rr);
                }
            }
            __syncthreads();
            n_prev = n;
        }
        float gk[V], GA[V], Sc[V], Ss[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            gk[q] = act[q] ? g[fsw_rowoff(s, ld_g) + g_col0 + k0 + q] : 0.f;
            GA[q] = gk[q] * (1.f + xi[q]) * A0[q];
            Sc[q] = 0.f;
            Ss[q] = 0.f;
        }
        if (lane_in_row) {
            const unsigned short* rp = ranks + fsw_rowoff(e0, ldr) + k0;
#pragma unroll 2
            for (int i = warp; i < n; i += W) {
                int r[V];
                fsw_unpack_ranks<V>(rp + fsw_rowoff(i, ldr), r);
                int64_t row = e0 + i;
                if (HAS_COL) row = __ldg(a.col + e0 + i);
                float c[V], v[V];
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    if (!act[q]) r[q] = 0;  // padding columns of the rank matrix are never written
                    c[q] = tab_c[(r[q] * V + q) * 32 + lane];
                    v[q] = GA[q] * c[q];
                }
                if (NEED_DXI) {
                    float p[V];
                    fsw_load_vec<V>(a.Xp + fsw_rowoff(row, a.ldp) + k0, p);
                    if (a.Ep) {
                        float pe[V];
                        fsw_load_vec<V>(a.Ep + fsw_rowoff(e0 + i, a.ldp) + k0, pe);
#pragma unroll
                        for (int q = 0; q < V; ++q) p[q] += pe[q];
                    }
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        Sc[q] = fmaf(p[q], c[q], Sc[q]);
                        Ss[q] = fmaf(p[q], tab_t[(r[q] * V + q) * 32 + lane], Ss[q]);
                    }
                }
                if (HAS_COL)
                    fsw_red_add<V>(dXp + fsw_rowoff(row, a.ldp) + k0, v);
                else
                    fsw_store_vec<V>(dXp + fsw_rowoff(row, a.ldp) + k0, v);
                if (dEp) fsw_store_vec<V>(dEp + fsw_rowoff(e0 + i, a.ldp) + k0, v);
            }
        }
        if (NEED_DXI) {
#pragma unroll
            for (int q = 0; q < V; ++q)
                dxi_acc[q] += (double)gk[q] * ((double)A0[q] * (double)Sc[q] + (1.0 + xid[q]) * ((double)A0p[q] * (double)Sc[q] - (double)A0[q] * (double)Ss[q]));
        }
    }
    if (NEED_DXI) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
            __syncthreads();
            red[warp][lane] = dxi_acc[q];
            __syncthreads();
            if (warp == 0 && act[q]) {
                double tot = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < W; ++w2) tot += red[w2][lane];
                atomicAdd(dfreqs + k0 + q, tot);
            }
        }
    }
}

template <int V, int W, bool HAS_COL, bool NEED_DXI>
int launch_rank_bwdv(const SegArgs<float>& a, int lo, int hi, int cap, const unsigned short* ranks, int64_t ldr, const float* g,
                     int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs, cudaStream_t st) {
    const int nchunks = (a.K + 32 * V - 1) / (32 * V);
    int64_t G = (int64_t)(hi - lo) * nchunks / (148 * 16);
    if (G < 1) G = 1;
    if (G > 64) G = 64;
    const int64_t blocks = fsw_cdiv(hi - lo, G) * nchunks;
    const size_t smem = (size_t)cap * V * 32 * sizeof(float) * (NEED_DXI ? 2 : 1);
    auto kern = fsw_rank_bwdv_kernel<V, W, HAS_COL, NEED_DXI>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const std::string label = std::string("bwd_rank_u") + std::to_string(cap) + "_f32";
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, W * 32, smem, st>>>(a, lo, hi, (int)G, nchunks, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_rank_bwdv_kernel");
    return FSW_OK;
}

template <int V, int W>
int dispatch_rank_bwdv(const SegArgs<float>& a, int lo, int hi, int cap, const unsigned short* ranks, int64_t ldr, const float* g,
                       int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs, cudaStream_t st) {
    const bool has_col = a.col != nullptr;
    if (dfreqs)
        return has_col ? launch_rank_bwdv<V, W, true, true>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                       : launch_rank_bwdv<V, W, false, true>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
    return has_col ? launch_rank_bwdv<V, W, true, false>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                   : launch_rank_bwdv<V, W, false, false>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
}

// ---------------------------------------------------------------------------------------------------
// fp32 rank-based backward for n <= FSW_GTAB_NMAX with the coefficient tables in GLOBAL memory
// (a few MB, L1/L2 resident, rebuilt once per backward call by fsw_build_rank_tables_kernel):
//   tab[(n*(n-1)/2 + r) * ldp + k]   for n = 1..NMAX, r = 0..n-1
// No shared memory, no barriers: every WARP is independent (segment run x 128-slice chunk), so occupancy is
// set by registers only, and U elements are in flight per warp.  4 slices per thread -> REDG.E.ADD.F32x4.
// ---------------------------------------------------------------------------------------------------
constexpr int FSW_GTAB_NMAX = 128;
constexpr int64_t FSW_GTAB_ROWS = (int64_t)FSW_GTAB_NMAX * (FSW_GTAB_NMAX + 1) / 2;

__global__ void __launch_bounds__(256) fsw_build_rank_tables_kernel(const float* __restrict__ freqs, int K, int ldp,
                                                                    float* __restrict__ tab_c, float* __restrict__ tab_t,
                                                                    float* __restrict__ tab_A, float* __restrict__ tab_Ap,
                                                                    float2* __restrict__ tab_u) {
    // grid.x = n (1..NMAX), threads over (r, k)
    const int n = blockIdx.x + 1;
    const float wn = (float)(1.0 / (double)n);
    const int64_t off = (int64_t)n * (n - 1) / 2;
    for (int idx = threadIdx.x; (tab_c || tab_t) && idx < n * ldp; idx += blockDim.x) {
        const int r = idx / ldp, k = idx - r * ldp;
        float c = 0.f, t = 0.f;
        if (k < K) {
            const double u = (double)freqs[k] / (double)n;
            const float rr = Num<float>::reduce(u * (double)(2 * r + 1));
            c = cospif(rr);
            t = (float)M_PI * wn * (float)(2 * r + 1) * sinpif(rr);
        }
        if (tab_c) tab_c[(off + r) * ldp + k] = c;
        if (tab_t) tab_t[(off + r) * ldp + k] = t;
    }
    for (int k = threadIdx.x; k < ldp; k += blockDim.x) {
        float A0 = 0.f, A0p = 0.f;
        if (k < K) {
            const float xi = freqs[k];
            fsw_amplitude<float, true>((double)xi / (double)n, wn, xi, A0, A0p);
        }
        if (tab_A) tab_A[(int64_t)(n - 1) * ldp + k] = A0;
        if (tab_Ap) tab_Ap[(int64_t)(n - 1) * ldp + k] = A0p;
        if (tab_u) {  // xi / n as a double-float pair: lets the phase (2r+1) xi / n mod 2 be formed in fp32
            const double u = (k < K) ? (double)freqs[k] / (double)n : 0.0;
            const float hi = (float)u;
            tab_u[(int64_t)(n - 1) * ldp + k] = make_float2(hi, (float)(u - (double)hi));
        }
    }
}

}  // namespace

// tab_c / tab_t [(n (n-1)/2 + r) * ldp + k] for n = 1..nmax, r < n; tab_A / tab_Ap [(n-1) * ldp + k] (each may be NULL)
int fsw_build_coef_tables(const float* freqs, int K, int ldp, int nmax, float* tab_c, float* tab_t, float* tab_A, float* tab_Ap,
                          cudaStream_t st, float2* tab_u) {
    fsw_prof_begin("coef_tables", st);
    fsw_build_rank_tables_kernel<<<nmax, 256, 0, st>>>(freqs, K, ldp, tab_c, tab_t, tab_A, tab_Ap, tab_u);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_build_rank_tables_kernel");
    return FSW_OK;
}

namespace {

template <bool HAS_COL, bool NEED_DXI>
__global__ void __launch_bounds__(128, 5) fsw_rank_bwdg_kernel(SegArgs<float> a, int seg_lo, int seg_hi, int G, int nchunks,
                                                            const unsigned short* __restrict__ ranks, int64_t ldr,
                                                            const float* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                            float* __restrict__ dXp, float* __restrict__ dEp,
                                                            double* __restrict__ dfreqs, const float* __restrict__ tab_c,
                                                            const float* __restrict__ tab_t, const float* __restrict__ tab_A,
                                                            const float* __restrict__ tab_Ap) {
    constexpr int V = 4, U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first64 = (int64_t)seg_lo + item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k0 = (chunk * 32 + lane) * V;
    if (k0 >= a.ldp) return;  // rows are padded to a multiple of 8 >= K: a 4-vector never straddles a row end
    const int ldp = (int)a.ldp;
    float xi[V];
    bool act[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        act[q] = k0 + q < a.K;
        xi[q] = __ldg(a.freqs + (act[q] ? k0 + q : a.K - 1));
    }
    double dxi_acc[V] = {0.0, 0.0, 0.0, 0.0};

    for (int qs = first; qs < last; ++qs) {
        const int s = fsw_ld_order(a, qs, last);
        int64_t e0;
        int n;
        fsw_ld_range(a, s, e0, n);
        float GA[V], Sc[V], Ss[V];
        {
            float A0[V];
            fsw_load_vec<V>(tab_A + (int64_t)(n - 1) * ldp + k0, A0);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const float gk = act[q] ? __ldg(g + fsw_rowoff(s, ld_g) + g_col0 + k0 + q) : 0.f;
                GA[q] = gk * (1.f + xi[q]) * A0[q];
                Sc[q] = 0.f;
                Ss[q] = 0.f;
            }
        }
        const float* tc = tab_c + ((int64_t)n * (n - 1) / 2) * ldp + k0;
        const float* tt = NEED_DXI ? tab_t + ((int64_t)n * (n - 1) / 2) * ldp + k0 : nullptr;
        const unsigned short* rp = ranks + fsw_rowoff(e0, ldr) + k0;
        for (int i0 = 0; i0 < n; i0 += U) {
            int64_t row[U];
            int r[U][V];
            float p[U][V];
#pragma unroll
            for (int j = 0; j < U; ++j) {  // column ids: warp-uniform addresses (one broadcast transaction each)
                const int i = min(i0 + j, n - 1);  // clamped: the tail re-reads the last element, masked below
                row[j] = HAS_COL ? (int64_t)__ldg(a.col + e0 + i) : e0 + i;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int i = min(i0 + j, n - 1);
                fsw_unpack_ranks<V>(rp + fsw_rowoff(i, ldr), r[j]);
                if (NEED_DXI) fsw_load_vec<V>(a.Xp + fsw_rowoff(row[j], ldp) + k0, p[j]);
            }
            if (NEED_DXI && a.Ep) {
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    float pe[V];
                    fsw_load_vec<V>(a.Ep + fsw_rowoff(e0 + min(i0 + j, n - 1), ldp) + k0, pe);
#pragma unroll
                    for (int q = 0; q < V; ++q) p[j][q] += pe[q];
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                if (i0 + j < n) {
                    float c[V], v[V];
#pragma unroll
                    for (int q = 0; q < V; ++q) {
                        const int rq = act[q] ? r[j][q] : 0;  // padding columns of the rank matrix are never written
                        c[q] = __ldg(tc + (int64_t)rq * ldp + q);
                        v[q] = GA[q] * c[q];
                        if (NEED_DXI) {
                            Sc[q] = fmaf(p[j][q], c[q], Sc[q]);
                            Ss[q] = fmaf(p[j][q], __ldg(tt + (int64_t)rq * ldp + q), Ss[q]);
                        }
                    }
                    if (HAS_COL)
                        fsw_red_add<V>(dXp + fsw_rowoff(row[j], ldp) + k0, v);
                    else
                        fsw_store_vec<V>(dXp + fsw_rowoff(row[j], ldp) + k0, v);
                    if (dEp) fsw_store_vec<V>(dEp + fsw_rowoff(e0 + i0 + j, ldp) + k0, v);
                }
            }
        }
        if (NEED_DXI) {
            // amplitudes and g are re-read here (L1 hits) instead of being kept live across the element loop
            float A0[V], A0p[V];
            fsw_load_vec<V>(tab_A + (int64_t)(n - 1) * ldp + k0, A0);
            fsw_load_vec<V>(tab_Ap + (int64_t)(n - 1) * ldp + k0, A0p);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const float gk = act[q] ? __ldg(g + fsw_rowoff(s, ld_g) + g_col0 + k0 + q) : 0.f;
                dxi_acc[q] += (double)gk * ((double)A0[q] * (double)Sc[q] + (1.0 + (double)xi[q]) * ((double)A0p[q] * (double)Sc[q] - (double)A0[q] * (double)Ss[q]));
            }
        }
    }
    if (NEED_DXI) {
#pragma unroll
        for (int q = 0; q < V; ++q)
            if (act[q]) atomicAdd(dfreqs + k0 + q, dxi_acc[q]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Dense batches (no rowptr, no column ids, unit weights): every element row belongs to ONE segment of n elements and
// receives exactly one value per slice, so the rank backward is a pure streaming kernel - 2 B of rank in, 4 B of
// gradient out per (element, slice), coalesced, no tables: xi/n (double-float) and g (1+xi) A0(n) live in registers.
// Block = 64 column groups (4 slices each) x 4 rows; blockIdx.x = (segment, chunk of 64 rows), blockIdx.y = 256-slice chunk.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fsw_rank_bwd_dense_kernel(SegArgs<float> a, int n, int chunks_per_seg,
                                                                 const unsigned short* __restrict__ ranks, int64_t ldr,
                                                                 const float* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                                 float* __restrict__ dXp, float* __restrict__ dEp) {
    constexpr int V = 4;
    const int cg = blockIdx.y * 64 + (threadIdx.x & 63);
    const int rsub = threadIdx.x >> 6;
    const int k0 = cg * V;
    const int ldp = (int)a.ldp;
    if (k0 >= ldp) return;
    const int64_t s = blockIdx.x / chunks_per_seg;
    const int chunk = (int)(blockIdx.x - s * chunks_per_seg);
    float uh[V], ul[V], ga[V];
    const float wn = (float)(1.0 / (double)n);
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const int k = k0 + q;
        uh[q] = ul[q] = ga[q] = 0.f;
        if (k < a.K) {
            const float xi = __ldg(a.freqs + k);
            const double u = (double)xi / (double)n;
            uh[q] = (float)u;
            ul[q] = (float)(u - (double)uh[q]);
            float A0, A0p;
            fsw_amplitude<float, false>(u, wn, xi, A0, A0p);
            ga[q] = __ldg(g + fsw_rowoff(s, ld_g) + g_col0 + k) * (1.f + xi) * A0;
        }
    }
    const int64_t e0 = s * n;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
        const int r = chunk * 64 + it * 4 + rsub;
        if (r >= n) break;
        const int64_t e = e0 + r;
        int rk[V];
        fsw_unpack_ranks<V>(ranks + fsw_rowoff(e, ldr) + k0, rk);
        float v[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const float m = __uint_as_float(0x4B000000u | (unsigned)(2 * rk[q] + 1)) - 8388608.0f;  // exact below 2^23
            const float ph = m * uh[q];
            const float pe = fmaf(m, uh[q], -ph);
            const float pl = fmaf(m, ul[q], pe);
            const float hq = (0.5f * ph + 12582912.0f) - 12582912.0f;
            const float red = fmaf(hq, -2.0f, ph);
            v[q] = ga[q] * fsw_cospi_unit(red + pl);
        }
        fsw_store_vec<V>(dXp + fsw_rowoff(e, ldp) + k0, v);
        if (dEp) fsw_store_vec<V>(dEp + fsw_rowoff(e, ldp) + k0, v);
    }
}

// ---------------------------------------------------------------------------------------------------
// SOURCE-major rank backward (graphs): warp = (point row j, 128-slice chunk).  It walks the transposed
// structure (all (segment, slot) pairs that reference row j), accumulates GA[seg] * cos(...) in registers
// and writes dXp[j] ONCE with a plain 128-bit store: no atomics, no read-modify-write of dXp.
// Measured (profiles/micro/atomic_bw.cu): scattered red.add to a 1.9 GB matrix sustains 2.3 TB/s of
// payload (each update is a DRAM read + write) while gathers run at 5.7 TB/s.
//   GA[s, k] = g[s, k] (1 + xi_k) A0(n_s, k) is formed once per segment by fsw_scale_grad_kernel (aligned,
//   zero for segments that are not eligible), so that one 128-bit gather per pair replaces 4 misaligned
//   scalar loads of g plus an amplitude lookup.
// Pairs whose segment is not eligible (tn == 0: more than 128 elements or non-uniform weights) are
// skipped here and added afterwards by the destination-major kernels (atomics).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fsw_scale_grad_kernel(SegArgs<float> a, int64_t S, const float* __restrict__ g, int64_t ld_g,
                                                             int64_t g_col0, const float* __restrict__ tab_A, int nmax,
                                                             float* __restrict__ GA) {
    // one warp per segment row: n and the eligibility are read once, the row is streamed in coalesced pieces
    const int ldp = (int)a.ldp;
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= S) return;
    const int n = __ldg(a.rowptr + s + 1) - __ldg(a.rowptr + s);
    const bool ok = (__ldg(a.info + s) & FSW_INFO_UNIFORM) && n >= 1 && n <= nmax;
    const float* gr = g + fsw_rowoff(s, ld_g) + g_col0;
    float* out = GA + fsw_rowoff(s, ldp);
    if (!ok || n <= FSW_RANKT_TAB) {
        const float* ar = tab_A + fsw_rowoff(max(n, 1) - 1, ldp);
        for (int k = lane; k < ldp; k += 32) {
            float v = 0.f;
            if (ok && k < a.K) v = __ldg(gr + k) * (1.f + __ldg(a.freqs + k)) * __ldg(ar + k);
            out[k] = v;
        }
    } else {  // the few segments beyond the amplitude table
        for (int k = lane; k < ldp; k += 32) {
            float v = 0.f;
            if (k < a.K) {
                const float xi = __ldg(a.freqs + k);
                float A0, A0p;
                fsw_amplitude<float, false>((double)xi / (double)n, (float)(1.0 / (double)n), xi, A0, A0p);
                v = __ldg(gr + k) * (1.f + xi) * A0;
            }
            out[k] = v;
        }
    }
}

// One batch of U (segment, slot) pairs of a source row: packed ranks and pre-scaled gradients of V = 4 P slices per lane.
template <int U, int P>
struct FswPairBatch {
    uint2 rk[U][P];
    float4 ga[U][P];
    int nn[U];
    int slot[U];
};

// V slices per lane: 4 (rows of up to 128 slices per warp) or 8 (up to 256: one warp per row for K = 199)
template <int V, int U, int MINB, bool PF>
__global__ void __launch_bounds__(128, MINB) fsw_rank_bwdT_kernel(SegArgs<float> a, int64_t Nrows, int nchunks,
                                                            const int32_t* __restrict__ tptr, const int32_t* __restrict__ tseg,
                                                            const int32_t* __restrict__ tslot, const int32_t* __restrict__ tn,
                                                            const unsigned short* __restrict__ ranks, int64_t ldr,
                                                            const float* __restrict__ GA, float* __restrict__ dXp,
                                                            float* __restrict__ dEp, const float2* __restrict__ tab_u) {
    constexpr int P = V / 4;
    const int lane = threadIdx.x & 31;
    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t j = wglobal / nchunks;
    const int chunk = (int)(wglobal - j * nchunks);
    if (j >= Nrows) return;
    const int k0 = (chunk * 32 + lane) * V;
    if (k0 >= a.ldp) return;
    const int ldp = (int)a.ldp, ldri = (int)ldr;
    float2 acc2[V / 2], xi2[V / 2];
#pragma unroll
    for (int q = 0; q < V / 2; ++q) {
        acc2[q] = make_float2(0.f, 0.f);
        const int ka = k0 + 2 * q;
        xi2[q] = make_float2(ka < a.K ? __ldg(a.freqs + ka) : 0.f, ka + 1 < a.K ? __ldg(a.freqs + ka + 1) : 0.f);
    }
    const int t_beg = __ldg(tptr + j), t_end = __ldg(tptr + j + 1);
    const unsigned lanemask = __activemask();  // lanes beyond the padded row width have left
    const int nlanes = __popc(lanemask);       // active lanes are a prefix 0..nlanes-1
    const int step = nlanes >= U ? (nlanes & ~(U - 1)) : nlanes;
    const unsigned short* rbase = ranks + k0;
    const float* gbase = GA + k0;
    int my_seg = 0, my_slot = 0, my_n = 0, cnt = 0;
    const int kc0 = chunk * 32 * V;                                     // first slice of this warp's chunk
    const int pf_bytes_r = PF ? min(ldp - kc0, 32 * V) * 2 : 0;         // bytes of a rank row inside the chunk (gradient row: twice)

    // ranks and gradients of the batch starting at triple t0 (loads only: issued one batch ahead of their use)
    auto load_batch = [&](int t0, FswPairBatch<U, P>& B) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int src = min(t0 + u, cnt - 1);
            const int seg = __shfl_sync(lanemask, my_seg, src);
            B.slot[u] = __shfl_sync(lanemask, my_slot, src);
            const int n = __shfl_sync(lanemask, my_n, src);
            B.nn[u] = t0 + u < cnt ? n : 0;
            const uint2* rp = reinterpret_cast<const uint2*>(rbase + fsw_rowoff(B.slot[u], ldri));
            const float4* gp = reinterpret_cast<const float4*>(gbase + fsw_rowoff(seg, ldp));
#pragma unroll
            for (int h = 0; h < P; ++h) {
                B.rk[u][h] = __ldg(rp + h);
                B.ga[u][h] = __ldg(gp + h);
            }
        }
    };
    // coefficients cos(pi (2r+1) xi / n) evaluated directly: a table lookup would scatter the 32 lanes of a warp over
    // 32 cache lines (each lane has its own rank) and bind the kernel on L1 wavefronts.  All of the arithmetic runs on PAIRS
    // of slices with Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot for two slices) - the kernel
    // is bound by issue slots, not by DRAM:
    //   xi/n = uh + ul   double-float, from 1/n = ih + il (Newton residual), formed with negated constants so that no
    //                    per-slice sign flip is needed (nul = -ul);
    //   m = 2r+1         exact float through the mantissa trick, nm = -m by one FMA;
    //   ph = m uh, npl = -(m ul + (m uh - ph))   phase = ph + pl in units of pi, |pl| ~ 1e-7 ph;
    //   k = rint(ph) by the magic-number add, red = ph - k in [-1/2, 1/2], cos(pi ph) = (-1)^k cos(pi red): the parity
    //   of k is the low mantissa bit of ph + 1.5 * 2^23 and is XORed into the sign of the gradient factor;
    //   cos(pi x) on [-1/2, 1/2]: even Taylor polynomial up to x^12 (truncation < 1e-8), Horner in x^2.
    // Padding columns carry GA = 0 and xi = 0, so whatever their (never written) ranks hold contributes exactly 0.
    auto f2c = [](float c) { return make_float2(c, c); };
    auto consume = [&](const FswPairBatch<U, P>& B) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (B.nn[u] > 0) {
                const float nf = (float)B.nn[u];                   // exact: n <= 32768
                const float ih = __frcp_rn(nf);
                const float il = fmaf(-nf, ih, 1.0f) * ih;         // 1/n = ih + il
                const float2 ih2 = f2c(ih), nih2 = f2c(-ih), nil2 = f2c(-il);
                float v[V];
                (void)v;
#pragma unroll
                for (int pr = 0; pr < V / 2; ++pr) {
                    const float2 uh = __fmul2_rn(xi2[pr], ih2);
                    const float2 nue = __ffma2_rn(xi2[pr], nih2, uh);          // -(xi ih - uh), exact
                    const float2 nul = __ffma2_rn(xi2[pr], nil2, nue);         // -ul
                    const uint2 rk = B.rk[u][pr >> 1];
                    const unsigned w = (pr & 1) ? rk.y : rk.x;                 // two uint16 ranks
                    const float2 mb = make_float2(__uint_as_float(0x4B000001u | ((w << 1) & 0x1fffeu)),
                                                  __uint_as_float(0x4B000001u | ((w >> 15) & 0x1fffeu)));
                    const float2 m = __fadd2_rn(mb, f2c(-8388608.0f));         // 2r+1, exact below 2^23
                    const float2 nm = __ffma2_rn(mb, f2c(-1.0f), f2c(8388608.0f));
                    const float2 ph = __fmul2_rn(m, uh);
                    const float2 nqe = __ffma2_rn(nm, uh, ph);                 // ph - m uh, exact
                    const float2 npl = __ffma2_rn(m, nul, nqe);                // -(m ul + m uh - ph)
                    const float2 t = __fadd2_rn(ph, f2c(12582912.0f));
                    const float2 kk = __fadd2_rn(t, f2c(-12582912.0f));        // rint(ph)
                    const float2 red = __ffma2_rn(kk, f2c(-1.0f), ph);         // exact
                    const float2 x = __ffma2_rn(npl, f2c(-1.0f), red);
                    const float2 y2 = __fmul2_rn(x, x);
                    float2 c = __ffma2_rn(y2, f2c(1.929574e-3f), f2c(-2.580689e-2f));
                    c = __ffma2_rn(y2, c, f2c(2.353306e-1f));
                    c = __ffma2_rn(y2, c, f2c(-1.335263f));
                    c = __ffma2_rn(y2, c, f2c(4.058712f));
                    c = __ffma2_rn(y2, c, f2c(-4.934802f));
                    c = __ffma2_rn(y2, c, f2c(1.0f));
                    const float4 g4 = B.ga[u][pr >> 1];
                    const float g0 = (pr & 1) ? g4.z : g4.x, g1 = (pr & 1) ? g4.w : g4.y;
                    const float2 gs = make_float2(__uint_as_float(__float_as_uint(g0) ^ (__float_as_uint(t.x) << 31)),
                                                  __uint_as_float(__float_as_uint(g1) ^ (__float_as_uint(t.y) << 31)));
                    if (dEp) {
                        const float2 vv = __fmul2_rn(gs, c);
                        v[2 * pr] = vv.x;
                        v[2 * pr + 1] = vv.y;
                        acc2[pr] = __fadd2_rn(acc2[pr], vv);
                    } else {
                        acc2[pr] = __ffma2_rn(gs, c, acc2[pr]);
                    }
                }
                if (dEp) {
                    float* ep = dEp + fsw_rowoff(B.slot[u], ldp) + k0;
#pragma unroll
                    for (int h = 0; h < P; ++h) reinterpret_cast<float4*>(ep)[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
                }
            }
        }
    };

    for (int tb = t_beg; tb < t_end; tb += step) {
        // one coalesced load of up to `step` (segment, slot, n) triples, broadcast by shuffles in load_batch
        cnt = min(step, t_end - tb);
        if (lane < cnt) {
            my_seg = __ldg(tseg + tb + lane);
            my_slot = __ldg(tslot + tb + lane);
            my_n = __ldg(tn + tb + lane);
            // lane l pulls the rank row and the gradient row of pair l towards L2 now: the demand loads of load_batch, two
            // pairs at a time, then find them there instead of paying a DRAM round trip each (more loads in flight
            // without spending registers on them)
            if (pf_bytes_r > 0) {
                const char* rrow = reinterpret_cast<const char*>(ranks + fsw_rowoff(my_slot, ldri) + kc0);
                const char* grow = reinterpret_cast<const char*>(GA + fsw_rowoff(my_seg, ldp) + kc0);
                for (int off = 0; off < pf_bytes_r; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + pf_bytes_r - 1));
                for (int off = 0; off < 2 * pf_bytes_r; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + 2 * pf_bytes_r - 1));
            }
        }
        FswPairBatch<U, P> A, B;
        load_batch(0, A);
        for (int t0 = 0; t0 < cnt; t0 += 2 * U) {
            const bool more = t0 + U < cnt;
            if (more) load_batch(t0 + U, B);
            consume(A);
            if (more) {
                if (t0 + 2 * U < cnt) load_batch(t0 + 2 * U, A);
                consume(B);
            }
        }
    }
    float* op = dXp + fsw_rowoff(j, ldp) + k0;
#pragma unroll
    for (int h = 0; h < P; ++h)
        reinterpret_cast<float4*>(op)[h] = make_float4(acc2[2 * h].x, acc2[2 * h].y, acc2[2 * h + 1].x, acc2[2 * h + 1].y);
}

// ---------------------------------------------------------------------------------------------------
// Source-major rank backward, streaming version: the (segment, slot) pairs of a run of consecutive source rows are ONE stream
// per warp, and the rank row and the pre-scaled gradient row of every pair travel global -> shared memory as asynchronous
// copies (LDGSTS, 16 bytes per lane) into a ring of D pairs per warp.  Nothing is staged in registers, so D - 1 pairs
// (~1.2 KB each at K = 199) are in flight per warp whatever the register budget - the register-staged kernel above holds 2 to 4
// and is bound by the DRAM round trip (profiles/r2/README.md).  Every lane reads back only the bytes it copied itself
// (its own 8 slices), so no barrier is needed; row boundaries only matter to the consumer, the copies run ahead across them.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fsw_ldgsts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fsw_ldgsts_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int V, int D>
struct FswRankStream {
    static constexpr int RB = V * 2;                          // rank bytes per lane and pair
    static constexpr int SLOT = 32 * RB + 32 * V * 4;         // ranks [32][RB] + gradient planes [V / 4][32][16]
    static constexpr int PER_WARP = D * SLOT;
};

// predicated asynchronous copy (no branch around it) to a 32-bit shared-memory address: BYTES = 8 or 16
template <int BYTES>
__device__ __forceinline__ void fsw_ldgsts_if(unsigned on, unsigned sdst, const void* gsrc) {
    if constexpr (BYTES == 16)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.cg.shared.global [%0], [%1], 16; }" ::"r"(sdst), "l"(gsrc), "r"(on) : "memory");
    else
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"(sdst), "l"(gsrc), "r"(on) : "memory");
}
__device__ __forceinline__ uint4 fsw_lds128(unsigned saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 fsw_lds64(unsigned saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}

template <int V, int D, int MINB, bool HAS_EP>
__global__ void __launch_bounds__(128, MINB) fsw_rank_bwdS_kernel(SegArgs<float> a, int64_t Nrows, int nchunks, int rpw,
                                                            const int32_t* __restrict__ tptr, const int32_t* __restrict__ tseg,
                                                            const int32_t* __restrict__ tslot, const int32_t* __restrict__ tn,
                                                            const unsigned short* __restrict__ ranks, int64_t ldr,
                                                            const float* __restrict__ GA, float* __restrict__ dXp,
                                                            float* __restrict__ dEp) {
    using RS = FswRankStream<V, D>;
    constexpr int P = V / 4;
    static_assert(V == 4 || V == 8, "4 or 8 slices per lane");
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    const int lane = threadIdx.x & 31;
    // this lane's bytes of a ring slot: ranks at lane * RB, gradient plane h at 32 RB + 512 h + lane * 16
    // (32-bit shared-window addresses, formed once: generic pointers cost an address conversion per access)
    const unsigned ring_r = (unsigned)__cvta_generic_to_shared(fsw_smem_raw) + (threadIdx.x >> 5) * RS::PER_WARP + lane * RS::RB;
    const unsigned ring_g = (unsigned)__cvta_generic_to_shared(fsw_smem_raw) + (threadIdx.x >> 5) * RS::PER_WARP + 32 * RS::RB + lane * 16;
    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t j0 = item * rpw;
    if (j0 >= Nrows) return;
    const int nrows = (int)((j0 + rpw < Nrows) ? rpw : Nrows - j0);   // <= 32
    const int k0 = (chunk * 32 + lane) * V;
    const int ldp = (int)a.ldp, ldri = (int)ldr;
    const bool lane_on = k0 < ldp;
    const unsigned on = lane_on ? 1u : 0u;
    float2 acc2[V / 2], xi2[V / 2];
#pragma unroll
    for (int q = 0; q < V / 2; ++q) {
        acc2[q] = make_float2(0.f, 0.f);
        const int ka = k0 + 2 * q;
        xi2[q] = make_float2(ka < a.K ? __ldg(a.freqs + ka) : 0.f, ka + 1 < a.K ? __ldg(a.freqs + ka + 1) : 0.f);
    }
    // lane l: end of the pair list of row j0 + l
    const int my_end = __ldg(tptr + j0 + min(lane + 1, nrows));
    const int T0 = __ldg(tptr + j0);
    const int T1 = __shfl_sync(FSW_FULL, my_end, nrows - 1);
    const unsigned short* rbase = ranks + (lane_on ? k0 : 0);
    const float* gbase = GA + (lane_on ? k0 : 0);

    // blocks of 32 (segment, slot, n) triples, one per lane: `cur` serves the copies being issued, `prv` the pairs still
    // being consumed behind a block boundary, `nxt` is in flight
    const int tlast = max(T1 - 1, 0);
    int bs = T0;
    int cur_seg = __ldg(tseg + min(bs + lane, tlast)), cur_slot = __ldg(tslot + min(bs + lane, tlast)), cur_n = __ldg(tn + min(bs + lane, tlast));
    int nxt_seg = __ldg(tseg + min(bs + 32 + lane, tlast)), nxt_slot = __ldg(tslot + min(bs + 32 + lane, tlast)), nxt_n = __ldg(tn + min(bs + 32 + lane, tlast));
    int prv_slot = 0, prv_n = 0;
    (void)prv_slot;

    int issued = T0;
    unsigned ioff = 0, coff = 0;   // byte offsets of the ring slots of the next copy / the next pair to consume
    const char* rbase_b = reinterpret_cast<const char*>(rbase);
    const char* gbase_b = reinterpret_cast<const char*>(gbase);
    const int ldr_b = ldri * 2, ldp_b = ldp * 4;   // row pitches in bytes (< 2^31 / rows by the int32 CSR contract)
    auto rotate = [&]() {          // the copies move on to the next block of 32 triples
        if (HAS_EP) prv_slot = cur_slot;
        prv_n = cur_n;
        cur_seg = nxt_seg;
        cur_slot = nxt_slot;
        cur_n = nxt_n;
        bs += 32;
        nxt_seg = __ldg(tseg + min(bs + 32 + lane, tlast));
        nxt_slot = __ldg(tslot + min(bs + 32 + lane, tlast));
        nxt_n = __ldg(tn + min(bs + 32 + lane, tlast));
    };
    auto issue = [&]() {           // needs issued < T1 and issued < bs + 32
        const int src = issued - bs;
        const int seg = __shfl_sync(FSW_FULL, cur_seg, src);
        const int slot = __shfl_sync(FSW_FULL, cur_slot, src);
        const char* rp;            // base + row * pitch as one IMAD.WIDE each
        const char* gp;
        asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(rp) : "r"(slot), "r"(ldr_b), "l"(rbase_b));
        asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(gp) : "r"(seg), "r"(ldp_b), "l"(gbase_b));
        fsw_ldgsts_if<RS::RB>(on, ring_r + ioff, rp);
#pragma unroll
        for (int h = 0; h < P; ++h) fsw_ldgsts_if<16>(on, ring_g + ioff + h * 512, gp + 16 * h);
        ++issued;
        ioff = (ioff + RS::SLOT == D * RS::SLOT) ? 0u : ioff + RS::SLOT;
    };
    auto f2c = [](float c) { return make_float2(c, c); };

    int jr = 0;
    int row_end = __shfl_sync(FSW_FULL, my_end, 0);
    auto flush_row = [&]() {
        if (lane_on) {
            float* op = dXp + fsw_rowoff(j0 + jr, ldp) + k0;
#pragma unroll
            for (int h = 0; h < P; ++h)
                reinterpret_cast<float4*>(op)[h] = make_float4(acc2[2 * h].x, acc2[2 * h].y, acc2[2 * h + 1].x, acc2[2 * h + 1].y);
        }
#pragma unroll
        for (int q = 0; q < V / 2; ++q) acc2[q] = make_float2(0.f, 0.f);
        ++jr;
        row_end = __shfl_sync(FSW_FULL, my_end, min(jr, 31));
    };
    // one pair: its copies have landed.  The arithmetic is that of fsw_rank_bwdT_kernel (see there): cos(pi (2r+1) xi / n) on
    // pairs of slices in packed fp32.  A pair of a segment that is not served here (tn = 0: general weights, more than nmax
    // elements) carries a zero gradient row (fsw_scale_grad_kernel), so it runs through with n = 1 and adds exactly 0.
    auto consume = [&](int tc) {
        while (tc >= row_end) flush_row();   // warp-uniform; also steps over rows without pairs
        const int src = tc - bs;             // >= -32: the consumer trails the copies by less than D <= 32 pairs
        const int n_c = __shfl_sync(FSW_FULL, cur_n, src & 31), n_p = __shfl_sync(FSW_FULL, prv_n, src & 31);
        const int n_raw = src >= 0 ? n_c : n_p;
        const int n = max(n_raw, 1);
        unsigned w[V / 2];
        float ga[V];
        if constexpr (V == 8) {
            const uint4 t = fsw_lds128(ring_r + coff);
            w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
        } else {
            const uint2 t = fsw_lds64(ring_r + coff);
            w[0] = t.x; w[1] = t.y;
        }
#pragma unroll
        for (int h = 0; h < P; ++h) {
            const uint4 t = fsw_lds128(ring_g + coff + h * 512);
            ga[4 * h] = __uint_as_float(t.x); ga[4 * h + 1] = __uint_as_float(t.y); ga[4 * h + 2] = __uint_as_float(t.z); ga[4 * h + 3] = __uint_as_float(t.w);
        }
        coff = (coff + RS::SLOT == D * RS::SLOT) ? 0u : coff + RS::SLOT;
        const float nf = (float)n;                         // exact: n <= 32768
        float ih;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ih) : "f"(nf));   // any ih within a few ulp of 1/n: il carries the rest
        const float il = fmaf(-nf, ih, 1.0f) * ih;         // 1/n = ih + il
        const float2 ih2 = f2c(ih), nih2 = f2c(-ih), nil2 = f2c(-il);
        float v[V];
        (void)v;
#pragma unroll
        for (int pr = 0; pr < V / 2; ++pr) {
            const float2 uh = __fmul2_rn(xi2[pr], ih2);
            const float2 nue = __ffma2_rn(xi2[pr], nih2, uh);          // -(xi ih - uh), exact
            const float2 nul = __ffma2_rn(xi2[pr], nil2, nue);         // -ul
            // the two uint16 ranks of w as floats 2^23 + r (one byte permute each), then 2r+1 and -(2r+1) by one FMA each (exact)
            const float2 mb = make_float2(__uint_as_float(__byte_perm(w[pr], 0x4B00u, 0x5410)), __uint_as_float(__byte_perm(w[pr], 0x4B00u, 0x5432)));
            const float2 m = __ffma2_rn(mb, f2c(2.0f), f2c(-16777215.0f));
            const float2 nm = __ffma2_rn(mb, f2c(-2.0f), f2c(16777215.0f));
            const float2 ph = __fmul2_rn(m, uh);
            const float2 nqe = __ffma2_rn(nm, uh, ph);                 // ph - m uh, exact
            const float2 npl = __ffma2_rn(m, nul, nqe);                // -(m ul + m uh - ph)
            const float2 t = __fadd2_rn(ph, f2c(12582912.0f));
            const float2 kk = __fadd2_rn(t, f2c(-12582912.0f));        // rint(ph)
            const float2 red = __ffma2_rn(kk, f2c(-1.0f), ph);         // exact
            const float2 x = __ffma2_rn(npl, f2c(-1.0f), red);
            const float2 y2 = __fmul2_rn(x, x);
            float2 c = __ffma2_rn(y2, f2c(1.929574e-3f), f2c(-2.580689e-2f));
            c = __ffma2_rn(y2, c, f2c(2.353306e-1f));
            c = __ffma2_rn(y2, c, f2c(-1.335263f));
            c = __ffma2_rn(y2, c, f2c(4.058712f));
            c = __ffma2_rn(y2, c, f2c(-4.934802f));
            c = __ffma2_rn(y2, c, f2c(1.0f));
            const float2 gs = make_float2(__uint_as_float(__float_as_uint(ga[2 * pr]) ^ (__float_as_uint(t.x) << 31)),
                                          __uint_as_float(__float_as_uint(ga[2 * pr + 1]) ^ (__float_as_uint(t.y) << 31)));
            if constexpr (HAS_EP) {
                const float2 vv = __fmul2_rn(gs, c);
                v[2 * pr] = vv.x;
                v[2 * pr + 1] = vv.y;
                acc2[pr] = __fadd2_rn(acc2[pr], vv);
            } else {
                acc2[pr] = __ffma2_rn(gs, c, acc2[pr]);
            }
        }
        if constexpr (HAS_EP) {
            const int s_c = __shfl_sync(FSW_FULL, cur_slot, src & 31), s_p = __shfl_sync(FSW_FULL, prv_slot, src & 31);
            if (lane_on && n_raw > 0) {
                float* ep = dEp + fsw_rowoff(src >= 0 ? s_c : s_p, ldp) + k0;
#pragma unroll
                for (int h = 0; h < P; ++h) reinterpret_cast<float4*>(ep)[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
            }
        }
    };

    // prologue: D - 1 groups (the copies of the first D - 1 pairs, or empty groups behind a short list)
#pragma unroll 1
    for (int i = 0; i < D - 1; ++i) {
        if (issued < T1) {
            if (issued - bs == 32) rotate();
            issue();
        }
        fsw_ldgsts_commit();
    }
    // steady state: issue pair tc + D - 1, wait for all but the newest D - 1 groups (the copies of pair tc have landed);
    // runs of pairs up to the next triple-block boundary, so that the per-pair path carries no block bookkeeping
    int tc = T0;
    const int Tsteady = T1 - (D - 1);
#pragma unroll 1
    while (tc < Tsteady) {
        if (issued - bs == 32) rotate();
        const int run_end = min(Tsteady, tc + (bs + 32 - issued));
#pragma unroll 1
        for (; tc < run_end; ++tc) {
            issue();
            fsw_ldgsts_commit();
            fsw_ldgsts_wait<D - 1>();
            consume(tc);
        }
    }
    // tail: everything has been issued
    fsw_ldgsts_wait<0>();
#pragma unroll 1
    for (; tc < T1; ++tc) consume(tc);
    while (jr < nrows) flush_row();   // the last row with pairs and the rows without any
}

template <bool HAS_COL, bool NEED_DXI>
int launch_rank_bwdg(const SegArgs<float>& a, int lo, int hi, const unsigned short* ranks, int64_t ldr, const float* g, int64_t ld_g,
                     int64_t g_col0, float* dXp, float* dEp, double* dfreqs, float* tables, cudaStream_t st) {
    const int ldp = (int)a.ldp;
    float* tab_c = tables;
    float* tab_t = tab_c + FSW_GTAB_ROWS * ldp;
    float* tab_A = tab_t + FSW_GTAB_ROWS * ldp;
    float* tab_Ap = tab_A + (int64_t)FSW_GTAB_NMAX * ldp;
    int rc0 = fsw_build_coef_tables(a.freqs, a.K, ldp, FSW_GTAB_NMAX, tab_c, NEED_DXI ? tab_t : nullptr, tab_A, tab_Ap, st);
    if (rc0) return rc0;
    const int nchunks = (a.K + 127) / 128;
    int64_t G = (int64_t)(hi - lo) * nchunks / (148 * 64);
    if (G < 1) G = 1;
    if (G > 64) G = 64;
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    fsw_prof_begin("bwd_rank_u128_f32", st);
    fsw_rank_bwdg_kernel<HAS_COL, NEED_DXI><<<(unsigned)blocks, 128, 0, st>>>(a, lo, hi, (int)G, nchunks, ranks, ldr, g, ld_g, g_col0, dXp, dEp,
                                                                               dfreqs, tab_c, tab_t, tab_A, tab_Ap);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_rank_bwdg_kernel");
    return FSW_OK;
}

int pick_G(int64_t cnt, int nchunks, int64_t target_warps, int maxG) {
    int64_t g = cnt * nchunks / target_warps;
    if (g < 1) g = 1;
    if (g > maxG) g = maxG;
    return (int)g;
}

template <typename T, int NP, bool HAS_COL, bool SAVE_RANK>
int launch_small_fwd(const SegArgs<T>& a, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                     unsigned short* ranks, int64_t ldr, T* dxi_out, int64_t ld_dxi, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    // up to 128 segments per warp: the coefficient table is rebuilt at the start of a warp's run and when n changes (with 32
    // segments per warp the rebuilds were 7 % of the instructions of the 13..16 class)
    const int G = pick_G(hi - lo, nchunks, 148 * 32, 128);
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    const size_t smem = (size_t)4 * NP * 32 * (sizeof(T) + (SAVE_RANK ? sizeof(T) + sizeof(int) : 0));
    auto kern = fsw_small_fwd_kernel<T, NP, HAS_COL, SAVE_RANK>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const std::string label = std::string(SAVE_RANK ? "fwdr_small_u" : "fwd_small_u") + std::to_string(NP) + (sizeof(T) == 4 ? "_f32" : "_f64");
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, 128, smem, st>>>(a, lo, hi, G, nchunks, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_small_fwd_kernel");
    return FSW_OK;
}

template <typename T, int NP>
int launch_small_fwd_np(const SegArgs<T>& a, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                        unsigned short* ranks, int64_t ldr, T* dxi_out, int64_t ld_dxi, cudaStream_t st) {
    const bool has_col = a.col != nullptr;
    if (ranks) {
        return has_col ? launch_small_fwd<T, NP, true, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st)
                       : launch_small_fwd<T, NP, false, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
    }
    return has_col ? launch_small_fwd<T, NP, true, false>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, nullptr, 0, st)
                   : launch_small_fwd<T, NP, false, false>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, nullptr, 0, st);
}

template <typename T, bool HAS_COL, bool NEED_DXI>
int launch_rank_bwd(const SegArgs<T>& a, int lo, int hi, int cap, const unsigned short* ranks, int64_t ldr, const T* g,
                    int64_t ld_g, int64_t g_col0, T* dXp, T* dEp, double* dfreqs, cudaStream_t st) {
    constexpr int W = 4;
    const int nchunks = (a.K + 31) / 32;
    const int G = pick_G(hi - lo, nchunks, 148 * 16, 64);
    const int64_t blocks = fsw_cdiv(hi - lo, G) * nchunks;
    const size_t smem = (size_t)cap * 32 * sizeof(T) * (NEED_DXI ? 2 : 1);
    auto kern = fsw_rank_bwd_kernel<T, W, HAS_COL, NEED_DXI>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const std::string label = std::string("bwd_rank_u") + std::to_string(cap) + (sizeof(T) == 4 ? "_f32" : "_f64");
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, W * 32, smem, st>>>(a, lo, hi, G, nchunks, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_rank_bwd_kernel");
    return FSW_OK;
}

}  // namespace

// uniform-weight segments order[lo, hi) with n <= np, np in {4, 8, 12, 16, 24, 32, 48, 64}
template <typename T>
int fsw_small_forward_u(const SegArgs<T>& a, int np, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                        unsigned short* ranks, int64_t ldr, T* dxi_out, int64_t ld_dxi, cudaStream_t st) {
    switch (np) {
        case 4: return launch_small_fwd_np<T, 4>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 8: return launch_small_fwd_np<T, 8>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 12: return launch_small_fwd_np<T, 12>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 16: return launch_small_fwd_np<T, 16>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 24: return launch_small_fwd_np<T, 24>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 32: return launch_small_fwd_np<T, 32>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
        case 48:
            if constexpr (sizeof(T) == 4) return launch_small_fwd_np<T, 48>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
            break;
        case 64:
            if constexpr (sizeof(T) == 4) return launch_small_fwd_np<T, 64>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
            break;
    }
    return fsw_fail(FSW_ERR_INVALID, "fsw_small_forward_u: class %d", np);
}

// uniform-weight segments order[lo, hi) with n <= cap (cap <= 512: tables in shared memory)
template <typename T>
int fsw_rank_backward_u(const SegArgs<T>& a, int lo, int hi, int cap, const unsigned short* ranks, int64_t ldr, const T* g,
                        int64_t ld_g, int64_t g_col0, T* dXp, T* dEp, double* dfreqs, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        // tables must fit shared memory: cap * V * 128 B (x2 with d/dxi)
        if (cap <= 128) return dispatch_rank_bwdv<4, 4>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        if (cap <= 256) return dispatch_rank_bwdv<2, 8>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        return dispatch_rank_bwdv<1, 8>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
    }
    const bool has_col = a.col != nullptr;
    if (dfreqs) {
        return has_col ? launch_rank_bwd<T, true, true>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                       : launch_rank_bwd<T, false, true>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
    }
    return has_col ? launch_rank_bwd<T, true, false>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                   : launch_rank_bwd<T, false, false>(a, lo, hi, cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
}

template int fsw_small_forward_u<float>(const SegArgs<float>&, int, int, int, float*, int64_t, int64_t, const float*, unsigned short*, int64_t, float*, int64_t, cudaStream_t);
template int fsw_small_forward_u<double>(const SegArgs<double>&, int, int, int, double*, int64_t, int64_t, const double*, unsigned short*, int64_t, double*, int64_t, cudaStream_t);
template int fsw_rank_backward_u<float>(const SegArgs<float>&, int, int, int, const unsigned short*, int64_t, const float*, int64_t, int64_t, float*, float*, double*, cudaStream_t);
template int fsw_rank_backward_u<double>(const SegArgs<double>&, int, int, int, const unsigned short*, int64_t, const double*, int64_t, int64_t, double*, double*, double*, cudaStream_t);

// fp32, uniform-weight segments order[lo, hi) with n <= 128: rank-based backward with global coefficient tables
size_t fsw_rank_tables_bytes(int64_t ldp) { return (size_t)((2 * FSW_GTAB_ROWS + 4 * FSW_GTAB_NMAX) * ldp) * sizeof(float); }

int fsw_rank_backward_g128(const SegArgs<float>& a, int lo, int hi, const unsigned short* ranks, int64_t ldr, const float* g,
                           int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs, void* tables, cudaStream_t st) {
    const bool has_col = a.col != nullptr;
    float* tb = (float*)tables;
    if (dfreqs)
        return has_col ? launch_rank_bwdg<true, true>(a, lo, hi, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, tb, st)
                       : launch_rank_bwdg<false, true>(a, lo, hi, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, tb, st);
    return has_col ? launch_rank_bwdg<true, false>(a, lo, hi, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, tb, st)
                   : launch_rank_bwdg<false, false>(a, lo, hi, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs, tb, st);
}

// fp32 dense batches with unit weights: S segments of n elements each, ranks recorded by the forward for all of them
int fsw_rank_backward_dense(const SegArgs<float>& a, int64_t S, int n, const unsigned short* ranks, int64_t ldr, const float* g,
                            int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, cudaStream_t st) {
    const int chunks = (n + 63) / 64;
    if (S * chunks > 0x7fffffff) return fsw_fail(FSW_ERR_INVALID, "fsw_rank_backward_dense: batch too large for one launch");
    dim3 grid((unsigned)(S * chunks), (unsigned)fsw_cdiv(a.ldp / 4, 64));
    fsw_prof_begin("bwd_rank_dense_f32", st);
    fsw_rank_bwd_dense_kernel<<<grid, 256, 0, st>>>(a, n, chunks, ranks, ldr, g, ld_g, g_col0, dXp, dEp);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_rank_bwd_dense_kernel");
    return FSW_OK;
}

// fp32 graphs: source-major rank backward over the transposed structure (eligible: uniform weights, n <= nmax =
// FSW_RANKT_ELIGIBLE(max n_eff)).
// Writes EVERY row of dXp (plain stores); must run before the kernels that add with atomics.
// `ga_buf` [S, ldp] floats is scratch for the pre-scaled upstream gradient.
int fsw_rank_backward_T(const SegArgs<float>& a, int64_t S, int64_t Nrows, int nmax, const int32_t* tptr, const int32_t* tseg,
                        const int32_t* tslot, const int32_t* tn, const unsigned short* ranks, int64_t ldr, const float* g,
                        int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, void* tables, float* ga_buf, cudaStream_t st) {
    const int ldp = (int)a.ldp;
    // only the amplitudes and xi/n are needed here, tabulated for n <= FSW_RANKT_TAB: 4 * 512 * ldp floats, well
    // inside fsw_rank_tables_bytes; larger segments (up to `nmax`) compute them on the fly
    static_assert(4 * FSW_RANKT_TAB <= 2 * FSW_GTAB_ROWS, "source-major tables must fit the rank-table scratch");
    float* tab_A = (float*)tables;
    float* tab_Ap = tab_A + (int64_t)FSW_RANKT_TAB * ldp;
    float2* tab_u = reinterpret_cast<float2*>(tab_Ap + (int64_t)FSW_RANKT_TAB * ldp);
    int rc0 = fsw_build_coef_tables(a.freqs, a.K, ldp, FSW_RANKT_TAB, nullptr, nullptr, tab_A, tab_Ap, st, tab_u);
    if (rc0) return rc0;
    fsw_prof_begin("bwd_scale_grad", st);
    fsw_scale_grad_kernel<<<(unsigned)fsw_cdiv(S, 8), 256, 0, st>>>(a, S, g, ld_g, g_col0, tab_A, nmax, ga_buf);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_scale_grad_kernel");
    // more than 128 slices: 8 per lane, so that a row of up to 256 slices is one warp (K = 199: 25 lanes of one warp
    // instead of 32 + 18 lanes of two, and half the per-row / per-pair bookkeeping)
    const bool wide = a.K > 128;
    const int nchunks = wide ? (a.K + 255) / 256 : (a.K + 127) / 128;
    const int64_t warps = Nrows * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    fsw_prof_begin("bwd_rankT_u32768_f32", st);
    // default: the streaming kernel (asynchronous copies into a shared-memory ring, 16 source rows per warp);
    // FSW_RANKT_STREAM=0 selects the register-staged kernel
    static const int stream_mode = [] { const char* e = getenv("FSW_RANKT_STREAM"); return e ? atoi(e) : 1; }();
    if (stream_mode != 0) {
        static const int rpw_env = [] { const char* e = getenv("FSW_RANKT_RPW"); return e ? atoi(e) : 16; }();
        const int rpw = rpw_env < 1 ? 1 : (rpw_env > 32 ? 32 : rpw_env);
        const int64_t swarps = fsw_cdiv(Nrows, rpw) * nchunks;
        const int64_t sblocks = fsw_cdiv(swarps, 4);
#define FSW_RANKS_LAUNCH(V_, D_, MINB_)                                                                                              \
    do {                                                                                                                             \
        auto kern = dEp ? fsw_rank_bwdS_kernel<V_, D_, MINB_, true> : fsw_rank_bwdS_kernel<V_, D_, MINB_, false>;                    \
        const size_t smem = (size_t)4 * FswRankStream<V_, D_>::PER_WARP;                                                             \
        FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                                \
        kern<<<(unsigned)sblocks, 128, smem, st>>>(a, Nrows, nchunks, rpw, tptr, tseg, tslot, tn, ranks, ldr, ga_buf, dXp, dEp);     \
    } while (0)
        if (wide) {
            // measured at configs[3] (ms per launch; register-staged kernel 22.35): ring of 6 pairs x 5 CTAs 17.97,
            // 8 x 4 CTAs 18.74, 5 x 6 CTAs 18.28, 12 x 3 CTAs 20.66; 8 / 16 / 32 rows per warp within 0.3 %
            if (stream_mode == 2) FSW_RANKS_LAUNCH(8, 8, 4);
            else FSW_RANKS_LAUNCH(8, 6, 5);
        } else {
            FSW_RANKS_LAUNCH(4, 8, 6);
        }
#undef FSW_RANKS_LAUNCH
        fsw_prof_end(st);
        FSW_CHECK_LAUNCH("fsw_rank_bwdS_kernel");
        return FSW_OK;
    }
    // 4 CTAs per SM: 125 registers, no spills (5 CTAs at 96 registers spill inside the pair loop: 28.2 vs 22.9 ms per launch)
    // L2 prefetch of every pair row of a block of triples: measured SLOWER (25.2 vs 22.9 ms per launch at configs[3]) - the
    // kernel is not waiting on those loads; kept behind a knob for other shapes
    static const bool pf = getenv("FSW_RANKT_PREFETCH") != nullptr;
#define FSW_RANKT_LAUNCH(V_, U_, PF_) \
    fsw_rank_bwdT_kernel<V_, U_, 4, PF_><<<(unsigned)blocks, 128, 0, st>>>(a, Nrows, nchunks, tptr, tseg, tslot, tn, ranks, ldr, ga_buf, dXp, dEp, tab_u)
    if (wide) {
        if (pf) FSW_RANKT_LAUNCH(8, 2, true);
        else FSW_RANKT_LAUNCH(8, 2, false);
    } else {
        if (pf) FSW_RANKT_LAUNCH(4, 4, true);
        else FSW_RANKT_LAUNCH(4, 4, false);
    }
#undef FSW_RANKT_LAUNCH
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_rank_bwdT_kernel");
    return FSW_OK;
}
