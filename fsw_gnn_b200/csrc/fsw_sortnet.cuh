// Compile-time sorting network (Batcher odd-even merge sort) over register arrays + tiny helpers shared by
// the embedding kernels.
#pragma once
#include <string>
#include <utility>

#include "fsw_common.cuh"

#define FSW_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------------
// Merge-exchange sorting network over NP compile-time indexed slots (any NP).
// ---------------------------------------------------------------------------------------------------
// The comparator list is produced at compile time and applied through a fold expression, so every
// index is a constant and the arrays stay in registers whatever NP is.
template <int NP>
struct FswNet {
    static constexpr int kMax = (NP <= 4) ? 8 : NP * 12;  // >= number of comparators (543 for NP = 64, 1471 for 128)
    struct Pairs {
        int a[kMax];
        int b[kMax];
        int n;
    };
    // Knuth, TAOCP 5.2.2 Algorithm M (merge exchange): valid for every NP, equals Batcher's odd-even
    // merge sort for powers of two (191 comparators at 32, 543 at 64; 127 at 24, 367 at 48).
    static constexpr Pairs make() {
        Pairs P{};
        int c = 0;
        if (NP >= 2) {
            int t = 0;
            while ((1 << t) < NP) ++t;
            for (int p = 1 << (t - 1); p > 0; p >>= 1) {
                int q = 1 << (t - 1), r = 0, d = p;
                while (true) {
                    for (int i = 0; i < NP - d; ++i)
                        if ((i & p) == r) {
                            P.a[c] = i;
                            P.b[c] = i + d;
                            ++c;
                        }
                    if (q == p) break;
                    d = q - p;
                    q >>= 1;
                    r = p;
                }
            }
        }
        P.n = c;
        return P;
    }
    static constexpr int count = make().n;
};

// scalar compile-time constants are usable from device code (aggregate constexpr members are not)
template <int NP, int I>
struct FswPair {
    static constexpr int a = FswNet<NP>::make().a[I];
    static constexpr int b = FswNet<NP>::make().b[I];
};

template <int NP, typename CE, int... Is>
__device__ __forceinline__ void fsw_sort_network_apply(CE&& ce, std::integer_sequence<int, Is...>) {
    (ce(FswPair<NP, Is>::a, FswPair<NP, Is>::b), ...);
}

template <int NP, typename CE>
__device__ __forceinline__ void fsw_sort_network(CE&& ce) {
    fsw_sort_network_apply<NP>(ce, std::make_integer_sequence<int, FswNet<NP>::count>{});
}

template <typename T>
__device__ __forceinline__ T fsw_ldg(const T* p) {
    return __ldg(p);
}


// ---------------------------------------------------------------------------------------------------
// Gather of up to NP projected keys of one segment chunk into registers.
//   rows e0+base .. e0+base+cnt-1 of the segment; lanes are slices (column kk of Xp).
// All loads are issued back to back before any of them is consumed (memory-level parallelism inside
// the warp: an earlier version interleaved each load with its use and ran at one DRAM round trip per
// element, see profiles/r1).  Column ids are loaded coalesced (lane j holds element j) and broadcast
// with shuffles; c0/c1 are returned for the scatter of the backward pass.  Slots >= cnt get +big.
// ---------------------------------------------------------------------------------------------------
template <typename T, int NP>
__device__ __forceinline__ void fsw_gather_keys(const SegArgs<T>& a, int64_t ebase, int cnt, int kk, int lane, T (&key)[NP],
                                                int& c0, int& c1) {
    c0 = 0;
    c1 = 0;
    const T* __restrict__ xp = a.Xp + kk;
    if (a.col) {
        if (lane < cnt) c0 = __ldg(a.col + ebase + lane);
        if (NP > 32 && lane + 32 < cnt) c1 = __ldg(a.col + ebase + 32 + lane);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int row = __shfl_sync(FSW_FULL, (j < 32) ? c0 : c1, j & 31);
            T v = Num<T>::big();
            if (j < cnt) v = __ldg(xp + fsw_rowoff(row, a.ldp));
            key[j] = v;
        }
    } else {
        const T* __restrict__ xr = xp + fsw_rowoff(ebase, a.ldp);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = Num<T>::big();
            if (j < cnt) v = __ldg(xr + fsw_rowoff(j, a.ldp));
            key[j] = v;
        }
    }
    if (a.Ep) {
        const T* __restrict__ er = a.Ep + fsw_rowoff(ebase, a.ldp) + kk;
        T ep[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = (T)0;
            if (j < cnt) v = __ldg(er + fsw_rowoff(j, a.ldp));
            ep[j] = v;
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) key[j] += ep[j];  // +big stays +big (finite + 0)
    }
}

// ---------------------------------------------------------------------------------------------------
// Lean gather used by the uniform-weight kernels: 5 instructions per element
//   SHFL (column id) + IMAD.WIDE (byte address) + LDG + ISETP/FSEL (mask slots >= cnt to +big).
// Column ids are loaded with a clamped index, so every lane holds a valid row and the loads need no
// predicates (slots >= cnt re-read the last element's row, an L1 hit).  Requires cnt >= 1.
// ---------------------------------------------------------------------------------------------------
template <int NP, bool HAS_COL>
__device__ __forceinline__ void fsw_load_cols(const int32_t* __restrict__ col, int64_t ebase, int cnt, int lane, int& c0, int& c1) {
    c0 = 0;
    c1 = 0;
    if (HAS_COL) {
        c0 = __ldg(col + ebase + min(lane, cnt - 1));
        if (NP > 32) c1 = __ldg(col + ebase + min(lane + 32, cnt - 1));
    }
}

template <typename T, int NP, bool HAS_COL>
__device__ __forceinline__ void fsw_gather_lean(const char* __restrict__ xp_bytes, int ldb, const char* __restrict__ ep_bytes,
                                                int64_t ebase, int cnt, int c0, int c1, T (&key)[NP]) {
    if (HAS_COL) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const int row = __shfl_sync(FSW_FULL, (j < 32) ? c0 : c1, j & 31);
            key[j] = __ldg(reinterpret_cast<const T*>(xp_bytes + fsw_rowoff(row, ldb)));   // one IMAD.WIDE per gathered row
        }
    } else {
        const char* __restrict__ base = xp_bytes + ebase * ldb;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = (T)0;
            if (j < cnt) v = __ldg(reinterpret_cast<const T*>(base + fsw_rowoff(j, ldb)));
            key[j] = v;
        }
    }
    if (ep_bytes != nullptr) {  // edge features: per-slot additive projection (rare path)
        const char* __restrict__ eb = ep_bytes + ebase * ldb;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = (T)0;
            if (j < cnt) v = __ldg(reinterpret_cast<const T*>(eb + fsw_rowoff(j, ldb)));
            key[j] += v;
        }
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) key[j] = (j < cnt) ? key[j] : Num<T>::big();
}

// ---------------------------------------------------------------------------------------------------
// Block-wide bitonic network over a tile [n_pad rows][32 lanes] (shared or global memory): rows = elements, lanes = slices.
// ---------------------------------------------------------------------------------------------------
template <typename KT, typename PT, bool HAS_PAY>
__device__ __forceinline__ void fsw_block_bitonic(KT* keys, PT* pay, int n_pad) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = warp; t < (n_pad >> 1); t += nw) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool asc = ((i & k) == 0);
                KT x = keys[i * 32 + lane], y = keys[l * 32 + lane];
                if ((x > y) == asc) {
                    keys[i * 32 + lane] = y;
                    keys[l * 32 + lane] = x;
                    if (HAS_PAY) {
                        PT px = pay[i * 32 + lane], py = pay[l * 32 + lane];
                        pay[i * 32 + lane] = py;
                        pay[l * 32 + lane] = px;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int fsw_next_pow2(int n) {
    int p = 2;
    while (p < n) p <<= 1;
    return p;
}

