// Compile-time sorting network (Batcher odd-even merge sort) over register arrays + tiny helpers shared by
// the embedding kernels.
#pragma once
#include <string>
#include <utility>

#include "fsw_common.cuh"

#define FSW_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------------
// Batcher odd-even merge sort network over NP compile-time indexed slots (NP power of two).
// ---------------------------------------------------------------------------------------------------
// The comparator list is produced at compile time and applied through a fold expression, so every
// index is a constant and the arrays stay in registers whatever NP is.
template <int NP>
struct FswNet {
    static constexpr int kMax = (NP <= 4) ? 8 : NP * 10;  // >= number of comparators (543 for NP = 64)
    struct Pairs {
        int a[kMax];
        int b[kMax];
        int n;
    };
    static constexpr Pairs make() {
        Pairs P{};
        int c = 0;
        for (int p = 1; p < NP; p <<= 1)
            for (int k = p; k >= 1; k >>= 1)
                for (int j = k % p; j <= NP - 1 - k; j += 2 * k)
                    for (int i = 0; i <= ((k - 1) < (NP - j - k - 1) ? (k - 1) : (NP - j - k - 1)); ++i)
                        if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) {
                            P.a[c] = i + j;
                            P.b[c] = i + j + k;
                            ++c;
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
            if (j < cnt) v = __ldg(xp + (int64_t)row * a.ldp);
            key[j] = v;
        }
    } else {
        const T* __restrict__ xr = xp + ebase * a.ldp;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = Num<T>::big();
            if (j < cnt) v = __ldg(xr + (int64_t)j * a.ldp);
            key[j] = v;
        }
    }
    if (a.Ep) {
        const T* __restrict__ er = a.Ep + ebase * a.ldp + kk;
        T ep[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            T v = (T)0;
            if (j < cnt) v = __ldg(er + (int64_t)j * a.ldp);
            ep[j] = v;
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) key[j] += ep[j];  // +big stays +big (finite + 0)
    }
}
