// K2, uniform-weight fp32 forward with the whole sort in registers: packed keys, L cooperating lanes per slice.
//
//   fsw_coop_fwd_kernel<R, L, HAS_COL, SAVE_RANK>
//     A (segment, slice) is sorted by L lanes of one warp holding R elements each (R * L slots; L = 1, 2, 4, 8;
//     R = 32 when L > 1).  A warp therefore works on 32 / L consecutive slices of one segment.
//     * packed keys: the network sorts ONE 32-bit word per element - the order-preserving integer image of the
//       key with its low log2(R L) bits replaced by the element index.  A comparator is two integer min/max
//       instructions instead of a compare and four selects, and no index array travels with the keys.
//     * each lane sorts its R words with a merge-exchange network, then the lanes are merged with bitonic
//       merges whose cross-lane steps are shuffles (static register indices) - the code stays small enough
//       for the instruction cache and needs ~R registers, so occupancy is high (an earlier monolithic
//       128-element network per thread stalled 70 % of its cycles on instruction fetch, profiles/r1).
//     * the dropped low bits are restored exactly: the full keys stay in shared memory; adjacent sorted
//       positions whose truncated keys coincide (min over XORs per block of positions, voted across the warp)
//       are bubble-ordered by their full keys.  Equal full keys keep element order, as a stable sort would.
//     * Fourier coefficients come from the global tables cos(pi xi (2j+1)/n) and d/dxi
//       (fsw_build_coef_tables); the L partial sums of a slice are combined with shuffles.
//     * SAVE_RANK: the slot of each consumed full key is overwritten by its sorted position and streamed out
//       as uint16 rank[(e0+e), k] for the sort-free backward (fsw_embed_small.cu).
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

// compare-exchange of two packed words held by this lane
#define FSW_PK_CMPX(x, y)        \
    {                            \
        const int lo_ = min(x, y); \
        y = max(x, y);           \
        x = lo_;                 \
    }

template <int R, int L, bool HAS_COL, bool SAVE_RANK>
__global__ void __launch_bounds__(128, 5) fsw_coop_fwd_kernel(SegArgs<float> a, int seg_lo, int seg_hi, int G, int nchunks,
                                                           float* __restrict__ out, int64_t ld_out, int64_t out_col0,
                                                           const float* __restrict__ bias, unsigned short* __restrict__ ranks,
                                                           int64_t ldr, float* __restrict__ dxi_out, int64_t ld_dxi,
                                                           const float* __restrict__ gtab_c, const float* __restrict__ gtab_t) {
    static_assert(L == 1 || (R & (R - 1)) == 0, "bitonic cross-lane merges need a power-of-two run length");
    static_assert((L & (L - 1)) == 0 && L <= 32, "lanes per slice: power of two");
    constexpr int SW = 32 / L;            // slices per warp
    constexpr int NS = R * L;             // element slots per (segment, slice)
    constexpr int NC = (NS + 31) / 32;    // column-id registers per lane
    constexpr int IDXB = fsw_clog2(NS);   // low bits that carry the element index
    constexpr int IMASK = (1 << IDXB) - 1;
    constexpr int LOGL = fsw_clog2(L);
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int g = lane / SW;   // which run of the slice this lane holds
    const int sl = lane % SW;  // slice within the warp
    float* fkw = reinterpret_cast<float*>(fsw_smem_raw) + (size_t)warp * NS * SW;  // full keys [NS][SW] of this warp
    float* fks = fkw + sl;                                                          // column of this lane's slice
    float* fkl = fkw + lane;                                                        // element i*L+g of this lane: fkl[i*32]
    int* fksi = reinterpret_cast<int*>(fks);
    const int* fkli = reinterpret_cast<const int*>(fkl);

    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first64 = (int64_t)seg_lo + item * G;
    if (first64 >= seg_hi) return;
    const int first = (int)first64;
    const int last = (int)((first64 + G < seg_hi) ? first64 + G : seg_hi);
    const int k = chunk * SW + sl;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const float xi = __ldg(a.freqs + kk);
    const double xid = (double)xi;
    const float bk = (bias != nullptr) ? __ldg(bias + kk) : 0.f;
    const int ldp = (int)a.ldp;
    const int ldb = ldp * (int)sizeof(float);
    const char* xp_bytes = reinterpret_cast<const char*>(a.Xp + kk);
    const char* ep_bytes = a.Ep ? reinterpret_cast<const char*>(a.Ep + kk) : nullptr;
    const bool want_dxi = SAVE_RANK && dxi_out != nullptr;

    // software pipeline over segments: order two ahead, row range one ahead, column ids one ahead
    PkMeta cur, nx1;
    int s2;
    int c[NC];
    cur.s = fsw_pk_order(a, first, last);
    fsw_pk_range(a, cur.s, cur.e0, cur.n);
    fsw_pk_cols<NC, HAS_COL>(a.col, cur.e0, cur.n, lane, c);
    nx1.s = fsw_pk_order(a, first + 1, last);
    fsw_pk_range(a, nx1.s, nx1.e0, nx1.n);
    s2 = fsw_pk_order(a, first + 2, last);

    int n_prev = -1;
    float A = 0.f, A0 = 0.f, A0p = 0.f;
    const float* tc = gtab_c;
    const float* tt = gtab_t;

    for (int q = first; q < last; ++q) {
        const int n = cur.n;
        int s[R];
        // ---- gather: lane (g, sl) takes elements e = i L + g; all its loads are in flight together ----
        if (HAS_COL) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int row = __shfl_sync(FSW_FULL, c[(i * L) >> 5], ((i * L) & 31) + g);
                s[i] = 0;
                if (i * L + g < n) s[i] = __float_as_int(__ldg(reinterpret_cast<const float*>(xp_bytes + fsw_rowoff(row, ldb))));
            }
        } else {
            const char* __restrict__ base = xp_bytes + (cur.e0 + g) * ldb;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                s[i] = 0;
                if (i * L + g < n) s[i] = __float_as_int(__ldg(reinterpret_cast<const float*>(base + fsw_rowoff(i * L, ldb))));
            }
        }
        if (ep_bytes != nullptr) {  // edge features: per-slot additive projection (rare path)
            const char* __restrict__ eb = ep_bytes + (cur.e0 + g) * ldb;
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (i * L + g < n)
                    s[i] = __float_as_int(__int_as_float(s[i]) + __ldg(reinterpret_cast<const float*>(eb + fsw_rowoff(i * L, ldb))));
        }
        // prefetches for the following segments (their addresses were loaded one iteration ago)
        int cn[NC];
        fsw_pk_cols<NC, HAS_COL>(a.col, nx1.e0, nx1.n, lane, cn);
        PkMeta nx2;
        nx2.s = s2;
        fsw_pk_range(a, s2, nx2.e0, nx2.n);
        const int s3 = fsw_pk_order(a, q + 3, last);

        if (n != n_prev) {
            const double u = xid / (double)n;
            const float wn = (float)(1.0 / (double)n);
            fsw_amplitude<float, SAVE_RANK>(u, wn, xi, A0, A0p);
            A = (1.f + xi) * A0;
            const int64_t row0 = (int64_t)n * (n - 1) / 2 + g * R;  // table row of this lane's first sorted position
            tc = gtab_c + row0 * ldp + kk;
            tt = gtab_t + row0 * ldp + kk;
            n_prev = n;
        }

        // ---- pack: full key to shared memory, sortable image | element index to the register ----
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int e = i * L + g;
            const bool valid = e < n;
            const float v = valid ? __int_as_float(s[i]) + 0.0f : 0.f;  // -0 -> +0: zeros tie (torch.sort compares values)
            fkl[i * 32] = v;                                            // = fkw[e * SW + sl]
            const int b = __float_as_int(v);
            const int t = b ^ ((b >> 31) & 0x7fffffff);
            // padding slots: above +inf, every slot its own group, in slot order
            s[i] = valid ? ((t & ~IMASK) | e) : (0x7f800000 | (e << IDXB) | e);
        }
        if (L > 1) __syncwarp();
        // ---- sort the lane's run, then merge the runs of the L lanes ----
        fsw_sort_network<R>([&](int i, int l) { FSW_PK_CMPX(s[i], s[l]); });
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
            // the lane now holds a bitonic run: half-cleaners at distance R/2 .. 1
            fsw_static_for<fsw_clog2(R)>([&](auto hc) {
                constexpr int h = R >> (decltype(hc)::value + 1);
                fsw_static_for<R / 2>([&](auto ic) {
                    constexpr int t = decltype(ic)::value;
                    constexpr int i = (t / h) * 2 * h + (t % h);
                    FSW_PK_CMPX(s[i], s[i + h]);
                });
            });
        });
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
        float acc = 0.f, acc2 = 0.f;
        const int p0 = g * R;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (p0 + i < n) {
                const int idx = s[i] & IMASK;
                const float key = fks[idx * SW];
                acc = fmaf(key, __ldg(tc + fsw_rowoff(i, ldp)), acc);
                if (want_dxi) acc2 = fmaf(key, __ldg(tt + fsw_rowoff(i, ldp)), acc2);
                if (SAVE_RANK) fksi[idx * SW] = p0 + i;
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
        if constexpr (SAVE_RANK) {
            if (L > 1) __syncwarp();
            unsigned short* rp = ranks + fsw_rowoff(cur.e0 + g, ldr) + k;
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (i * L + g < n && act) rp[fsw_rowoff(i * L, ldr)] = (unsigned short)fkli[i * 32];
        }
        if (L > 1) __syncwarp();
        // rotate the pipeline
        cur = nx1;
#pragma unroll
        for (int m = 0; m < NC; ++m) c[m] = cn[m];
        nx1 = nx2;
        s2 = s3;
    }
}

template <int R, int L, bool HAS_COL, bool SAVE_RANK>
int launch_coop_fwd(const SegArgs<float>& a, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0, const float* bias,
                    unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t,
                    cudaStream_t st) {
    constexpr int SW = 32 / L;
    const int nchunks = (a.K + SW - 1) / SW;
    int64_t G = (int64_t)(hi - lo) * nchunks / (148 * 64);
    if (G < 1) G = 1;
    if (G > 32) G = 32;
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    const size_t smem = (size_t)4 * R * 32 * sizeof(float);
    auto kern = fsw_coop_fwd_kernel<R, L, HAS_COL, SAVE_RANK>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const std::string label = std::string(SAVE_RANK ? "fwdr_coop_u" : "fwd_coop_u") + std::to_string(R * L) + "_f32";
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, 128, smem, st>>>(a, lo, hi, (int)G, nchunks, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_coop_fwd_kernel");
    return FSW_OK;
}

template <int R, int L>
int launch_coop(const SegArgs<float>& a, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0, const float* bias,
                unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t,
                cudaStream_t st) {
    const bool has_col = a.col != nullptr;
    if (ranks) {
        return has_col ? launch_coop_fwd<R, L, true, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st)
                       : launch_coop_fwd<R, L, false, true>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st);
    }
    return has_col ? launch_coop_fwd<R, L, true, false>(a, lo, hi, out, ld_out, out_col0, bias, nullptr, 0, nullptr, 0, gtab_c, gtab_t, st)
                   : launch_coop_fwd<R, L, false, false>(a, lo, hi, out, ld_out, out_col0, bias, nullptr, 0, nullptr, 0, gtab_c, gtab_t, st);
}

}  // namespace

// uniform-weight fp32 segments order[lo, hi) with n <= np, np in {64, 128, 256}; gtab_c (and gtab_t when dxi_out is
// given) must cover n <= np (fsw_build_coef_tables layout)
int fsw_packed_forward_u(const SegArgs<float>& a, int np, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0,
                         const float* bias, unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c,
                         const float* gtab_t, cudaStream_t st) {
    if (gtab_c == nullptr) return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: coefficient table missing");
    switch (np) {
        case 64: return launch_coop<32, 2>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st);
        case 128: return launch_coop<32, 4>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st);
        case 256: return launch_coop<32, 8>(a, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, st);
    }
    return fsw_fail(FSW_ERR_INVALID, "fsw_packed_forward_u: class %d", np);
}
