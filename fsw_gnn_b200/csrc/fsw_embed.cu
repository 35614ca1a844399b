// K2 / K3: fused FSW embedding forward and backward over CSR segments - entry points and dispatch, plus the kernels
// for general (non-uniform) weights and fp64.
//
// Replaces FSW_embedding.forward_helper (fsw_embedding.py:894-1112) and the autograd functions it
// drives (class ag, :1232-2258; helpers sp, :2266-2775).  The segment plan sorts the segments into size buckets;
// every bucket range is served by one launch of the kernel family that fits it:
//
//   uniform weights, fp32 (the case of FSW_conv's default adjacency and of unit-weight point clouds)
//     n <= 32        fsw_small_fwd_kernel        (fsw_embed_small.cu)   thread per (segment, slice), register network
//     33 .. 512      fsw_coop_fwd_kernel         (fsw_embed_packed.cu)  L lanes per slice, packed keys; dense <= 1024
//     larger         fsw_medium_kernel           (fsw_embed_medium.cu)  CTA tile, register runs + merge path
//     backward       fsw_rank_bwdT_kernel / fsw_rank_bwd_dense_kernel (fsw_embed_small.cu): no sort, the forward
//                    recorded every element's sorted position; re-sorting kernels only beyond 32768 elements
//   general weights / fp64 (this file)
//     small  (n_eff <= 64 fp32 / 32 fp64): one THREAD owns one (segment, slice).  Lanes of a warp are 32
//          consecutive slices, so the gather Xp[col[e], k0..k0+31] is one coalesced 128-byte line per
//          element.  Keys (and raw weights) live in registers and are sorted with a data-oblivious merge-exchange
//          network (no shuffles, no divergence).
//     generic (anything larger): one CTA owns a tile [n_pad][32 slices] in shared memory (or in an
//          L2-resident global scratch when it does not fit) and runs a block-wide bitonic network
//          with rows as elements and lanes as slices (bank = lane, conflict free).
//
// Numerics (SURVEY.md 7 hard part 2): the phase xi*(2C - w) is formed in fp64 from fp64 cumulative
// weights and reduced mod 2 before it is rounded to fp32; everything else is fp32 for fp32 inputs.
// The sparse product form D_j = 2 w sinc(xi w) cos(pi xi (2C - w)) (fsw_embedding.py:1047-1075) is
// used for both value and gradient; it is well conditioned for every xi >= 0.
#include <cstdlib>

#include "fsw_sortnet.cuh"

// ---------------------------------------------------------------------------------------------------
// Small path, forward
// ---------------------------------------------------------------------------------------------------
template <typename T, int NP, bool UNIFORM>
__global__ void __launch_bounds__(128) fsw_fwd_small_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks,
                                                            T* __restrict__ out, int64_t ld_out, int64_t out_col0,
                                                            const T* __restrict__ bias) {
    const int lane = threadIdx.x & 31;
    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first = (int64_t)seg_lo + item * G;
    if (first >= seg_hi) return;
    const int last = (int)((first + G < seg_hi) ? first + G : seg_hi);
    const int k = chunk * 32 + lane;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const T xi = fsw_ldg(a.freqs + kk);
    const double xid = (double)xi;
    const T bk = (bias != nullptr) ? fsw_ldg(bias + kk) : (T)0;

    T coef[NP];
    T A = (T)0;
    int n_prev = -1;
    (void)coef;
    (void)A;
    (void)n_prev;

    for (int q = (int)first; q < last; ++q) {
        const int s = a.order ? a.order[q] : q;
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);

        // ---- gather (column ids are loaded coalesced, then broadcast by shuffle; all loads in flight at once) ----
        int c0, c1;
        T key[NP];
        fsw_gather_keys<T, NP>(a, e0, n, kk, lane, key, c0, c1);

        T result;
        if constexpr (UNIFORM) {
            if (n != n_prev) {
                const double u = xid / (double)n;
#pragma unroll
                for (int j = 0; j < NP; ++j)
                    coef[j] = (j < n) ? Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * j + 1))) : (T)0;
                T a0, a0p;
                fsw_amplitude<T, false>(u, (T)(1.0 / (double)n), xi, a0, a0p);
                A = ((T)1 + xi) * a0;
                n_prev = n;
            }
            fsw_sort_network<NP>([&](int i, int l) {
                T x = key[i], y = key[l];
                key[i] = fmin(x, y);
                key[l] = fmax(x, y);
            });
            T acc = (T)0;
#pragma unroll
            for (int j = 0; j < NP; ++j) acc = fma(key[j], coef[j], acc);
            result = A * acc;
        } else {
            const int n_eff = a.info[s] & FSW_INFO_NMASK;
            const bool padded = n_eff > n;
            const double Ts = a.mass[s];
            const double invS = 1.0 / fmax(Ts, a.thresh);
            const double padw = a.thresh - Ts;  // raw weight of the deficit pad point (x = 0)
            T w0 = (T)1, w1 = (T)1;
            if (a.W) {
                w0 = (lane < n) ? a.W[e0 + lane] : (T)0;
                if (NP > 32) w1 = (lane + 32 < n) ? a.W[e0 + 32 + lane] : (T)0;
            }
            T pay[NP];  // raw weight; -1 flags the pad point, 0 for unused slots
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                T w = __shfl_sync(FSW_FULL, (j < 32) ? w0 : w1, j & 31);
                if (j >= n) w = (T)0;
                if (padded && j == n) {
                    w = (T)-1;
                    key[j] = (T)0;
                }
                pay[j] = w;
            }
            fsw_sort_network<NP>([&](int i, int l) {
                T x = key[i], y = key[l];
                T px = pay[i], py = pay[l];
                bool sw = x > y;
                key[i] = sw ? y : x;
                key[l] = sw ? x : y;
                pay[i] = sw ? py : px;
                pay[l] = sw ? px : py;
            });
            double Craw = 0.0;
            T acc = (T)0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (j < n_eff) {
                    const double wr = (pay[j] < (T)0) ? padw : (double)pay[j];
                    Craw += wr;
                    const double wn = wr * invS;
                    const double phi = xid * (2.0 * Craw * invS - wn);
                    const T c = Num<T>::cospi_(Num<T>::reduce(phi));
                    T aj, ajp;
                    fsw_amplitude<T, false>(xid * wn, (T)wn, xi, aj, ajp);
                    acc = fma(key[j], aj * c, acc);
                }
            }
            result = ((T)1 + xi) * acc;
        }
        if (act) out[(int64_t)s * ld_out + out_col0 + k] = result + bk;
    }
}

// ---------------------------------------------------------------------------------------------------
// Small path, backward.  Dynamic shared memory per warp: [3][NP][32] T  (un-permute buffer, cos table,
// d/dxi table).
// ---------------------------------------------------------------------------------------------------
template <typename T, int NP, bool UNIFORM, bool NEED_DXI>
__global__ void __launch_bounds__(128) fsw_bwd_small_kernel(SegArgs<T> a, int seg_lo, int seg_hi, int G, int nchunks,
                                                            const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                            T* __restrict__ dXp, T* __restrict__ dEp,
                                                            double* __restrict__ dfreqs) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    T* sm_val = reinterpret_cast<T*>(fsw_smem_raw) + (size_t)warp * 3 * NP * 32;
    T* sm_c = sm_val + NP * 32;
    T* sm_t = sm_c + NP * 32;
    (void)sm_c;
    (void)sm_t;

    const int64_t wglobal = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const int64_t item = wglobal / nchunks;
    const int chunk = (int)(wglobal - item * nchunks);
    const int64_t first = (int64_t)seg_lo + item * G;
    if (first >= seg_hi) return;
    const int last = (int)((first + G < seg_hi) ? first + G : seg_hi);
    const int k = chunk * 32 + lane;
    const bool act = k < a.K;
    const int kk = act ? k : a.K - 1;
    const T xi = fsw_ldg(a.freqs + kk);
    const double xid = (double)xi;

    double dxi_acc = 0.0;
    T A0 = (T)0, A0p = (T)0;
    int n_prev = -1;
    (void)A0p;
    (void)n_prev;

    for (int q = (int)first; q < last; ++q) {
        const int s = a.order ? a.order[q] : q;
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);
        const int n_eff = UNIFORM ? n : (a.info[s] & FSW_INFO_NMASK);
        const bool padded = n_eff > n;

        int c0, c1;
        T key[NP];
        int idx[NP];
        fsw_gather_keys<T, NP>(a, e0, n, kk, lane, key, c0, c1);
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            if (!UNIFORM && padded && j == n) key[j] = (T)0;
            idx[j] = j;
        }
        fsw_sort_network<NP>([&](int i, int l) {
            T x = key[i], y = key[l];
            int px = idx[i], py = idx[l];
            bool sw = x > y;
            key[i] = sw ? y : x;
            key[l] = sw ? x : y;
            idx[i] = sw ? py : px;
            idx[l] = sw ? px : py;
        });

        const T gk = act ? g[(int64_t)s * ld_g + g_col0 + k] : (T)0;
        const T Gk = gk * ((T)1 + xi);

        if constexpr (UNIFORM) {
            if (n != n_prev) {
                const double u = xid / (double)n;
                const T wn = (T)(1.0 / (double)n);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    if (j < n) {
                        const T r = Num<T>::reduce(u * (double)(2 * j + 1));
                        sm_c[j * 32 + lane] = Num<T>::cospi_(r);
                        if (NEED_DXI) sm_t[j * 32 + lane] = (T)M_PI * wn * (T)(2 * j + 1) * Num<T>::sinpi_(r);
                    }
                }
                fsw_amplitude<T, NEED_DXI>(u, wn, xi, A0, A0p);
                n_prev = n;
            }
            T Sc = (T)0, Ss = (T)0;
            const T GA = Gk * A0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (j < n) {
                    const T c = sm_c[j * 32 + lane];
                    sm_val[idx[j] * 32 + lane] = GA * c;
                    if (NEED_DXI) {
                        Sc = fma(key[j], c, Sc);
                        Ss = fma(key[j], sm_t[j * 32 + lane], Ss);
                    }
                }
            }
            if (NEED_DXI) dxi_acc += (double)(gk * (A0 * Sc + ((T)1 + xi) * (A0p * Sc - A0 * Ss)));
        } else {
            const double Ts = a.mass[s];
            const double invS = 1.0 / fmax(Ts, a.thresh);
            const double padw = a.thresh - Ts;
            T w0 = (T)1, w1 = (T)1;
            if (a.W) {
                w0 = (lane < n) ? a.W[e0 + lane] : (T)0;
                if (NP > 32) w1 = (lane + 32 < n) ? a.W[e0 + 32 + lane] : (T)0;
            }
            // raw weights in ORIGINAL order go through shared memory so that they can be fetched by idx
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                T w = __shfl_sync(FSW_FULL, (j < 32) ? w0 : w1, j & 31);
                sm_c[j * 32 + lane] = (j < n) ? w : (T)0;
            }
            __syncwarp();
            double Craw = 0.0;
            T sPD = (T)0, sPdD = (T)0;
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                if (j < n_eff) {
                    const int id = idx[j];
                    const double wr = (id >= n) ? padw : (double)sm_c[id * 32 + lane];
                    Craw += wr;
                    const double wn = wr * invS;
                    const double two_c_minus_w = 2.0 * Craw * invS - wn;
                    const T r = Num<T>::reduce(xid * two_c_minus_w);
                    const T c = Num<T>::cospi_(r);
                    T aj, ajp = (T)0;
                    fsw_amplitude<T, NEED_DXI>(xid * wn, (T)wn, xi, aj, ajp);
                    const T D = aj * c;
                    sm_val[id * 32 + lane] = Gk * D;
                    if (NEED_DXI) {
                        const T sn = Num<T>::sinpi_(r);
                        const T dD = ajp * c - aj * (T)M_PI * (T)two_c_minus_w * sn;
                        sPD = fma(key[j], D, sPD);
                        sPdD = fma(key[j], dD, sPdD);
                    }
                }
            }
            if (NEED_DXI) dxi_acc += (double)(gk * (sPD + ((T)1 + xi) * sPdD));
        }
        __syncwarp();
        // ---- coalesced scatter in ORIGINAL element order ----
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            int64_t row = e0 + i;
            if (a.col) row = __shfl_sync(FSW_FULL, (i < 32) ? c0 : c1, i & 31);
            if (i < n && act) {
                const T v = sm_val[i * 32 + lane];
                if (a.col)
                    atomicAdd(dXp + row * a.ldp + k, v);
                else
                    dXp[row * a.ldp + k] = v;
                if (dEp) dEp[(e0 + i) * a.ldp + k] = v;
            }
        }
        __syncwarp();
    }
    if (NEED_DXI && act) atomicAdd(dfreqs + k, dxi_acc);
}

// ---------------------------------------------------------------------------------------------------
// Generic path: block-wide bitonic network (fsw_block_bitonic, fsw_sortnet.cuh), rows = elements, lanes = slices.
// ---------------------------------------------------------------------------------------------------
// forward.  MODE 0: uniform weights (keys only).  MODE 1: general (payload = raw weight, -1 = pad point)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) fsw_fwd_generic_kernel(SegArgs<T> a, int seg_lo, int nchunks, int64_t ntiles,
                                                              T* __restrict__ out, int64_t ld_out, int64_t out_col0,
                                                              const T* __restrict__ bias, int cap_rows,
                                                              unsigned char* gscratch) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const size_t tile_elems = (size_t)cap_rows * 32;
    unsigned char* base = gscratch ? gscratch + (size_t)blockIdx.x * tile_elems * sizeof(T) * (MODE == 1 ? 2 : 1) : fsw_smem_raw;
    T* keys = reinterpret_cast<T*>(base);
    T* pay = keys + tile_elems;
    (void)pay;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int q = seg_lo + (int)(tile / nchunks);
        const int chunk = (int)(tile % nchunks);
        const int s = a.order ? a.order[q] : q;
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);
        const int n_eff = (MODE == 0) ? n : (a.info[s] & FSW_INFO_NMASK);
        const bool padded = n_eff > n;
        const int n_pad = fsw_next_pow2(n_eff);
        const int k = chunk * 32 + lane;
        const bool act = k < a.K;
        const int kk = act ? k : a.K - 1;
        const T xi = fsw_ldg(a.freqs + kk);
        const double xid = (double)xi;

        for (int r = warp; r < n_pad; r += nw) {
            T v = Num<T>::big();
            T w = (T)0;
            if (r < n) {
                const int64_t row = a.col ? (int64_t)a.col[e0 + r] : e0 + r;
                v = fsw_ldg(a.Xp + row * a.ldp + kk);
                if (a.Ep) v += fsw_ldg(a.Ep + (e0 + r) * a.ldp + kk);
                if (MODE == 1) w = a.W ? a.W[e0 + r] : (T)1;
            } else if (MODE == 1 && padded && r == n) {
                v = (T)0;
                w = (T)-1;
            }
            keys[r * 32 + lane] = v;
            if (MODE == 1) pay[r * 32 + lane] = w;
        }
        __syncthreads();
        fsw_block_bitonic<T, T, MODE == 1>(keys, pay, n_pad);

        const int rpw = (n_eff + nw - 1) / nw;
        const int r0 = warp * rpw;
        const int r1 = (r0 + rpw < n_eff) ? r0 + rpw : n_eff;
        double acc = 0.0;
        if (MODE == 0) {
            const double u = xid / (double)n;
            for (int r = r0; r < r1; ++r) {
                const T c = Num<T>::cospi_(Num<T>::reduce(u * (double)(2 * r + 1)));
                acc += (double)(keys[r * 32 + lane] * c);
            }
            T a0, a0p;
            fsw_amplitude<T, false>(u, (T)(1.0 / (double)n), xi, a0, a0p);
            acc *= (double)(((T)1 + xi) * a0);
        } else {
            const double Ts = a.mass[s];
            const double invS = 1.0 / fmax(Ts, a.thresh);
            const double padw = a.thresh - Ts;
            double part = 0.0;
            for (int r = r0; r < r1; ++r) {
                const T w = pay[r * 32 + lane];
                part += (w < (T)0) ? padw : (double)w;
            }
            red[warp][lane] = part;
            __syncthreads();
            double Craw = 0.0;
            for (int w2 = 0; w2 < warp; ++w2) Craw += red[w2][lane];
            __syncthreads();
            for (int r = r0; r < r1; ++r) {
                const T w = pay[r * 32 + lane];
                const double wr = (w < (T)0) ? padw : (double)w;
                Craw += wr;
                const double wn = wr * invS;
                const T c = Num<T>::cospi_(Num<T>::reduce(xid * (2.0 * Craw * invS - wn)));
                T aj, ajp;
                fsw_amplitude<T, false>(xid * wn, (T)wn, xi, aj, ajp);
                acc += (double)(keys[r * 32 + lane] * (aj * c));
            }
            acc *= (double)((T)1 + xi);
        }
        red[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            double tot = 0.0;
            for (int w2 = 0; w2 < nw; ++w2) tot += red[w2][lane];
            if (act) out[(int64_t)s * ld_out + out_col0 + k] = (T)tot + ((bias != nullptr) ? bias[k] : (T)0);
        }
        __syncthreads();
    }
}

// backward generic.  Tiles: keys (T) and idx (int32).  After the key sort and the evaluation the key
// tile holds dL/dp in SORTED order; a second network sorts by idx to restore the ORIGINAL order so
// that the scatter to dXp is one coalesced row per element.
template <typename T, bool UNIFORM, bool NEED_DXI>
__global__ void __launch_bounds__(256) fsw_bwd_generic_kernel(SegArgs<T> a, int seg_lo, int nchunks, int64_t ntiles,
                                                              const T* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                              T* __restrict__ dXp, T* __restrict__ dEp,
                                                              double* __restrict__ dfreqs, int cap_rows,
                                                              unsigned char* gscratch) {
    extern __shared__ __align__(16) unsigned char fsw_smem_raw[];
    __shared__ double red[8][32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const size_t tile_elems = (size_t)cap_rows * 32;
    unsigned char* base = gscratch ? gscratch + (size_t)blockIdx.x * tile_elems * (sizeof(T) + sizeof(int)) : fsw_smem_raw;
    T* keys = reinterpret_cast<T*>(base);
    int* idx = reinterpret_cast<int*>(keys + tile_elems);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int q = seg_lo + (int)(tile / nchunks);
        const int chunk = (int)(tile % nchunks);
        const int s = a.order ? a.order[q] : q;
        int64_t e0;
        int n;
        fsw_seg_range(a, s, e0, n);
        const int n_eff = UNIFORM ? n : (a.info[s] & FSW_INFO_NMASK);
        const bool padded = n_eff > n;
        const int n_pad = fsw_next_pow2(n_eff);
        const int k = chunk * 32 + lane;
        const bool act = k < a.K;
        const int kk = act ? k : a.K - 1;
        const T xi = fsw_ldg(a.freqs + kk);
        const double xid = (double)xi;
        const T gk = act ? g[(int64_t)s * ld_g + g_col0 + k] : (T)0;
        const T Gk = gk * ((T)1 + xi);

        for (int r = warp; r < n_pad; r += nw) {
            T v = Num<T>::big();
            if (r < n) {
                const int64_t row = a.col ? (int64_t)a.col[e0 + r] : e0 + r;
                v = fsw_ldg(a.Xp + row * a.ldp + kk);
                if (a.Ep) v += fsw_ldg(a.Ep + (e0 + r) * a.ldp + kk);
            } else if (!UNIFORM && padded && r == n) {
                v = (T)0;
            }
            keys[r * 32 + lane] = v;
            idx[r * 32 + lane] = r;
        }
        __syncthreads();
        fsw_block_bitonic<T, int, true>(keys, idx, n_pad);

        const int rpw = (n_eff + nw - 1) / nw;
        const int r0 = warp * rpw;
        const int r1 = (r0 + rpw < n_eff) ? r0 + rpw : n_eff;
        double dxi_local = 0.0;
        if (UNIFORM) {
            const double u = xid / (double)n;
            const T wn = (T)(1.0 / (double)n);
            T A0, A0p = (T)0;
            fsw_amplitude<T, NEED_DXI>(u, wn, xi, A0, A0p);
            double Sc = 0.0, Ss = 0.0;
            for (int r = r0; r < r1; ++r) {
                const T rr = Num<T>::reduce(u * (double)(2 * r + 1));
                const T c = Num<T>::cospi_(rr);
                const T p = keys[r * 32 + lane];
                keys[r * 32 + lane] = Gk * A0 * c;
                if (NEED_DXI) {
                    Sc += (double)(p * c);
                    Ss += (double)(p * ((T)M_PI * wn * (T)(2 * r + 1) * Num<T>::sinpi_(rr)));
                }
            }
            if (NEED_DXI)
                dxi_local = (double)gk * ((double)A0 * Sc + (1.0 + xid) * ((double)A0p * Sc - (double)A0 * Ss));
        } else {
            const double Ts = a.mass[s];
            const double invS = 1.0 / fmax(Ts, a.thresh);
            const double padw = a.thresh - Ts;
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
            double sPD = 0.0, sPdD = 0.0;
            for (int r = r0; r < r1; ++r) {
                const int id = idx[r * 32 + lane];
                const double wr = (id >= n) ? padw : (a.W ? (double)a.W[e0 + id] : 1.0);
                Craw += wr;
                const double wn = wr * invS;
                const double tcw = 2.0 * Craw * invS - wn;
                const T rr = Num<T>::reduce(xid * tcw);
                const T c = Num<T>::cospi_(rr);
                T aj, ajp = (T)0;
                fsw_amplitude<T, NEED_DXI>(xid * wn, (T)wn, xi, aj, ajp);
                const T D = aj * c;
                const T p = keys[r * 32 + lane];
                keys[r * 32 + lane] = Gk * D;
                if (NEED_DXI) {
                    const T sn = Num<T>::sinpi_(rr);
                    const T dD = ajp * c - aj * (T)M_PI * (T)tcw * sn;
                    sPD += (double)(p * D);
                    sPdD += (double)(p * dD);
                }
            }
            if (NEED_DXI) dxi_local = (double)gk * (sPD + (1.0 + xid) * sPdD);
        }
        if (NEED_DXI) {
            red[warp][lane] = dxi_local;
            __syncthreads();
            if (warp == 0 && act) {
                double tot = 0.0;
                for (int w2 = 0; w2 < nw; ++w2) tot += red[w2][lane];
                atomicAdd(dfreqs + k, tot);
            }
        }
        // rows >= n_eff keep idx >= n_eff (they were sorted to the end with +big keys); give them 0 gradient
        for (int r = n_eff + warp; r < n_pad; r += nw) keys[r * 32 + lane] = (T)0;
        __syncthreads();
        fsw_block_bitonic<int, T, true>(idx, keys, n_pad);
        for (int r = warp; r < n; r += nw) {
            if (act) {
                const T v = keys[r * 32 + lane];
                if (a.col)
                    atomicAdd(dXp + (int64_t)a.col[e0 + r] * a.ldp + k, v);
                else
                    dXp[(e0 + r) * a.ldp + k] = v;
                if (dEp) dEp[(e0 + r) * a.ldp + k] = v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Host dispatch
// ---------------------------------------------------------------------------------------------------
namespace {

const int kSmemBudget = 200 * 1024;  // dynamic shared memory we are willing to ask for per CTA

struct ClassRange {
    int lo, hi;  // bucket range [lo, hi]
    int np;
};
const ClassRange kSmall[] = {{0, 4, 4}, {5, 8, 8}, {9, 16, 16}, {17, 32, 32}, {33, 64, 64}};

template <typename T>
int max_small_np() {
    return sizeof(T) == 4 ? 64 : 32;
}

struct SizeRange {
    int lo, hi;  // size-bucket range [lo, hi]
    int cap;     // upper bound of n_eff in the range
};

// ranges served by the generic kernels, beyond the small-path limit `msn`
int generic_ranges(int msn, int64_t max_n_eff, SizeRange* out) {
    int c = 0;
    if (msn < 64) out[c++] = {msn + 1, 64, 64};
    out[c++] = {65, 128, 128};
    out[c++] = {129, 256, 256};
    out[c++] = {257, 512, 512};
    out[c++] = {513, 513, 1024};
    out[c++] = {514, 514, 2048};
    out[c++] = {515, 515, 4096};
    out[c++] = {516, 516, 8192};
    out[c++] = {517, 517, 32768};
    int64_t big = 65536;   // hubs: one class, sized for the largest segment
    while (big < max_n_eff) big <<= 1;
    out[c++] = {518, 518, (int)big};
    return c;
}

int pick_G(int64_t cnt, int nchunks) {
    int64_t g = cnt * nchunks / (148 * 24);
    if (g < 1) g = 1;
    if (g > 32) g = 32;
    return (int)g;
}

template <typename T, int NP, bool UNIFORM>
int launch_fwd_small(const SegArgs<T>& a, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                     cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int G = pick_G(hi - lo, nchunks);
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    static const std::string label = std::string("fwd_small_") + (UNIFORM ? "u" : "g") + std::to_string(NP) + (sizeof(T) == 4 ? "_f32" : "_f64");
    fsw_prof_begin(label.c_str(), st);
    fsw_fwd_small_kernel<T, NP, UNIFORM><<<(unsigned)blocks, 128, 0, st>>>(a, lo, hi, G, nchunks, out, ld_out, out_col0, bias);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_fwd_small_kernel");
    return FSW_OK;
}

template <typename T, bool UNIFORM>
int dispatch_fwd_small(const SegArgs<T>& a, int np, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0,
                       const T* bias, cudaStream_t st) {
    switch (np) {
        case 4: return launch_fwd_small<T, 4, UNIFORM>(a, lo, hi, out, ld_out, out_col0, bias, st);
        case 8: return launch_fwd_small<T, 8, UNIFORM>(a, lo, hi, out, ld_out, out_col0, bias, st);
        case 16: return launch_fwd_small<T, 16, UNIFORM>(a, lo, hi, out, ld_out, out_col0, bias, st);
        case 32: return launch_fwd_small<T, 32, UNIFORM>(a, lo, hi, out, ld_out, out_col0, bias, st);
        case 64:
            if constexpr (sizeof(T) == 4) return launch_fwd_small<T, 64, UNIFORM>(a, lo, hi, out, ld_out, out_col0, bias, st);
            break;
    }
    return fsw_fail(FSW_ERR_INVALID, "bad small class %d", np);
}

template <typename T, int NP, bool UNIFORM, bool NEED_DXI>
int launch_bwd_small(const SegArgs<T>& a, int lo, int hi, const T* g, int64_t ld_g, int64_t g_col0, T* dXp, T* dEp,
                     double* dfreqs, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int G = pick_G(hi - lo, nchunks);
    const int64_t warps = fsw_cdiv(hi - lo, G) * nchunks;
    const int64_t blocks = fsw_cdiv(warps, 4);
    const size_t smem = (size_t)4 * 3 * NP * 32 * sizeof(T);
    auto kern = fsw_bwd_small_kernel<T, NP, UNIFORM, NEED_DXI>;
    if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static const std::string label = std::string("bwd_small_") + (UNIFORM ? "u" : "g") + std::to_string(NP) + (sizeof(T) == 4 ? "_f32" : "_f64");
    fsw_prof_begin(label.c_str(), st);
    kern<<<(unsigned)blocks, 128, smem, st>>>(a, lo, hi, G, nchunks, g, ld_g, g_col0, dXp, dEp, dfreqs);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_bwd_small_kernel");
    return FSW_OK;
}

template <typename T, bool UNIFORM, bool NEED_DXI>
int dispatch_bwd_small(const SegArgs<T>& a, int np, int lo, int hi, const T* g, int64_t ld_g, int64_t g_col0, T* dXp,
                       T* dEp, double* dfreqs, cudaStream_t st) {
    switch (np) {
        case 4: return launch_bwd_small<T, 4, UNIFORM, NEED_DXI>(a, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        case 8: return launch_bwd_small<T, 8, UNIFORM, NEED_DXI>(a, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        case 16: return launch_bwd_small<T, 16, UNIFORM, NEED_DXI>(a, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        case 32: return launch_bwd_small<T, 32, UNIFORM, NEED_DXI>(a, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
        case 64:
            if constexpr (sizeof(T) == 4) return launch_bwd_small<T, 64, UNIFORM, NEED_DXI>(a, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
            break;
    }
    return fsw_fail(FSW_ERR_INVALID, "bad small class %d", np);
}

// number of resident CTAs we use for the persistent generic kernels
const int kGenericGrid = 148 * 2;

template <typename T>
size_t generic_tile_bytes(int cap_rows, bool backward, bool uniform) {
    const size_t per = backward ? (sizeof(T) + sizeof(int)) : (uniform ? sizeof(T) : 2 * sizeof(T));
    return (size_t)cap_rows * 32 * per;
}

template <typename T>
SegArgs<T> make_args(const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr, int64_t n_fixed,
                     const int32_t* col, const void* W, const double* mass, const int32_t* info, const int32_t* order,
                     const void* freqs, int64_t K, double thresh) {
    SegArgs<T> a;
    a.Xp = (const T*)Xp;
    a.Ep = (const T*)Ep;
    a.rowptr = rowptr;
    a.col = col;
    a.W = (const T*)W;
    a.mass = mass;
    a.info = info;
    a.order = order;
    a.freqs = (const T*)freqs;
    a.ldp = ldp;
    a.n_fixed = n_fixed;
    a.K = (int)K;
    a.thresh = thresh;
    return a;
}

// size classes of the uniform-weight small path (exact merge-exchange networks of these sizes)
const ClassRange kSmallU[] = {{0, 4, 4}, {5, 8, 8}, {9, 12, 12}, {13, 16, 16}, {17, 24, 24}, {25, 32, 32}, {33, 48, 48}, {49, 64, 64}};

template <typename T>
int embed_forward_t(const SegArgs<T>& a, const int32_t* bo, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                    int64_t max_n_eff, void* scratch, size_t scratch_bytes, unsigned short* ranks, int64_t ldr,
                    T* dxi_out, int64_t ld_dxi, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int msn = max_small_np<T>();
    // fp32: slice-major coefficient tables for n <= FSW_FWD_TAB_NMAX at the front of the scratch, when a size class
    // served by the packed-key kernels is present
    const float* gtab_c = nullptr;
    const float* gtab_t = nullptr;
    int tab_n0 = 0, tab_ld4 = 0, tab_nmax = 0;   // n covered by the tables: tab_n0 .. tab_nmax
    if constexpr (sizeof(T) == 4) {
        const bool dense_single = a.rowptr == nullptr && a.n_fixed > FSW_FWD_TAB_NMAX && a.n_fixed <= 1024 && bo[0] < bo[FSW_PLAN_BUCKETS_PER_KIND];
        const int nmax_present = fsw_fwd_tables_nmax(bo);
        const size_t tb = dense_single ? fsw_fwd_tables_bytes_single(a.K, a.n_fixed) : fsw_fwd_tables_bytes(a.K, nmax_present);
        if (scratch_bytes >= tb && (dense_single || nmax_present > 0)) {
            float* tc = (float*)scratch;
            float* tt = tc + tb / (2 * sizeof(float));
            int rc;
            if (dense_single) {  // a dense batch has one segment size: a table for that n alone (a.K * n * 4 floats)
                tab_n0 = tab_nmax = (int)a.n_fixed;
                tab_ld4 = (int)((a.n_fixed + 3) / 4);
                rc = fsw_build_fwd_tables(a.freqs, a.K, tab_n0, tab_n0, tab_ld4, tc, tt, st);
            } else {
                tab_nmax = nmax_present;
                tab_ld4 = (nmax_present + 3) / 4;
                rc = fsw_build_fwd_tables(a.freqs, a.K, 1, nmax_present, tab_ld4, tc + (int64_t)tab_ld4 * a.K * 4,
                                          tt + (int64_t)tab_ld4 * a.K * 4, st);
            }
            if (rc) return rc;
            gtab_c = tc;
            gtab_t = tt;
            scratch = (unsigned char*)scratch + tb;
            scratch_bytes -= tb;
        }
    }
    for (int kind = 0; kind < 2; ++kind) {
        const int base = kind * FSW_PLAN_BUCKETS_PER_KIND;
        if (kind == 0) {
            for (const ClassRange& c : kSmallU) {
                if (c.np > msn) continue;
                const int lo = bo[base + c.lo], hi = bo[base + c.hi + 1];
                if (hi <= lo) continue;
                if constexpr (sizeof(T) == 4) {
                    if (gtab_c != nullptr && c.np > 32 && tab_n0 == 0) {  // 33..48, 49..64: packed keys, 4 cooperating lanes
                        const int np = c.np;
                        int rc = fsw_packed_forward_u(a, np, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
                        if (rc) return rc;
                        continue;
                    }
                }
                if (a.projX != nullptr) return fsw_fail(FSW_ERR_UNSUPPORTED, "point-cloud mode: only the packed-key classes (33..1024 points) form keys on the fly");
                int rc = fsw_small_forward_u<T>(a, c.np, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, st);
                if (rc) return rc;
            }
        } else {
            for (const ClassRange& c : kSmall) {
                if (c.np > msn) continue;
                const int lo = bo[base + c.lo], hi = bo[base + c.hi + 1];
                if (hi <= lo) continue;
                if (a.projX != nullptr) return fsw_fail(FSW_ERR_UNSUPPORTED, "point-cloud mode needs uniform weights and total mass >= the pad threshold");
                int rc = dispatch_fwd_small<T, false>(a, c.np, lo, hi, out, ld_out, out_col0, bias, st);
                if (rc) return rc;
            }
        }
        SizeRange rr[12];
        const int nr = generic_ranges(msn, max_n_eff, rr);
        for (int ri = 0; ri < nr; ++ri) {
            const int lo = bo[base + rr[ri].lo], hi = bo[base + rr[ri].hi + 1];
            const int cap = rr[ri].cap;
            if (hi <= lo) continue;
            if constexpr (sizeof(T) == 4) {
                // 65..512 elements (and dense batches of up to 1024): packed keys, cooperating lanes.  The graph tables
                // (tab_n0 == 0) reach the largest size <= 512 that occurs; a dense single-size table serves its own class.
                if (kind == 0 && gtab_c != nullptr && cap <= 1024 &&
                    (tab_n0 == 0 ? cap <= (tab_nmax > FSW_FWD_TAB_NMAX ? FSW_FWD_TAB_GRAPH_NMAX : FSW_FWD_TAB_NMAX) : (a.n_fixed > cap / 2 && a.n_fixed <= cap))) {
                    int rc = FSW_OK;
                    if (tab_n0 == 0 && cap <= FSW_FWD_TAB_NMAX) {
                        // two classes per power of two: 3/4 cap slots for the lower part of the range (less padding)
                        const int mid = bo[base + 3 * cap / 4 + 1];
                        if (mid > lo) rc = fsw_packed_forward_u(a, 3 * cap / 4, lo, mid, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
                        if (rc == FSW_OK && hi > mid) rc = fsw_packed_forward_u(a, cap, mid, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
                    } else {
                        rc = fsw_packed_forward_u(a, cap, lo, hi, out, ld_out, out_col0, bias, ranks, ldr, dxi_out, ld_dxi, gtab_c, gtab_t, tab_n0, tab_ld4, st);
                    }
                    if (rc) return rc;
                    continue;
                }
                if (a.projX != nullptr) return fsw_fail(FSW_ERR_UNSUPPORTED, "point-cloud mode: segment class %d is not served by the packed-key kernels", cap);
                if (kind == 0 && cap >= 128) {  // medium / large path: uniform weights, more than 128 elements
                    int rc = fsw_medium_forward_f32(a, lo, hi, cap, out, ld_out, out_col0, bias, scratch, scratch_bytes, ranks, ldr, dxi_out, ld_dxi, nullptr, nullptr, st);
                    if (rc) return rc;
                    continue;
                }
            }
            const int64_t ntiles = (int64_t)(hi - lo) * nchunks;
            const size_t tb = generic_tile_bytes<T>(cap, false, kind == 0);
            unsigned grid = (unsigned)(ntiles < kGenericGrid ? ntiles : kGenericGrid);
            unsigned char* gs = nullptr;
            size_t smem = tb;
            if (tb > (size_t)kSmemBudget) {
                if ((size_t)grid * tb > scratch_bytes) {
                    grid = (unsigned)(scratch_bytes / tb);
                    if (grid == 0) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small: need >= %zu bytes", tb);
                }
                gs = (unsigned char*)scratch;
                smem = 0;
            }
            const std::string label = std::string("fwd_generic_") + (kind == 0 ? "u" : "g") + std::to_string(cap) + (sizeof(T) == 4 ? "_f32" : "_f64");
            fsw_prof_begin(label.c_str(), st);
            if (kind == 0) {
                auto kern = fsw_fwd_generic_kernel<T, 0>;
                if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, 256, smem, st>>>(a, lo, nchunks, ntiles, out, ld_out, out_col0, bias, cap, gs);
            } else {
                auto kern = fsw_fwd_generic_kernel<T, 1>;
                if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                kern<<<grid, 256, smem, st>>>(a, lo, nchunks, ntiles, out, ld_out, out_col0, bias, cap, gs);
            }
            fsw_prof_end(st);
            FSW_CHECK_LAUNCH("fsw_fwd_generic_kernel");
        }
    }
    return FSW_OK;
}

// Transposed segment structure for the source-major backward (optional; graphs only)
struct TransposeArgs {
    const int32_t* tptr;
    const int32_t* tseg;
    const int32_t* tslot;
    const int32_t* tn;
    int64_t nrows;
    int64_t S;
};

// `dxi_fwd`: the forward already produced d out / d xi for the uniform-weight classes it ran with rank
// recording (all small classes; fp32 medium classes up to 32768 elements), so their backward runs without the
// frequency gradient; every other class accumulates it into `dfreqs` (when requested).
template <typename T>
int embed_backward_t(const SegArgs<T>& a, const int32_t* bo, const T* g, int64_t ld_g, int64_t g_col0, T* dXp, T* dEp,
                     double* dfreqs, bool dxi_fwd, int64_t max_n_eff, void* scratch, size_t scratch_bytes,
                     const unsigned short* ranks, int64_t ldr, const TransposeArgs& tr, cudaStream_t st) {
    const int nchunks = (a.K + 31) / 32;
    const int msn = max_small_np<T>();
    const bool have_ranks = ranks != nullptr;
    double* dfreqs_cov = (have_ranks && dxi_fwd) ? nullptr : dfreqs;  // for classes covered by the forward d/dxi
    if constexpr (sizeof(T) == 4) {
        // dense batch of unit-weight multisets (every segment uniform, n elements, ranks recorded for all of them):
        // one streaming kernel does the whole backward
        if (have_ranks && dfreqs_cov == nullptr && a.rowptr == nullptr && a.col == nullptr && a.W == nullptr && a.n_fixed >= 1 &&
            a.n_fixed <= 32768 && (double)a.n_fixed >= a.thresh && bo[FSW_PLAN_BUCKETS_PER_KIND] == bo[2 * FSW_PLAN_BUCKETS_PER_KIND]) {
            const int64_t S = bo[2 * FSW_PLAN_BUCKETS_PER_KIND];
            return fsw_rank_backward_dense(a, S, (int)a.n_fixed, ranks, ldr, g, ld_g, g_col0, dXp, dEp, st);
        }
    }
    bool source_major = false;   // uniform segments of up to tnmax elements done by the source-major rank backward
    const int tnmax = FSW_RANKT_ELIGIBLE(max_n_eff);
    for (int kind = 0; kind < 2; ++kind) {
        const int base = kind * FSW_PLAN_BUCKETS_PER_KIND;
        if (kind == 0 && have_ranks) {
            // rank-based backward: no sorting
            if constexpr (sizeof(T) == 4) {
                const size_t tb = fsw_rank_tables_bytes(a.ldp);
                if (scratch_bytes < tb) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small for the rank tables");
                // the pre-scaled gradient GA [S, ldp] sits at the END of the scratch
                const size_t ga_bytes = (size_t)tr.S * a.ldp * sizeof(float);
                source_major = a.col != nullptr && tr.tptr != nullptr && dfreqs_cov == nullptr && scratch_bytes >= tb + ga_bytes + 256;
                if (source_major) {
                    // graphs: every uniform segment up to tnmax elements, one plain store per row of dXp
                    // (must precede every kernel that adds with atomics)
                    if (bo[base + FSW_PLAN_BUCKETS_PER_KIND] > bo[base + 0]) {
                        const size_t ga_off = (scratch_bytes - ga_bytes) & ~(size_t)255;  // keep 128-bit accesses aligned
                        float* ga_buf = (float*)((unsigned char*)scratch + ga_off);
                        scratch_bytes = ga_off;
                        int rc = fsw_rank_backward_T(a, tr.S, tr.nrows, tnmax, tr.tptr, tr.tseg, tr.tslot, tr.tn, ranks, ldr, g, ld_g, g_col0, dXp, dEp, scratch, ga_buf, st);
                        if (rc) return rc;
                    } else {
                        // contract: when the transposition is passed the caller need not zero dXp - nothing will be written
                        // by the source-major kernel here, so clear it before the atomics
                        FSW_CUDA(cudaMemsetAsync(dXp, 0, (size_t)tr.nrows * a.ldp * sizeof(T), st));
                    }
                } else {
                    if (tr.tptr != nullptr) FSW_CUDA(cudaMemsetAsync(dXp, 0, (size_t)tr.nrows * a.ldp * sizeof(T), st));
                    const int lo0 = bo[base + 0], hi0 = bo[base + 128 + 1];
                    if (hi0 > lo0) {
                        int rc = fsw_rank_backward_g128(a, lo0, hi0, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs_cov, scratch, st);
                        if (rc) return rc;
                    }
                }
                scratch = (unsigned char*)scratch + tb;
                scratch_bytes -= tb;
                struct { int lo, hi, cap; } rk[2] = {{129, 256, 256}, {257, 512, 512}};
                for (int i = 0; i < 2 && !source_major; ++i) {
                    const int lo = bo[base + rk[i].lo], hi = bo[base + rk[i].hi + 1];
                    if (hi <= lo) continue;
                    int rc = fsw_rank_backward_u<T>(a, lo, hi, rk[i].cap, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs_cov, st);
                    if (rc) return rc;
                }
            } else {
                const int lo = bo[base + 0], hi = bo[base + msn + 1];
                if (hi > lo) {
                    int rc = fsw_rank_backward_u<T>(a, lo, hi, msn, ranks, ldr, g, ld_g, g_col0, dXp, dEp, dfreqs_cov, st);
                    if (rc) return rc;
                }
            }
        }
        for (const ClassRange& c : kSmall) {
            if (kind == 0 && have_ranks) break;
            if (c.np > msn) continue;
            const int lo = bo[base + c.lo], hi = bo[base + c.hi + 1];
            if (hi <= lo) continue;
            int rc;
            if (dfreqs)
                rc = kind == 0 ? dispatch_bwd_small<T, true, true>(a, c.np, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                               : dispatch_bwd_small<T, false, true>(a, c.np, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
            else
                rc = kind == 0 ? dispatch_bwd_small<T, true, false>(a, c.np, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st)
                               : dispatch_bwd_small<T, false, false>(a, c.np, lo, hi, g, ld_g, g_col0, dXp, dEp, dfreqs, st);
            if (rc) return rc;
        }
        SizeRange rr[12];
        const int nr = generic_ranges(msn, max_n_eff, rr);
        for (int ri = 0; ri < nr; ++ri) {
            const int lo = bo[base + rr[ri].lo], hi = bo[base + rr[ri].hi + 1];
            const int cap = rr[ri].cap;
            if (hi <= lo) continue;
            if constexpr (sizeof(T) == 4) {
                if (kind == 0 && cap >= 128 && cap <= 512 && have_ranks) continue;  // done by the rank kernels
                if (kind == 0 && source_major && cap <= tnmax) continue;  // source-major rank backward (hubs beyond: re-sorted below)
                if (kind == 0 && cap >= 128) {
                    // re-sorting backward; the forward produced d/dxi for these when it recorded ranks (cap <= 32768)
                    double* df = (have_ranks && cap <= 32768) ? dfreqs_cov : dfreqs;
                    int rc = fsw_medium_backward_f32(a, lo, hi, cap, g, ld_g, g_col0, dXp, dEp, df, scratch, scratch_bytes, st);
                    if (rc) return rc;
                    continue;
                }
            }
            const int64_t ntiles = (int64_t)(hi - lo) * nchunks;
            const size_t tb = generic_tile_bytes<T>(cap, true, kind == 0);
            unsigned grid = (unsigned)(ntiles < kGenericGrid ? ntiles : kGenericGrid);
            unsigned char* gs = nullptr;
            size_t smem = tb;
            if (tb > (size_t)kSmemBudget) {
                if ((size_t)grid * tb > scratch_bytes) {
                    grid = (unsigned)(scratch_bytes / tb);
                    if (grid == 0) return fsw_fail(FSW_ERR_WORKSPACE, "embed scratch too small: need >= %zu bytes", tb);
                }
                gs = (unsigned char*)scratch;
                smem = 0;
            }
            const std::string label = std::string("bwd_generic_") + (kind == 0 ? "u" : "g") + std::to_string(cap) + (sizeof(T) == 4 ? "_f32" : "_f64");
            fsw_prof_begin(label.c_str(), st);
#define FSW_GEN_BWD(UNI, DXI)                                                                                      \
    do {                                                                                                         \
        auto kern = fsw_bwd_generic_kernel<T, UNI, DXI>;                                                         \
        if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, 256, smem, st>>>(a, lo, nchunks, ntiles, g, ld_g, g_col0, dXp, dEp, dfreqs, cap, gs);        \
    } while (0)
            if (kind == 0) {
                if (dfreqs) FSW_GEN_BWD(true, true); else FSW_GEN_BWD(true, false);
            } else {
                if (dfreqs) FSW_GEN_BWD(false, true); else FSW_GEN_BWD(false, false);
            }
#undef FSW_GEN_BWD
            fsw_prof_end(st);
            FSW_CHECK_LAUNCH("fsw_bwd_generic_kernel");
        }
    }
    return FSW_OK;
}

}  // namespace

extern "C" size_t fsw_embed_scratch_bytes(int dtype, const int32_t* bo, int64_t K, int64_t max_n_eff, int backward) {
    // worst case over the non-empty generic ranges of grid * tile_bytes, only for tiles beyond the smem budget
    // (+ the coefficient tables of the rank-based backward, which sit in front of the tiles)
    size_t need = 0;
    const size_t es = dtype == FSW_F64 ? 8 : 4;
    const int msn = dtype == FSW_F64 ? 32 : 64;
    const int nchunks = (int)((K + 31) / 32);
    SizeRange rr[12];
    const int nr = generic_ranges(msn, max_n_eff, rr);
    for (int kind = 0; kind < 2; ++kind) {
        const int base = kind * FSW_PLAN_BUCKETS_PER_KIND;
        for (int ri = 0; ri < nr; ++ri) {
            const int cnt = bo[base + rr[ri].hi + 1] - bo[base + rr[ri].lo];
            if (cnt <= 0) continue;
            int64_t ntiles = (int64_t)cnt * nchunks;
            if (dtype == FSW_F32 && kind == 0 && rr[ri].cap >= 128) {  // medium / large path
                const size_t tb = fsw_medium_tile_bytes(rr[ri].cap, backward ? 2 : 1);
                size_t grid = (size_t)(ntiles < fsw_medium_grid() ? ntiles : fsw_medium_grid());
                size_t want = grid * tb;
                const size_t cap_bytes = (size_t)2 << 30;  // never ask for more than 2 GiB: fewer resident CTAs instead
                if (want > cap_bytes) want = (cap_bytes / tb ? cap_bytes / tb : 1) * tb;
                want += 1024;   // the grid-wide (team) mode keeps its global accumulators behind the tile
                if (want > need) need = want;
                continue;
            }
            const size_t per = backward ? (es + 4) : (kind == 0 ? es : 2 * es);
            const size_t tb = (size_t)rr[ri].cap * 32 * per;
            if (tb <= (size_t)kSmemBudget) continue;
            size_t grid = (size_t)(ntiles < kGenericGrid ? ntiles : kGenericGrid);
            if (grid * tb > need) need = grid * tb;
        }
    }
    if (dtype == FSW_F32) {
        size_t tabs = fsw_rank_tables_bytes((K + 7) / 8 * 8);
        if (!backward) {
            tabs = fsw_fwd_tables_bytes(K, fsw_fwd_tables_nmax(bo));
            if (max_n_eff > FSW_FWD_TAB_NMAX && max_n_eff <= 1024 && fsw_fwd_tables_bytes_single(K, max_n_eff) > tabs)
                tabs = fsw_fwd_tables_bytes_single(K, max_n_eff);
        }
        need += tabs;
    }
    return need;
}

extern "C" size_t fsw_embed_backward_extra_bytes(int dtype, int64_t S, int64_t K) {
    // pre-scaled gradient buffer of the source-major backward, appended to the scratch by the caller
    if (dtype != FSW_F32) return 0;
    return (size_t)S * ((K + 7) / 8 * 8) * sizeof(float) + 256;
}

extern "C" int fsw_embed_forward(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr,
                                 int64_t n_fixed, const int32_t* col, const void* W, const double* mass,
                                 const int32_t* info, const int32_t* order, const int32_t* bucket_offsets_host,
                                 int64_t S, int64_t K, const void* freqs, double thresh, void* out, int64_t ld_out,
                                 int64_t out_col0, const void* bias, int64_t max_n_eff, void* scratch,
                                 size_t scratch_bytes, void* ranks_out, int64_t ldr, void* dxi_out, int64_t ld_dxi,
                                 void* stream) {
    if (S == 0 || K == 0) return FSW_OK;
    if (!Xp || !mass || !info || !bucket_offsets_host || !freqs || !out)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_forward: null argument");
    if (!rowptr && n_fixed <= 0) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_forward: rowptr == NULL needs n_fixed > 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FSW_F32) {
        auto a = make_args<float>(Xp, ldp, Ep, rowptr, n_fixed, col, W, mass, info, order, freqs, K, thresh);
        return embed_forward_t<float>(a, bucket_offsets_host, (float*)out, ld_out, out_col0, (const float*)bias, max_n_eff, scratch, scratch_bytes, (unsigned short*)ranks_out, ldr, (float*)dxi_out, ld_dxi, st);
    } else if (dtype == FSW_F64) {
        auto a = make_args<double>(Xp, ldp, Ep, rowptr, n_fixed, col, W, mass, info, order, freqs, K, thresh);
        return embed_forward_t<double>(a, bucket_offsets_host, (double*)out, ld_out, out_col0, (const double*)bias, max_n_eff, scratch, scratch_bytes, (unsigned short*)ranks_out, ldr, (double*)dxi_out, ld_dxi, st);
    }
    return fsw_fail(FSW_ERR_INVALID, "fsw_embed_forward: dtype %d", dtype);
}

// Point-cloud mode of the forward (include/fsw_embedding.h section 5b): dense batch of S multisets of n points each, unit
// weights, d <= 4: keys formed on the fly from the points, ranks recorded slice-major.
extern "C" int fsw_embed_forward_cloud(int dtype, const void* X, int64_t d, const void* theta, int64_t ldt, int64_t n,
                                       const double* mass, const int32_t* info, const int32_t* order,
                                       const int32_t* bucket_offsets_host, int64_t S, int64_t K, const void* freqs, double thresh,
                                       void* out, int64_t ld_out, int64_t out_col0, const void* bias, void* scratch,
                                       size_t scratch_bytes, void* ranksT_out, void* dxi_out, int64_t ld_dxi, void* stream) {
    if (S == 0 || K == 0) return FSW_OK;
    if (dtype != FSW_F32) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_embed_forward_cloud: fp32 only");
    if (!X || !theta || !mass || !info || !bucket_offsets_host || !freqs || !out)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_forward_cloud: null argument");
    if (d < 1 || d > 4 || n < 33 || n > 1024)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_forward_cloud: needs 1 <= d <= 4 and 33 <= n <= 1024 (got d=%lld n=%lld)", (long long)d, (long long)n);
    auto a = make_args<float>(X, d, nullptr, nullptr, n, nullptr, nullptr, mass, info, order, freqs, K, thresh);
    a.projX = (const float*)X;
    a.projTheta = (const float*)theta;
    a.proj_d = (int)d;
    a.proj_ldt = ldt;
    a.rank_transposed = 1;
    return embed_forward_t<float>(a, bucket_offsets_host, (float*)out, ld_out, out_col0, (const float*)bias, n, scratch, scratch_bytes,
                                  (unsigned short*)ranksT_out, 0, (float*)dxi_out, ld_dxi, (cudaStream_t)stream);
}

extern "C" int fsw_embed_backward(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr,
                                  int64_t n_fixed, const int32_t* col, const void* W, const double* mass,
                                  const int32_t* info, const int32_t* order, const int32_t* bucket_offsets_host,
                                  int64_t S, int64_t K, const void* freqs, double thresh, const void* g, int64_t ld_g,
                                  int64_t g_col0, void* dXp, void* dEp, double* dfreqs_acc, void* dW,
                                  int64_t max_n_eff, void* scratch, size_t scratch_bytes, const void* ranks, int64_t ldr,
                                  int dxi_from_forward, const int32_t* tptr, const int32_t* tseg, const int32_t* tslot,
                                  const int32_t* tn, int64_t nrows, void* stream) {
    if (S == 0 || K == 0) {
        // with the transposition passed, dXp [nrows, ldp] is an output that need not be initialised: an empty shard yields zeros
        if (tptr != nullptr && dXp != nullptr && nrows > 0 && ldp > 0) {
            const size_t es = dtype == FSW_F64 ? 8 : 4;
            if (cudaMemsetAsync(dXp, 0, (size_t)nrows * ldp * es, (cudaStream_t)stream) != cudaSuccess)
                return fsw_fail(FSW_ERR_CUDA, "fsw_embed_backward: memset failed");
        }
        return FSW_OK;
    }
    if (dW != nullptr)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward: dW must be NULL (use fsw_embed_backward_weights)");
    if (!Xp || !mass || !info || !bucket_offsets_host || !freqs || !g || !dXp)
        return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward: null argument");
    if (!rowptr && n_fixed <= 0) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward: rowptr == NULL needs n_fixed > 0");
    cudaStream_t st = (cudaStream_t)stream;
    TransposeArgs tr{tptr, tseg, tslot, tn, nrows, S};
    if (dtype == FSW_F32) {
        auto a = make_args<float>(Xp, ldp, Ep, rowptr, n_fixed, col, W, mass, info, order, freqs, K, thresh);
        return embed_backward_t<float>(a, bucket_offsets_host, (const float*)g, ld_g, g_col0, (float*)dXp, (float*)dEp, dfreqs_acc,
                                       dxi_from_forward != 0, max_n_eff, scratch, scratch_bytes, (const unsigned short*)ranks, ldr, tr, st);
    } else if (dtype == FSW_F64) {
        auto a = make_args<double>(Xp, ldp, Ep, rowptr, n_fixed, col, W, mass, info, order, freqs, K, thresh);
        return embed_backward_t<double>(a, bucket_offsets_host, (const double*)g, ld_g, g_col0, (double*)dXp, (double*)dEp, dfreqs_acc,
                                        dxi_from_forward != 0, max_n_eff, scratch, scratch_bytes, (const unsigned short*)ranks, ldr, tr, st);
    }
    return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward: dtype %d", dtype);
}
