// K1: dense contractions at full input precision (fp32 FMA / fp64) for
//   op 0 (NT)  Xp      = X   . theta^T     slice projection, fsw_embedding.py:911/:913/:936
//   op 1 (NN)  dX      = dXp . theta       autograd of :911 w.r.t. X
//   op 2 (TN)  dtheta += dXp^T . X         autograd of :911 w.r.t. projVecs (split over the long N axis)
//
// The reference runs these as cuBLAS SGEMM with TF32 off, so the contraction must be true fp32:
// register-tiled SIMT kernels: a full-width strip kernel for the shapes of the path (small dimension <= 200, fp32)
// and a generic 128x64x16 kernel for everything else (fp64, odd strides).
// A tcgen05 3xTF32 variant for large d_in is future work.
#include <stdlib.h>

#include "fsw_common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

template <typename T>
__global__ void __launch_bounds__(256) fsw_gemm_kernel(int64_t M, int64_t N, int64_t Kd, const T* __restrict__ A, int64_t sa_m,
                                                       int64_t sa_k, const T* __restrict__ B, int64_t sb_n, int64_t sb_k,
                                                       T* __restrict__ C, int64_t ldc, int accumulate, int64_t k_per_split,
                                                       int use_atomics) {
    __shared__ T As[BK][BM + 4];
    __shared__ T Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15;   // column group
    const int ty = tid >> 4;   // row group
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < Kd) ? kbeg + k_per_split : Kd;

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = (T)0;

    const bool a_kfast = (sa_k == 1);
    const bool b_kfast = (sb_k == 1);

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- global -> shared (zero filled at the edges) ----
#pragma unroll
        for (int it = 0; it < (BM * BK) / 256; ++it) {
            const int idx = tid + it * 256;
            int m, k;
            if (a_kfast) {
                k = idx % BK;
                m = idx / BK;
            } else {
                m = idx % BM;
                k = idx / BM;
            }
            const int64_t gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < kend) ? __ldg(A + gm * sa_m + gk * sa_k) : (T)0;
        }
#pragma unroll
        for (int it = 0; it < (BN * BK) / 256; ++it) {
            const int idx = tid + it * 256;
            int n, k;
            if (b_kfast) {
                k = idx % BK;
                n = idx / BK;
            } else {
                n = idx % BN;
                k = idx / BN;
            }
            const int64_t gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < kend) ? __ldg(B + gn * sb_n + gk * sb_k) : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            T* c = C + gm * ldc + gn;
            if (use_atomics)
                atomicAdd(c, acc[i][j]);
            else if (accumulate)
                *c += acc[i][j];
            else
                *c = acc[i][j];
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// fp32 strip kernel: the CTA tile spans the WHOLE small dimension N (N <= 8 CG, CG = 13 or 25 column groups), so the
// long operand is read from HBM exactly once and nothing is wasted on padding N to a power of two (N = 199 slices or
// 100 features would lose 22 % in 128-wide tiles).  RG x CG threads own 8 x 8 outputs each (columns split 4 + 4 so that
// the 128-bit shared loads of a quarter warp touch 32 distinct banks), k-major shared tiles, register-prefetch double
// buffering: the loads of step i+1 are in flight while step i is multiplied.
//   AMODE / BMODE 0: element (r, k) at r * ld + k (rows of length Kd; read as float4 along k, transposed on the way in)
//                 1: element (r, k) at k * ld + r (float4 along r)
// Requirements (checked by the dispatcher): ld % 4 == 0, 16-byte aligned bases (split boundaries are multiples of 16).
// ---------------------------------------------------------------------------------------------------
constexpr int SBK = 16;

template <int RG, int CG>
struct StripCfg {
    static constexpr int BMs = 8 * RG, BNs = 8 * CG;
    static constexpr int NTHR = ((RG * CG + 31) / 32) * 32;
    static constexpr int LDA = BMs + 4, LDB = BNs + 4;
    static constexpr int UA = BMs * SBK / 4, UB = BNs * SBK / 4;  // float4 units per k-step
    static constexpr int PA = (UA + NTHR - 1) / NTHR, PB = (UB + NTHR - 1) / NTHR;
};

// one float4 unit of a [ROWS x SBK] tile: mode 0 -> (r, 4 consecutive k), mode 1 -> (4 consecutive r, k)
template <int MODE, int ROWS>
__device__ __forceinline__ float4 fsw_strip_load(const float* __restrict__ P, int64_t ld, int64_t r0, int64_t R, int64_t k0,
                                                 int64_t kend, int u) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 0) {
        const int r = u % ROWS, kq = u / ROWS;
        const int64_t gr = r0 + r, gk = k0 + kq * 4;
        if (gr < R && gk < kend) {
            const float* p = P + gr * ld + gk;
            if (gk + 3 < kend) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {  // tail of a row whose length is not a multiple of 4
                v.x = __ldg(p);
                if (gk + 1 < kend) v.y = __ldg(p + 1);
                if (gk + 2 < kend) v.z = __ldg(p + 2);
            }
        }
    } else {
        const int rq = u % (ROWS / 4), k = u / (ROWS / 4);
        const int64_t gr = r0 + rq * 4, gk = k0 + k;
        if (gk < kend && gr < R) {
            const float* p = P + gk * ld + gr;
            if (gr + 3 < R) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                v.x = __ldg(p);
                if (gr + 1 < R) v.y = __ldg(p + 1);
                if (gr + 2 < R) v.z = __ldg(p + 2);
            }
        }
    }
    return v;
}

template <int MODE, int ROWS, int LD>
__device__ __forceinline__ void fsw_strip_store(float* __restrict__ S, int u, float4 v) {
    if (MODE == 0) {
        const int r = u % ROWS, kq = u / ROWS;
        S[(kq * 4 + 0) * LD + r] = v.x;
        S[(kq * 4 + 1) * LD + r] = v.y;
        S[(kq * 4 + 2) * LD + r] = v.z;
        S[(kq * 4 + 3) * LD + r] = v.w;
    } else {
        const int rq = u % (ROWS / 4), k = u / (ROWS / 4);
        *reinterpret_cast<float4*>(S + k * LD + rq * 4) = v;
    }
}

template <int RG, int CG, int AMODE, int BMODE>
__global__ void __launch_bounds__(StripCfg<RG, CG>::NTHR, StripCfg<RG, CG>::NTHR <= 256 ? 2 : 1) fsw_gemm_strip_kernel(int64_t M, int64_t N, int64_t Kd, const float* __restrict__ A,
                                                                                int64_t lda, const float* __restrict__ B, int64_t ldb,
                                                                                float* __restrict__ C, int64_t ldc, int accumulate,
                                                                                int64_t k_per_split, int use_atomics) {
    using Cfg = StripCfg<RG, CG>;
    constexpr int BMs = Cfg::BMs, BNs = Cfg::BNs, NTHR = Cfg::NTHR, LDA = Cfg::LDA, LDB = Cfg::LDB;
    __shared__ __align__(16) float As[2][SBK * LDA];
    __shared__ __align__(16) float Bs[2][SBK * LDB];
    const int tid = threadIdx.x;
    const bool worker = tid < RG * CG;
    const int rg = worker ? tid / CG : 0, cg = worker ? tid % CG : 0;
    const int64_t m0 = (int64_t)blockIdx.x * BMs;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = (kbeg + k_per_split < Kd) ? kbeg + k_per_split : Kd;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 pa[Cfg::PA], pb[Cfg::PB];
    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int p = 0; p < Cfg::PA; ++p) {
            const int u = tid + p * NTHR;
            pa[p] = (u < Cfg::UA) ? fsw_strip_load<AMODE, BMs>(A, lda, m0, M, k0, kend, u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < Cfg::PB; ++p) {
            const int u = tid + p * NTHR;
            pb[p] = (u < Cfg::UB) ? fsw_strip_load<BMODE, BNs>(B, ldb, 0, N, k0, kend, u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int p = 0; p < Cfg::PA; ++p) {
            const int u = tid + p * NTHR;
            if (u < Cfg::UA) fsw_strip_store<AMODE, BMs, LDA>(As[buf], u, pa[p]);
        }
#pragma unroll
        for (int p = 0; p < Cfg::PB; ++p) {
            const int u = tid + p * NTHR;
            if (u < Cfg::UB) fsw_strip_store<BMODE, BNs, LDB>(Bs[buf], u, pb[p]);
        }
    };

    load_tiles(kbeg);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = kbeg; k0 < kend; k0 += SBK) {
        const bool more = k0 + SBK < kend;
        if (more) load_tiles(k0 + SBK);  // in flight during the multiply below
        if (worker) {
            const float* as = As[buf] + rg * 8;
            const float* bs = Bs[buf] + cg * 4;
#pragma unroll
            for (int k = 0; k < SBK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(as + k * LDA);
                const float4 a1 = *reinterpret_cast<const float4*>(as + k * LDA + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(bs + k * LDB);
                const float4 b1 = *reinterpret_cast<const float4*>(bs + k * LDB + BNs / 2);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (more) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    if (!worker) return;
    // thread's columns: cg*4 .. +3 and BNs/2 + cg*4 .. +3
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + rg * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t gn = (int64_t)h * (BNs / 2) + cg * 4;
            float* c = C + gm * ldc + gn;
            const float v[4] = {acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]};
            if (!use_atomics && !accumulate && gn + 3 < N && ((ldc & 3) == 0)) {
                *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gn + j >= N) continue;
                    if (use_atomics)
                        atomicAdd(c + j, v[j]);
                    else if (accumulate)
                        c[j] += v[j];
                    else
                        c[j] = v[j];
                }
            }
        }
    }
}

template <int RG, int CG, int AMODE, int BMODE>
int launch_strip(int op, int64_t M, int64_t N, int64_t Kd, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                 int64_t ldc, int accumulate, cudaStream_t st) {
    using Cfg = StripCfg<RG, CG>;
    const int64_t mt = fsw_cdiv(M, Cfg::BMs);
    int64_t splits = 1;
    if (op == 2) {  // split the long reduction axis so that the grid fills the machine
        splits = (148 * 2 + mt - 1) / mt;
        const int64_t max_splits = fsw_cdiv(Kd, 8 * SBK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int64_t k_per_split = fsw_cdiv(fsw_cdiv(Kd, splits), SBK) * SBK;
    splits = fsw_cdiv(Kd, k_per_split);
    dim3 grid((unsigned)mt, 1, (unsigned)splits);
    static const char* labels[3] = {"gemm_nt", "gemm_nn", "gemm_tn"};
    fsw_prof_begin(labels[op], st);
    fsw_gemm_strip_kernel<RG, CG, AMODE, BMODE><<<grid, Cfg::NTHR, 0, st>>>(M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, k_per_split,
                                                                         op == 2 ? 1 : 0);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_gemm_strip_kernel");
    return FSW_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// returns -1 when the shape / alignment is not covered (the generic kernel takes over)
int gemm_strip_f32(int op, int64_t M, int64_t N, int64_t Kd, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                   int64_t ldc, int accumulate, cudaStream_t st) {
    if (N > 200 || Kd < 16 || !aligned16(A) || !aligned16(B) || !aligned16(C) || (lda & 3) || (ldb & 3)) return -1;
    const bool wide = N > 104;
    switch (op) {
        case 0:
            return wide ? launch_strip<10, 25, 0, 0>(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st)
                        : launch_strip<16, 13, 0, 0>(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 1:
            return wide ? launch_strip<10, 25, 0, 1>(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st)
                        : launch_strip<16, 13, 0, 1>(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 2:
            if (M > 200 || wide) return -1;
            return launch_strip<25, 13, 1, 1>(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------------
// Backward contractions for small d_in (N <= 8, e.g. 3-d point clouds), fp32: both are ONE streaming pass over the
// [rows, K] gradient of the projections (memory bound), where a tiled GEMM would waste > 90 % of its 64-wide tile.
//   nn_small: C[M, N] (+)= A[M, Kd] . B[Kd, N]      dX = dXp . theta     warp per row, lanes over Kd, shuffle reduction
//   tn_small: C[M, N] += A[Kd, M]^T . B[Kd, N]      dtheta = dXp^T . X   lanes over M, register accumulators over a
//                                                                        run of rows, one atomic per warp and output
// KJ = ceil(width / 32) values per lane (width = Kd resp. M <= 32 KJ).
// ---------------------------------------------------------------------------------------------------
template <int N, int KJ>
__global__ void __launch_bounds__(256) fsw_gemm_nn_small_kernel(int64_t M, int Kd, const float* __restrict__ A, int64_t lda,
                                                                const float* __restrict__ B, int64_t ldb, float* __restrict__ C,
                                                                int64_t ldc, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float b[KJ][N];
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        const int k = lane + 32 * j;
#pragma unroll
        for (int n = 0; n < N; ++n) b[j][n] = (k < Kd) ? __ldg(B + (int64_t)k * ldb + n) : 0.f;
    }
#pragma unroll 2
    for (int64_t r = warp0; r < M; r += nwarps) {
        const float* ar = A + r * lda;
        float v[KJ];
#pragma unroll
        for (int j = 0; j < KJ; ++j) v[j] = (lane + 32 * j < Kd) ? __ldg(ar + lane + 32 * j) : 0.f;
        float s[N];
#pragma unroll
        for (int n = 0; n < N; ++n) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < KJ; ++j) t = fmaf(v[j], b[j][n], t);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            s[n] = t;
        }
        if (lane < N) {
            float out = s[0];
#pragma unroll
            for (int n = 1; n < N; ++n) out = (lane == n) ? s[n] : out;
            float* c = C + r * ldc + lane;
            *c = accumulate ? *c + out : out;
        }
    }
}

template <int N, int KJ>
__global__ void __launch_bounds__(256) fsw_gemm_tn_small_kernel(int64_t Kd, int M, const float* __restrict__ A, int64_t lda,
                                                                const float* __restrict__ B, int64_t ldb, float* __restrict__ C,
                                                                int64_t ldc) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float acc[KJ][N];
#pragma unroll
    for (int j = 0; j < KJ; ++j)
#pragma unroll
        for (int n = 0; n < N; ++n) acc[j][n] = 0.f;
#pragma unroll 4
    for (int64_t r = warp0; r < Kd; r += nwarps) {   // several rows in flight per warp
        const float* ar = A + r * lda;
        float x[N];
#pragma unroll
        for (int n = 0; n < N; ++n) x[n] = __ldg(B + r * ldb + n);   // same address in every lane: one broadcast load
#pragma unroll
        for (int j = 0; j < KJ; ++j) {
            const float v = (lane + 32 * j < M) ? __ldg(ar + lane + 32 * j) : 0.f;
#pragma unroll
            for (int n = 0; n < N; ++n) acc[j][n] = fmaf(v, x[n], acc[j][n]);
        }
    }
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        const int m = lane + 32 * j;
        if (m < M) {
#pragma unroll
            for (int n = 0; n < N; ++n) atomicAdd(C + (int64_t)m * ldc + n, acc[j][n]);
        }
    }
}

template <int N>
int launch_small_n(int op, int64_t M, int64_t Kd, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int accumulate, cudaStream_t st) {
    const int64_t width = (op == 1) ? Kd : M;      // the axis spread over the lanes
    const int KJ = (int)((width + 31) / 32);
    if (KJ > 8) return -1;
    const int64_t rows = (op == 1) ? M : Kd;
    int64_t blocks = fsw_cdiv(rows, 8 * 16);        // 8 warps per block, >= 16 rows per warp
    const int64_t cap = (op == 1) ? 148 * 8 : 148 * 2;   // op 2 ends with one atomic per warp and output: fewer, longer warps
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    fsw_prof_begin(op == 1 ? "gemm_nn" : "gemm_tn", st);
#define FSW_SMALLN(KJ_)                                                                                                              \
    case KJ_:                                                                                                                        \
        if (op == 1)                                                                                                                 \
            fsw_gemm_nn_small_kernel<N, KJ_><<<(unsigned)blocks, 256, 0, st>>>(M, (int)Kd, A, lda, B, ldb, C, ldc, accumulate);      \
        else                                                                                                                         \
            fsw_gemm_tn_small_kernel<N, KJ_><<<(unsigned)blocks, 256, 0, st>>>(Kd, (int)M, A, lda, B, ldb, C, ldc);                  \
        break;
    switch (KJ) {
        FSW_SMALLN(1) FSW_SMALLN(2) FSW_SMALLN(3) FSW_SMALLN(4) FSW_SMALLN(5) FSW_SMALLN(6) FSW_SMALLN(7) FSW_SMALLN(8)
    }
#undef FSW_SMALLN
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_gemm_small_n_kernel");
    return FSW_OK;
}

// ops 1 and 2 with N <= 8 columns and the lane axis <= 256 wide; returns -1 when not covered
int gemm_small_n_f32(int op, int64_t M, int64_t N, int64_t Kd, const float* A, int64_t lda, const float* B, int64_t ldb, float* C,
                     int64_t ldc, int accumulate, cudaStream_t st) {
    if ((op != 1 && op != 2) || N < 1 || N > 8) return -1;
    switch ((int)N) {
        case 1: return launch_small_n<1>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 2: return launch_small_n<2>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 3: return launch_small_n<3>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 4: return launch_small_n<4>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 5: return launch_small_n<5>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 6: return launch_small_n<6>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 7: return launch_small_n<7>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        case 8: return launch_small_n<8>(op, M, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
    }
    return -1;
}

// Small-d projection (d_in <= 8, e.g. 3-d point clouds): memory bound on the Xp write.  One thread
// produces 4 consecutive slices of one row; theta is read through the read-only cache.
template <typename T, int D>
__global__ void __launch_bounds__(256) fsw_project_small_kernel(int64_t M, int64_t N, const T* __restrict__ X, int64_t ldx,
                                                                const T* __restrict__ Th, int64_t ldt, T* __restrict__ C,
                                                                int64_t ldc) {
    const int64_t ngrp = (N + 3) / 4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * ngrp) return;
    const int64_t m = idx / ngrp;
    const int64_t n4 = (idx - m * ngrp) * 4;
    T x[D];
#pragma unroll
    for (int i = 0; i < D; ++i) x[i] = __ldg(X + m * ldx + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t n = n4 + j;
        if (n < N) {
            T acc = (T)0;
#pragma unroll
            for (int i = 0; i < D; ++i) acc = fma(x[i], __ldg(Th + n * ldt + i), acc);
            C[m * ldc + n] = acc;
        }
    }
}

template <typename T>
int gemm_t(int op, int64_t M, int64_t N, int64_t Kd, const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc,
           int accumulate, cudaStream_t st) {
    if (M == 0 || N == 0) return FSW_OK;
    int64_t sa_m, sa_k, sb_n, sb_k;
    switch (op) {
        case 0: sa_m = lda; sa_k = 1; sb_n = ldb; sb_k = 1; break;
        case 1: sa_m = lda; sa_k = 1; sb_n = 1; sb_k = ldb; break;
        case 2: sa_m = 1; sa_k = lda; sb_n = 1; sb_k = ldb; break;
        default: return fsw_fail(FSW_ERR_INVALID, "fsw_gemm: op %d", op);
    }
    if (Kd == 0) {
        if (!accumulate && op != 2) FSW_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(T), 0, (size_t)N * sizeof(T), (size_t)M, st));
        return FSW_OK;
    }
    if (op == 0 && !accumulate && Kd <= 8) {
        const int64_t total = M * ((N + 3) / 4);
        const unsigned grid = (unsigned)fsw_cdiv(total, 256);
#define FSW_PS(D)                                                                                       \
    case D:                                                                                             \
        fsw_project_small_kernel<T, D><<<grid, 256, 0, st>>>(M, N, A, lda, B, ldb, C, ldc);             \
        break;
        fsw_prof_begin(sizeof(T) == 4 ? "project_small_f32" : "project_small_f64", st);
        switch ((int)Kd) {
            FSW_PS(1) FSW_PS(2) FSW_PS(3) FSW_PS(4) FSW_PS(5) FSW_PS(6) FSW_PS(7) FSW_PS(8)
        }
#undef FSW_PS
        fsw_prof_end(st);
        FSW_CHECK_LAUNCH("fsw_project_small_kernel");
        return FSW_OK;
    }
    if constexpr (sizeof(T) == 4) {
        if (fsw_umma_enabled()) {
            // large fp32 contractions: tcgen05 + TMA at fp32 accuracy (fsw_umma.cu)
            const float* Ap = A;
            const float* Bp = B;
            const int urc = fsw_umma_gemm(op, M, N, 1, &Kd, &Ap, &lda, &Bp, &ldb, C, ldc, nullptr, accumulate, st);
            if (urc != FSW_UMMA_NA) return urc;
        }
        int rc = gemm_small_n_f32(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        if (rc >= 0) return rc;
        rc = gemm_strip_f32(op, M, N, Kd, A, lda, B, ldb, C, ldc, accumulate, st);
        if (rc >= 0) return rc;
    }
    int64_t splits = 1;
    if (op == 2) {
        // split the long reduction axis so that the grid fills the machine
        const int64_t tiles = fsw_cdiv(M, BM) * fsw_cdiv(N, BN);
        splits = (148 * 4 + tiles - 1) / tiles;
        const int64_t max_splits = fsw_cdiv(Kd, 4 * BK);
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int64_t k_per_split = fsw_cdiv(fsw_cdiv(Kd, splits), BK) * BK;
    splits = fsw_cdiv(Kd, k_per_split);
    dim3 grid((unsigned)fsw_cdiv(M, BM), (unsigned)fsw_cdiv(N, BN), (unsigned)splits);
    const int use_atomics = (op == 2) ? 1 : 0;
    static const char* labels[3] = {"gemm_nt", "gemm_nn", "gemm_tn"};
    fsw_prof_begin(labels[op], st);
    fsw_gemm_kernel<T><<<grid, 256, 0, st>>>(M, N, Kd, A, sa_m, sa_k, B, sb_n, sb_k, C, ldc, accumulate, k_per_split, use_atomics);
    fsw_prof_end(st);
    FSW_CHECK_LAUNCH("fsw_gemm_kernel");
    return FSW_OK;
}

}  // namespace

namespace {
bool g_umma_on = getenv("FSW_DISABLE_UMMA") == nullptr;

__global__ void fsw_add_bias_kernel(float* C, int64_t ldc, const float* bias, int64_t M, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M * N) C[(i / N) * ldc + (i % N)] += bias[i % N];
}
}  // namespace

bool fsw_umma_enabled() { return g_umma_on; }

extern "C" int fsw_set_tensor_cores(int on) {
    g_umma_on = on != 0;
    return FSW_OK;
}

// C[M, N] (ldc) (+)= sum_s A_s . B_s^T (+ bias): the contraction axis is the concatenation of up to two segments, so that
// cat(A_0, A_1) . W^T runs without materialising the concatenation (fsw_conv.py:357-361); fp32 only.
extern "C" int fsw_gemm_fused(int dtype, int64_t M, int64_t N, int nseg, const int64_t* Kd, const void* const* A, const int64_t* lda,
                              const void* const* B, const int64_t* ldb, void* C, int64_t ldc, const void* bias, int accumulate,
                              void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != FSW_F32) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_gemm_fused: fp32 only");
    if (nseg < 1 || nseg > 2) return fsw_fail(FSW_ERR_INVALID, "fsw_gemm_fused: nseg %d", nseg);
    if (M == 0 || N == 0) return FSW_OK;
    if (g_umma_on) {
        const int rc = fsw_umma_gemm(0, M, N, nseg, Kd, (const float* const*)A, lda, (const float* const*)B, ldb, (float*)C, ldc,
                                     (const float*)bias, accumulate, st);
        if (rc != FSW_UMMA_NA) return rc;
    }
    for (int s = 0; s < nseg; ++s) {
        const int rc = gemm_t<float>(0, M, N, Kd[s], (const float*)A[s], lda[s], (const float*)B[s], ldb[s], (float*)C, ldc,
                                     (accumulate || s > 0) ? 1 : 0, st);
        if (rc != FSW_OK) return rc;
    }
    if (bias) {
        fsw_add_bias_kernel<<<(unsigned)fsw_cdiv(M * N, 256), 256, 0, st>>>((float*)C, ldc, (const float*)bias, M, N);
        FSW_CHECK_LAUNCH("fsw_add_bias_kernel");
    }
    return FSW_OK;
}

extern "C" int fsw_gemm(int dtype, int op, int64_t M, int64_t N, int64_t Kd, const void* A, int64_t lda, const void* B,
                        int64_t ldb, void* C, int64_t ldc, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FSW_F32) return gemm_t<float>(op, M, N, Kd, (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc, accumulate, st);
    if (dtype == FSW_F64) return gemm_t<double>(op, M, N, Kd, (const double*)A, lda, (const double*)B, ldb, (double*)C, ldc, accumulate, st);
    return fsw_fail(FSW_ERR_INVALID, "fsw_gemm: dtype %d", dtype);
}
