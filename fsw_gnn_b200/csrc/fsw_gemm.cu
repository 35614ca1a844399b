// K1: dense contractions at full input precision (fp32 FMA / fp64) for
//   op 0 (NT)  Xp      = X   . theta^T     slice projection, fsw_embedding.py:911/:913/:936
//   op 1 (NN)  dX      = dXp . theta       autograd of :911 w.r.t. X
//   op 2 (TN)  dtheta += dXp^T . X         autograd of :911 w.r.t. projVecs (split over the long N axis)
//
// The reference runs these as cuBLAS SGEMM with TF32 off, so the contraction must be true fp32:
// this is a register-tiled SIMT kernel (128x64x16 CTA tile, 8x4 per thread, k-major shared tiles).
// A tcgen05 3xTF32 variant for large d_in is future work.
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

extern "C" int fsw_gemm(int dtype, int op, int64_t M, int64_t N, int64_t Kd, const void* A, int64_t lda, const void* B,
                        int64_t ldb, void* C, int64_t ldc, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == FSW_F32) return gemm_t<float>(op, M, N, Kd, (const float*)A, lda, (const float*)B, ldb, (float*)C, ldc, accumulate, st);
    if (dtype == FSW_F64) return gemm_t<double>(op, M, N, Kd, (const double*)A, lda, (const double*)B, ldb, (double*)C, ldc, accumulate, st);
    return fsw_fail(FSW_ERR_INVALID, "fsw_gemm: dtype %d", dtype);
}
