// Point-cloud mode, backward (dense batches of unit-weight multisets of low-dimensional points, configs[2]): no projected
// matrix and no projected gradient exist in memory.  With the slice-major ranks of the forward,
//     dL/dp[s, k, e] = g[s, k] (1 + xi_k) A0(n, k) cos(pi xi_k (2 r + 1) / n),   r = rank[s, k, e]
// is evaluated where it is consumed:
//   fsw_cloud_bwd_dx_kernel      thread = point e of cloud s, loop over the K slices: dX[s, e, :] = sum_k dL/dp theta_k
//   fsw_cloud_bwd_dtheta_kernel  block = slice k (x a range of clouds), threads over points: dtheta_k = sum_{s, e} dL/dp x_{s,e}
// Both read the ranks as contiguous 2-byte streams (64 bytes per warp instruction); the cosine costs ~25 instructions and is
// simply evaluated twice.  Reference semantics: autograd of fsw_embedding.py:925, :989-1004, :1101 (dense path).
#include "fsw_common.cuh"

namespace {

// cos(pi (2 r + 1) u) for the double-float u = uh + ul (u = xi / n), all in fp32: the product (2r+1) uh is split exactly with an
// FMA, reduced by its nearest integer through the magic-number add, and the parity of that integer flips the sign
__device__ __forceinline__ float fsw_cos_rank(unsigned r, float uh, float ul) {
    const float m = __uint_as_float(0x4B000001u | (r << 1)) - 8388608.0f;   // 2 r + 1, exact below 2^23
    const float ph = m * uh;
    const float pe = fmaf(m, uh, -ph);
    const float pl = fmaf(m, ul, pe);
    const float t = ph + 12582912.0f;
    const float k = t - 12582912.0f;                                         // rint(ph)
    const float x = (ph - k) + pl;                                           // in [-1/2, 1/2]
    const float y2 = x * x;
    float p = fmaf(y2, 1.929574e-3f, -2.580689e-2f);
    p = fmaf(y2, p, 2.353306e-1f);
    p = fmaf(y2, p, -1.335263f);
    p = fmaf(y2, p, 4.058712f);
    p = fmaf(y2, p, -4.934802f);
    p = fmaf(y2, p, 1.0f);
    return __uint_as_float(__float_as_uint(p) ^ (__float_as_uint(t) << 31));
}

// the same for two (rank, slice) pairs at once on Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2: one issue slot for
// both); nul = -ul.  Returns cos(pi (2 r + 1) u) WITHOUT the sign; `t` carries the parity of the rounded phase in its low bit.
__device__ __forceinline__ float2 fsw_cos_rank2(unsigned r0, unsigned r1, float2 uh, float2 nul, float2& t) {
    auto f2c = [](float c) { return make_float2(c, c); };
    const float2 mb = make_float2(__uint_as_float(0x4B000000u | r0), __uint_as_float(0x4B000000u | r1));   // 2^23 + r
    const float2 m = __ffma2_rn(mb, f2c(2.0f), f2c(-16777215.0f));     // 2 r + 1, exact
    const float2 nm = __ffma2_rn(mb, f2c(-2.0f), f2c(16777215.0f));
    const float2 ph = __fmul2_rn(m, uh);
    const float2 nqe = __ffma2_rn(nm, uh, ph);                         // ph - m uh, exact
    const float2 npl = __ffma2_rn(m, nul, nqe);                        // -(m ul + m uh - ph)
    t = __fadd2_rn(ph, f2c(12582912.0f));
    const float2 kk = __fadd2_rn(t, f2c(-12582912.0f));                // rint(ph)
    const float2 red = __ffma2_rn(kk, f2c(-1.0f), ph);
    const float2 x = __ffma2_rn(npl, f2c(-1.0f), red);
    const float2 y2 = __fmul2_rn(x, x);
    float2 c = __ffma2_rn(y2, f2c(1.929574e-3f), f2c(-2.580689e-2f));
    c = __ffma2_rn(y2, c, f2c(2.353306e-1f));
    c = __ffma2_rn(y2, c, f2c(-1.335263f));
    c = __ffma2_rn(y2, c, f2c(4.058712f));
    c = __ffma2_rn(y2, c, f2c(-4.934802f));
    c = __ffma2_rn(y2, c, f2c(1.0f));
    return c;
}
__device__ __forceinline__ float fsw_flip(float v, float t) { return __uint_as_float(__float_as_uint(v) ^ (__float_as_uint(t) << 31)); }

// per-slice constants of a cloud size n: double-float xi_k / n and the amplitude (1 + xi_k) A0(n, k)
__device__ __forceinline__ void fsw_slice_consts(float xi, int n, float& uh, float& ul, float& amp) {
    const double u = (double)xi / (double)n;
    uh = (float)u;
    ul = (float)(u - (double)uh);
    float A0, A0p;
    fsw_amplitude<float, false>(u, (float)(1.0 / (double)n), xi, A0, A0p);
    amp = (1.f + xi) * A0;
}

template <int D>
__global__ void __launch_bounds__(256) fsw_cloud_bwd_dx_kernel(const float* __restrict__ theta, int64_t ldt, const float* __restrict__ freqs,
                                                               int K, int n, const float* __restrict__ g, int64_t ld_g, int64_t g_col0,
                                                               const unsigned short* __restrict__ ranksT, float* __restrict__ dX) {
    extern __shared__ float fsw_cloud_smem[];
    float* s_uh = fsw_cloud_smem;            // [K]
    float* s_ul = s_uh + K;                  // [K]
    float* s_ga = s_ul + K;                  // [K]   g[s, k] (1 + xi_k) A0
    float* s_th = s_ga + K;                  // [K][D]
    const int s = blockIdx.y;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float uh, ul, amp;
        fsw_slice_consts(__ldg(freqs + k), n, uh, ul, amp);
        s_uh[k] = uh;
        s_ul[k] = ul;
        s_ga[k] = amp * __ldg(g + fsw_rowoff(s, ld_g) + g_col0 + k);
#pragma unroll
        for (int dd = 0; dd < D; ++dd) s_th[k * D + dd] = __ldg(theta + fsw_rowoff(k, ldt) + dd);
    }
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned short* rp = ranksT + (int64_t)s * K * n + e;
    float acc[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) acc[dd] = 0.f;
    int k = 0;
    for (; k + 16 <= K; k += 16) {   // 16 rank loads in flight, then two slices per packed instruction
        unsigned r[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) r[u] = rp[(int64_t)(k + u) * n];
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
            float2 t;
            const float2 c = fsw_cos_rank2(r[u], r[u + 1], make_float2(s_uh[k + u], s_uh[k + u + 1]), make_float2(-s_ul[k + u], -s_ul[k + u + 1]), t);
            const float dp0 = fsw_flip(s_ga[k + u], t.x) * c.x, dp1 = fsw_flip(s_ga[k + u + 1], t.y) * c.y;
#pragma unroll
            for (int dd = 0; dd < D; ++dd) acc[dd] = fmaf(dp1, s_th[(k + u + 1) * D + dd], fmaf(dp0, s_th[(k + u) * D + dd], acc[dd]));
        }
    }
    for (; k < K; ++k) {
        const unsigned r = rp[(int64_t)k * n];
        const float dp = s_ga[k] * fsw_cos_rank(r, s_uh[k], s_ul[k]);
#pragma unroll
        for (int dd = 0; dd < D; ++dd) acc[dd] = fmaf(dp, s_th[k * D + dd], acc[dd]);
    }
    float* o = dX + ((int64_t)s * n + e) * D;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) o[dd] = acc[dd];
}

template <int D>
__global__ void __launch_bounds__(256) fsw_cloud_bwd_dtheta_kernel(const float* __restrict__ X, const float* __restrict__ freqs, int K, int n,
                                                                   int64_t S, int clouds_per_block, const float* __restrict__ g,
                                                                   int64_t ld_g, int64_t g_col0, const unsigned short* __restrict__ ranksT,
                                                                   float* __restrict__ dtheta, int64_t ld_dt) {
    __shared__ float red[8][D];
    __shared__ float s_c[3];
    const int k = blockIdx.x;
    if (threadIdx.x == 0) {   // double-precision division and the amplitude: once per block, not per thread
        float uh0, ul0, amp0;
        fsw_slice_consts(__ldg(freqs + k), n, uh0, ul0, amp0);
        s_c[0] = uh0;
        s_c[1] = ul0;
        s_c[2] = amp0;
    }
    __syncthreads();
    const float uh = s_c[0], ul = s_c[1], amp = s_c[2];
    const int64_t s0 = (int64_t)blockIdx.y * clouds_per_block;
    const int64_t s1 = min(S, s0 + clouds_per_block);
    float acc[D];
#pragma unroll
    for (int dd = 0; dd < D; ++dd) acc[dd] = 0.f;
    for (int64_t s = s0; s < s1; ++s) {
        const float ga = amp * __ldg(g + fsw_rowoff(s, ld_g) + g_col0 + k);
        const unsigned short* rp = ranksT + ((int64_t)s * K + k) * n;
        const float* xs = X + (int64_t)s * n * D;
        int e = threadIdx.x;
        for (; e + (int)blockDim.x < n; e += 2 * blockDim.x) {   // two points per packed instruction
            const int e1 = e + blockDim.x;
            float2 t;
            const float2 c = fsw_cos_rank2(rp[e], rp[e1], make_float2(uh, uh), make_float2(-ul, -ul), t);
            const float dp0 = fsw_flip(ga, t.x) * c.x, dp1 = fsw_flip(ga, t.y) * c.y;
#pragma unroll
            for (int dd = 0; dd < D; ++dd)
                acc[dd] = fmaf(dp1, __ldg(xs + (int64_t)e1 * D + dd), fmaf(dp0, __ldg(xs + (int64_t)e * D + dd), acc[dd]));
        }
        for (; e < n; e += blockDim.x) {
            const float dp = ga * fsw_cos_rank(rp[e], uh, ul);
#pragma unroll
            for (int dd = 0; dd < D; ++dd) acc[dd] = fmaf(dp, __ldg(xs + (int64_t)e * D + dd), acc[dd]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int dd = 0; dd < D; ++dd) {
        float v = acc[dd];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[warp][dd] = v;
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        atomicAdd(dtheta + fsw_rowoff(k, ld_dt) + threadIdx.x, v);
    }
}

template <int D>
int cloud_backward_d(const float* X, const float* theta, int64_t ldt, const float* freqs, int K, int n, int64_t S, const float* g,
                     int64_t ld_g, int64_t g_col0, const unsigned short* ranksT, float* dX, float* dtheta, int64_t ld_dt, cudaStream_t st) {
    if (dX) {
        const size_t smem = (size_t)K * (3 + D) * sizeof(float);
        auto kern = fsw_cloud_bwd_dx_kernel<D>;
        if (smem > 40 * 1024) FSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fsw_prof_begin("bwd_cloud_dx", st);
        kern<<<dim3((unsigned)fsw_cdiv(n, 256), (unsigned)S), 256, smem, st>>>(theta, ldt, freqs, K, n, g, ld_g, g_col0, ranksT, dX);
        fsw_prof_end(st);
        FSW_CHECK_LAUNCH("fsw_cloud_bwd_dx_kernel");
    }
    if (dtheta) {
        // several waves of small blocks (K slices x groups of clouds) rather than one ragged wave: with 148 SMs x 8 resident
        // blocks, 1280 blocks of 52 clouds ran as one full wave plus a nearly empty one (0.27 ms; 0.1x ms with 8 clouds each)
        int cpb = (int)fsw_cdiv(S * (int64_t)K, 148 * 8 * 8);
        if (cpb < 1) cpb = 1;
        if (cpb > 8) cpb = 8;
        fsw_prof_begin("bwd_cloud_dtheta", st);
        fsw_cloud_bwd_dtheta_kernel<D><<<dim3((unsigned)K, (unsigned)fsw_cdiv(S, cpb)), 256, 0, st>>>(X, freqs, K, n, S, cpb, g, ld_g, g_col0, ranksT,
                                                                                                  dtheta, ld_dt);
        fsw_prof_end(st);
        FSW_CHECK_LAUNCH("fsw_cloud_bwd_dtheta_kernel");
    }
    return FSW_OK;
}

}  // namespace

// dX [S * n, d] is overwritten, dtheta [K, ld_dt] is ADDED to (either may be NULL); ranksT [S][K][n] from fsw_embed_forward_cloud
extern "C" int fsw_embed_backward_cloud(int dtype, const void* X, int64_t d, const void* theta, int64_t ldt, int64_t n, int64_t S,
                                        int64_t K, const void* freqs, const void* g, int64_t ld_g, int64_t g_col0, const void* ranksT,
                                        void* dX, void* dtheta, int64_t ld_dt, void* stream) {
    if (S == 0 || K == 0 || n == 0) return FSW_OK;
    if (dtype != FSW_F32) return fsw_fail(FSW_ERR_UNSUPPORTED, "fsw_embed_backward_cloud: fp32 only");
    if (!X || !theta || !freqs || !g || !ranksT) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_cloud: null argument");
    if (S > 65535) return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_cloud: more than 65535 clouds per call");
    cudaStream_t st = (cudaStream_t)stream;
#define FSW_CLOUD_CASE(D_) \
    case D_: return cloud_backward_d<D_>((const float*)X, (const float*)theta, ldt, (const float*)freqs, (int)K, (int)n, S, (const float*)g, ld_g, \
                                         g_col0, (const unsigned short*)ranksT, (float*)dX, (float*)dtheta, ld_dt, st);
    switch ((int)d) {
        FSW_CLOUD_CASE(1) FSW_CLOUD_CASE(2) FSW_CLOUD_CASE(3) FSW_CLOUD_CASE(4)
    }
#undef FSW_CLOUD_CASE
    return fsw_fail(FSW_ERR_INVALID, "fsw_embed_backward_cloud: d = %lld (1..4)", (long long)d);
}
