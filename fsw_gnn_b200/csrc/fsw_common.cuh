// Shared declarations for the libfsw_embedding.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fsw_embedding.h"

// ---- host-side error plumbing (fsw_api.cu) -------------------------------------------------------
int fsw_fail(int code, const char* fmt, ...);
void fsw_count_launch(int n = 1);

#define FSW_CHECK_LAUNCH(name)                                                                      \
    do {                                                                                            \
        cudaError_t e__ = cudaGetLastError();                                                       \
        if (e__ != cudaSuccess)                                                                     \
            return fsw_fail(FSW_ERR_CUDA, "%s: kernel launch failed: %s", name, cudaGetErrorString(e__)); \
        fsw_count_launch();                                                                         \
    } while (0)

#define FSW_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fsw_fail(FSW_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));         \
    } while (0)

// per-kernel event timers (fsw_api.cu); no-ops unless fsw_profile_enable(1)
void fsw_prof_begin(const char* label, cudaStream_t st);
void fsw_prof_end(cudaStream_t st);

// tensor-core contraction (fsw_umma.cu); FSW_UMMA_NA: shapes / alignment not covered, the SIMT kernels take over
#define FSW_UMMA_NA 1
int fsw_umma_gemm(int op, int64_t M, int64_t N, int nseg, const int64_t* Kd, const float* const* A, const int64_t* lda,
                  const float* const* B, const int64_t* ldb, float* C, int64_t ldc, const float* bias, int accumulate,
                  cudaStream_t st);
bool fsw_umma_enabled();

static inline int64_t fsw_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- info word of the segment plan ----------------------------------------------------------------
#define FSW_INFO_UNIFORM (1 << 30)
#define FSW_INFO_NMASK ((1 << 30) - 1)

// plan bucket of a segment of n_eff elements (0..512 exact, then power-of-two ranges up to 8192, <= 32768, hubs beyond)
__host__ __device__ static inline int fsw_size_bucket(int n_eff) {
    if (n_eff < FSW_PLAN_EXACT) return n_eff;
    if (n_eff <= 1024) return 513;
    if (n_eff <= 2048) return 514;
    if (n_eff <= 4096) return 515;
    if (n_eff <= 8192) return 516;
    if (n_eff <= 32768) return 517;
    return 518;
}

// ---- numeric helpers --------------------------------------------------------------------------------
template <typename T>
struct Num;

template <>
struct Num<float> {
    static __device__ __forceinline__ float big() { return FLT_MAX; }
    static __device__ __forceinline__ float cospi_(float x) { return cospif(x); }
    static __device__ __forceinline__ float sinpi_(float x) { return sinpif(x); }
    // phase given in units of pi, in double; reduce to [-1, 1] before dropping to fp32 so that the
    // large-frequency phases (xi up to ~2K, SURVEY.md 7 hard part 2) keep full fp32 accuracy.
    static __device__ __forceinline__ float reduce(double phi) { return (float)(phi - 2.0 * rint(0.5 * phi)); }
};

template <>
struct Num<double> {
    static __device__ __forceinline__ double big() { return DBL_MAX; }
    static __device__ __forceinline__ double cospi_(double x) { return cospi(x); }
    static __device__ __forceinline__ double sinpi_(double x) { return sinpi(x); }
    static __device__ __forceinline__ double reduce(double phi) { return phi; }
};

// row * stride for operands known to fit 31 bits (node / edge indices of the int32 CSR, padded widths): one
// IMAD.WIDE where an int64 stride would cost a full 64x64 multiply (5-7 instructions) per address.
// (Spelled as mul.wide.s32: written as a C++ product of two sign-extended ints the compiler still emitted the full 64 x 64
// sequence - IMAD.WIDE.U32 plus two sign fix-ups - in several kernels, profiles/r2/README.md.)
__device__ __forceinline__ int64_t fsw_rowoff(int64_t row, int64_t ld) {
    int64_t r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"((int)row), "r"((int)ld));
    return r;
}

// cos(pi t) for |t| <= 1 (phase already reduced): no conversions, no range reduction - fold to [0, 1/2] and
// evaluate the even Taylor polynomial up to t^12 (truncation error < 1e-8 there).  ~11 full-rate instructions.
__device__ __forceinline__ float fsw_cospi_unit(float t) {
    const float x = fabsf(t);
    const bool hi = x > 0.5f;
    const float y = hi ? 1.0f - x : x;
    const float y2 = y * y;
    float p = fmaf(y2, 1.929574e-3f, -2.580689e-2f);
    p = fmaf(y2, p, 2.353306e-1f);
    p = fmaf(y2, p, -1.335263f);
    p = fmaf(y2, p, 4.058712f);
    p = fmaf(y2, p, -4.934802f);
    p = fmaf(y2, p, 1.0f);
    return hi ? -p : p;
}

// sinc(x) = sin(pi x)/(pi x)   (torch.sinc, fsw_embedding.py:1002, :1767)
template <typename T>
__device__ __forceinline__ T fsw_sinc(T x) {
    T ax = fabs(x);
    if (ax < (T)1e-3) {
        T y = (T)M_PI * x;
        T y2 = y * y;
        return (T)1 - y2 * ((T)(1.0 / 6.0) - y2 * (T)(1.0 / 120.0));
    }
    return Num<T>::sinpi_(x) / ((T)M_PI * x);
}

// d/dx sinc(x) = (cos(pi x) - sinc(x)) / x   (what autograd of torch.sinc gives, sp.dsinc :2763-2774)
template <typename T>
__device__ __forceinline__ T fsw_dsinc(T x) {
    T ax = fabs(x);
    if (ax < (T)0.25) {
        // -pi^2 x/3 + pi^4 x^3/30 - pi^6 x^5/840 + pi^8 x^7/45360 - pi^10 x^9/3991680
        T y = (T)M_PI * x;
        T y2 = y * y;
        T s = (T)(1.0 / 3.0) - y2 * ((T)(1.0 / 30.0) - y2 * ((T)(1.0 / 840.0) - y2 * ((T)(1.0 / 45360.0) - y2 * (T)(1.0 / 3991680.0))));
        return -(T)M_PI * y * s;
    }
    return (Num<T>::cospi_(x) - fsw_sinc(x)) / x;
}

// Amplitude of one element of normalised weight w at frequency xi:
//   a  = 2 w sinc(xi w)          (so that D_j = a cos(pi xi (2C - w)), fsw_embedding.py:1047-1075)
//   ap = da/dxi = 2 w^2 sinc'(xi w)
// x = xi*w is passed in DOUBLE: for large xi*w the numerator sin(pi x) needs the phase reduced mod 2
// before it is rounded to fp32 (same reason as for the cosine phase).
template <typename T, bool NEED_AP>
__device__ __forceinline__ void fsw_amplitude(double x, T w, T xi, T& a, T& ap) {
    if (fabs(x) < 0.25) {
        const T xf = (T)x;
        a = (T)2 * w * fsw_sinc(xf);
        if (NEED_AP) ap = (T)2 * w * w * fsw_dsinc(xf);
    } else {
        const T r = Num<T>::reduce(x);
        const T sincx = Num<T>::sinpi_(r) / ((T)M_PI * (T)x);
        a = (T)2 * w * sincx;
        if (NEED_AP) ap = (T)2 * w * (Num<T>::cospi_(r) - sincx) / xi;
    }
}

// Arguments shared by the forward and backward embedding kernels.
template <typename T>
struct SegArgs {
    const T* Xp;            // [Nrows, ldp]
    const T* Ep;            // [E, ldp] or nullptr
    const int32_t* rowptr;  // [S+1] or nullptr
    const int32_t* col;     // [E] or nullptr
    const T* W;             // [E] or nullptr
    const double* mass;     // [S]
    const int32_t* info;    // [S]
    const int32_t* order;   // [S] or nullptr (identity)
    const T* freqs;         // [K]
    int64_t ldp;
    int64_t n_fixed;
    int K;
    double thresh;
    // point-cloud mode (dense batches of low-dimensional points, fsw_embed_forward_cloud): the packed-key forward forms the keys
    // <x_e, theta_k> on the fly from the points instead of reading a projected matrix, and records the ranks slice-major
    // [S][K][n] so that its stores and the backward's loads are contiguous
    const T* projX = nullptr;       // [S * n, proj_d] points (Xp is not read)
    const T* projTheta = nullptr;   // [K, proj_ldt] slices
    int proj_d = 0;
    int64_t proj_ldt = 0;
    int rank_transposed = 0;
};

// medium / large path (fsw_embed_medium.cu): uniform-weight fp32 segments of more than 64 elements
int fsw_medium_forward_f32(const SegArgs<float>& a, int lo, int hi, int cap, float* out, int64_t ld_out, int64_t out_col0,
                           const float* bias, void* scratch, size_t scratch_bytes, unsigned short* ranks, int64_t ldr,
                           float* dxi_out, int64_t ld_dxi, const float* gtab_c, const float* gtab_t, cudaStream_t st);
int fsw_medium_backward_f32(const SegArgs<float>& a, int lo, int hi, int cap, const float* g, int64_t ld_g, int64_t g_col0,
                            float* dXp, float* dEp, double* dfreqs, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t fsw_medium_tile_bytes(int cap, int mode);  // global scratch per CTA (0: shared-memory tile)
int fsw_medium_grid();

// uniform-weight small path + rank-based backward (fsw_embed_small.cu)
template <typename T>
int fsw_small_forward_u(const SegArgs<T>& a, int np, int lo, int hi, T* out, int64_t ld_out, int64_t out_col0, const T* bias,
                        unsigned short* ranks, int64_t ldr, T* dxi_out, int64_t ld_dxi, cudaStream_t st);
template <typename T>
int fsw_rank_backward_u(const SegArgs<T>& a, int lo, int hi, int cap, const unsigned short* ranks, int64_t ldr, const T* g,
                        int64_t ld_g, int64_t g_col0, T* dXp, T* dEp, double* dfreqs, cudaStream_t st);

// packed-key register sort (fsw_embed_packed.cu): uniform-weight fp32 segments with n <= np, np in {96, 128}
int fsw_packed_forward_u(const SegArgs<float>& a, int np, int lo, int hi, float* out, int64_t ld_out, int64_t out_col0,
                         const float* bias, unsigned short* ranks, int64_t ldr, float* dxi_out, int64_t ld_dxi, const float* gtab_c,
                         const float* gtab_t, int tab_n0, int tab_ld4, cudaStream_t st);

int fsw_build_coef_tables(const float* freqs, int K, int ldp, int nmax, float* tab_c, float* tab_t, float* tab_A, float* tab_Ap,
                          cudaStream_t st, float2* tab_u = nullptr);
size_t fsw_rank_tables_bytes(int64_t ldp);
// forward coefficient tables (cos and d/dxi) for n <= FSW_FWD_TAB_NMAX, [n][FSW_FWD_TAB_LD/4][K][4], in front of the
// forward scratch (fsw_embed_packed.cu)
#define FSW_FWD_TAB_NMAX 512
// extent of the tables for a plan: the largest uniform segment size in 33..FSW_FWD_TAB_NMAX that occurs (0: none)
#define FSW_FWD_TAB_GRAPH_NMAX 1024   // graphs: the 513..1024 bucket (not resolved by size) takes tables up to 1024
static inline int fsw_fwd_tables_nmax(const int32_t* bo) {
    if (bo[FSW_PLAN_EXACT] < bo[FSW_PLAN_EXACT + 1]) return FSW_FWD_TAB_GRAPH_NMAX;
    for (int n = FSW_FWD_TAB_NMAX; n > 32; --n)
        if (bo[n] < bo[n + 1]) return n;
    return 0;
}
// both tables (cos, d/dxi): (nmax + 1) size classes x ceil(nmax / 4) position blocks x K slices x 4 positions
static inline size_t fsw_fwd_tables_bytes(int64_t K, int nmax) {
    const int64_t ld4 = (nmax + 3) / 4;
    return (size_t)(2 * (int64_t)(nmax + 1) * ld4 * K * 4) * sizeof(float);
}
// a dense batch of FSW_FWD_TAB_NMAX+1 .. 1024 points per multiset: one size class
static inline size_t fsw_fwd_tables_bytes_single(int64_t K, int64_t n) {
    return (size_t)(2 * ((n + 3) / 4) * K * 4) * sizeof(float);
}
int fsw_build_fwd_tables(const float* freqs, int K, int n_lo, int n_hi, int ld4, float* tab_c, float* tab_t, cudaStream_t st);
int fsw_rank_backward_g128(const SegArgs<float>& a, int lo, int hi, const unsigned short* ranks, int64_t ldr, const float* g,
                           int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, double* dfreqs, void* tables, cudaStream_t st);

int fsw_rank_backward_dense(const SegArgs<float>& a, int64_t S, int n, const unsigned short* ranks, int64_t ldr, const float* g,
                            int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, cudaStream_t st);
#define FSW_RANKT_TAB 512   // xi/n and A0(n) come from tables up to this n, are computed on the fly beyond
int fsw_rank_backward_T(const SegArgs<float>& a, int64_t S, int64_t Nrows, int nmax, const int32_t* tptr, const int32_t* tseg,
                        const int32_t* tslot, const int32_t* tn, const unsigned short* ranks, int64_t ldr, const float* g,
                        int64_t ld_g, int64_t g_col0, float* dXp, float* dEp, void* tables, float* ga_buf, cudaStream_t st);

template <typename T>
__device__ __forceinline__ void fsw_seg_range(const SegArgs<T>& a, int s, int64_t& e0, int& n) {
    if (a.rowptr) {
        e0 = a.rowptr[s];
        n = a.rowptr[s + 1] - (int)e0;
    } else {
        e0 = (int64_t)s * a.n_fixed;
        n = (int)a.n_fixed;
    }
}
