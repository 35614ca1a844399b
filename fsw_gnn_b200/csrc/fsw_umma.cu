// K1 on the 5th-generation tensor cores: C (+)= A . B at fp32 accuracy with tcgen05.mma (kind::tf32), operands staged
// by TMA, accumulators in tensor memory (north star subsystem 1; replaces the cuBLAS sgemm behind torch.tensordot,
// fsw_embedding.py:911, and behind nn.Linear in FSW_conv, fsw_conv.py:361).
//
// fp32 accuracy from TF32 tensor cores: every fp32 operand tile is split in shared memory into hi = tf32(a) and
// lo = tf32(a - hi) (22 of the 24 mantissa bits survive), and three products are issued per k-step:
//     D_main += A_hi . B_hi          D_corr += A_hi . B_lo + A_lo . B_hi
// into TWO accumulators in tensor memory that are added in the epilogue, so the 2^-11-times smaller correction terms
// never ride through the rounding of the large accumulator.  The dropped A_lo . B_lo term is 2^-22 relative.
//
// One persistent CTA per SM, 18 warps (14 in mode 2):
//   warp 0      TMA producer   global -> shared (128-byte swizzle), ring of 2 stages, mbarrier complete_tx
//   warps 2-5, 14-17  splitter  hi/lo split of the landed tiles in place (generic proxy) + fence.proxy.async
//   warp 1      MMA issuer     one thread: 3 tcgen05.mma per k-step, tcgen05.commit releases the stage / publishes the tile
//   warps 6-13  epilogue       tcgen05.ld -> registers (main + corr [+ bias]) -> swizzled staging -> TMA store / reduce-add
// Modes: 0 = NT (A [M,Kd], B [N,Kd]: both K-major), 1 = NN (B [Kd,N]: MN-major), 2 = TN (A [Kd,M], B [Kd,N]: both MN-major,
// the long contraction split over CTAs; the TMEM accumulators are drained into registers every UM_FLUSH_TN k-blocks, partial
// tiles are reduced with TMA reduce-add).  Output tiles of up to 128 columns ping-pong between two accumulator stages in
// tensor memory, so the epilogue of one tile / round runs under the MMAs of the next.
#include <cuda.h>
#include <stdlib.h>

#include "fsw_common.cuh"
#include "fsw_sm100.cuh"

namespace {
using namespace sm100;

constexpr int UM_BM = 128;       // UMMA M = rows of the output tile
// k-blocks of 16 floats (two k-steps): half-size stages, twice as many of them in the same shared memory - the kernel is
// bound by the TMA -> split -> MMA latency chain of a stage, not by bytes (measured: 2 stages of 32 floats 1.21 ms for the
// configs[3] projection)
constexpr int UM_BK = 16;
constexpr int UM_ROWB = UM_BK * 4;       // bytes of one K-major tile row (64: SWIZZLE_64B)
constexpr int UM_BOXB = 32 * UM_ROWB;    // bytes of one MN-major TMA box: [UM_BK k-rows][32 floats]
constexpr int UM_MAX_STAGES = 8;
// NT / NN: 18 warps, 8 of them splitters (warps 2-5 and 14-17: the hi/lo split is the slowest stage of the ring for long
// contractions, C5 project 1.80 -> 1.54 ms); TN keeps 14 warps (its epilogue accumulators need the registers: 1.39 vs 1.46 ms)
template <int MODE>
constexpr int um_threads() { return MODE == 2 ? 448 : 576; }
template <int MODE>
constexpr int um_split_threads() { return MODE == 2 ? 128 : 256; }
// The tensor core adds into the fp32 accumulator with truncation (measured: a chain of 1024 k-steps of positive terms ends
// 5.5e-5 low, 5.4e-8 per k-step), so no chain is longer than a round: the accumulators are drained after UM_FLUSH_TN
// k-blocks (32 k-steps) in the long reductions of mode 2, after UM_FLUSH k-blocks (contraction length 512) otherwise.
constexpr int UM_FLUSH_TN = 16;
constexpr int UM_FLUSH = 32;
constexpr uint32_t UM_ACC_STAGE_COLS = 256;
constexpr uint32_t UM_A_BYTES = UM_BM * UM_ROWB;
constexpr uint32_t UM_STG_BYTES = UM_BM * 128;   // one epilogue staging buffer: [128 rows][32 columns]
constexpr uint32_t UM_STAGING = 2 * UM_STG_BYTES;
constexpr uint32_t UM_TMEM_COLS = 512;

struct UmmaParams {
    int M, N;
    int nseg;
    int nkb[2];
    int kd[2];
    int BN;    // UMMA N (multiple of 16, <= 256)
    int BNL;   // rows (K-major) / columns (MN-major) of B staged per k-block
    int mtiles, ntiles, splits, kb_per_split;
    int accumulate;
    const float* bias;
    uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;   // descriptor strides (bytes) of the staged tiles
    uint32_t dbg_idesc_xor, dbg_print;
    int nstages;          // depth of the TMA -> split -> MMA ring
    int round_len;        // k-blocks accumulated in tensor memory before the accumulators are drained
    int nacc;             // accumulator stages in tensor memory (2 when BN <= 128: drain under the next round's MMAs)
    uint32_t corr_col;    // column offset of the correction accumulator inside an accumulator stage
};

struct WorkItem {
    int mt, nt, kb0, kb1;
};

__device__ __forceinline__ WorkItem decode_item(const UmmaParams& p, int item) {
    WorkItem w;
    const int sp = item % p.splits;
    const int t = item / p.splits;
    w.nt = t % p.ntiles;
    w.mt = t / p.ntiles;
    const int total = p.nkb[0] + p.nkb[1];
    w.kb0 = sp * p.kb_per_split;
    w.kb1 = min(total, w.kb0 + p.kb_per_split);
    return w;
}

template <int MODE>
__global__ void __launch_bounds__(um_threads<MODE>(), 1)
fsw_umma_kernel(const __grid_constant__ CUtensorMap tA0, const __grid_constant__ CUtensorMap tB0,
                const __grid_constant__ CUtensorMap tA1, const __grid_constant__ CUtensorMap tB1,
                const __grid_constant__ CUtensorMap tC, const UmmaParams p) {
    constexpr bool A_MN = (MODE == 2), B_MN = (MODE >= 1);
    extern __shared__ uint8_t um_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)um_smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t bbytes = (uint32_t)p.BNL * UM_ROWB;
    const uint32_t stage_bytes = 2 * UM_A_BYTES + 2 * bbytes;
    const int UM_STAGES = p.nstages;
    uint8_t* staging = smem + UM_STAGES * stage_bytes;
    uint64_t* bars = (uint64_t*)(staging + UM_STAGING);
    uint64_t* full = bars;                       // [nstages] TMA landed
    uint64_t* xfrm = bars + UM_MAX_STAGES;       // [nstages] hi/lo split done
    uint64_t* empty = bars + 2 * UM_MAX_STAGES;  // [nstages] MMAs reading the stage completed
    uint64_t* tfull = bars + 3 * UM_MAX_STAGES;  // [2] accumulators of a round complete
    uint64_t* tempty = tfull + 2;            // [2] accumulators drained
    uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tA0);
        tma_prefetch_desc(&tB0);
        tma_prefetch_desc(&tA1);
        tma_prefetch_desc(&tB1);
        tma_prefetch_desc(&tC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < UM_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&xfrm[s], um_split_threads<MODE>());
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 256);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, UM_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int nitems = p.mtiles * p.ntiles * p.splits;
    const int round_len = p.round_len;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer ------------------------------------------------
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const WorkItem w = decode_item(p, item);
                for (int kb = w.kb0; kb < w.kb1; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    const int seg = (kb >= p.nkb[0]) ? 1 : 0;
                    const int kk = (kb - (seg ? p.nkb[0] : 0)) * UM_BK;
                    const CUtensorMap* ta = seg ? &tA1 : &tA0;
                    const CUtensorMap* tb = seg ? &tB1 : &tB0;
                    uint8_t* st = smem + s * stage_bytes;
                    mbar_arrive_expect_tx(&full[s], UM_A_BYTES + bbytes);
                    if (A_MN) {
#pragma unroll
                        for (int b = 0; b < UM_BM / 32; ++b) tma_load_2d(st + b * UM_BOXB, ta, w.mt * UM_BM + 32 * b, kk, &full[s]);
                    } else {
                        tma_load_2d(st, ta, kk, w.mt * UM_BM, &full[s]);
                    }
                    uint8_t* sb = st + 2 * UM_A_BYTES;
                    if (B_MN) {
                        for (int b = 0; b < p.BNL / 32; ++b) tma_load_2d(sb + b * UM_BOXB, tb, w.nt * p.BN + 32 * b, kk, &full[s]);
                    } else {
                        tma_load_2d(sb, tb, kk, w.nt * p.BN, &full[s]);
                    }
                    if (++s == UM_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer ---------------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(UM_BM, p.BN, A_MN, B_MN) ^ p.dbg_idesc_xor;
            int s = 0;
            uint32_t ph = 0, aph = 0;
            int as = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const WorkItem w = decode_item(p, item);
                for (int r0 = w.kb0; r0 < w.kb1; r0 += round_len) {
                    const int r1 = min(w.kb1, r0 + round_len);
                    mbar_wait(&tempty[as], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_main = tmem + as * UM_ACC_STAGE_COLS, d_corr = d_main + p.corr_col;
                    for (int kb = r0; kb < r1; ++kb) {
                        mbar_wait(&full[s], ph);
                        mbar_wait(&xfrm[s], ph);
                        tc_fence_after();
                        const int seg = (kb >= p.nkb[0]) ? 1 : 0;
                        const int kk = (kb - (seg ? p.nkb[0] : 0)) * UM_BK;
                        const int nks = (min(UM_BK, p.kd[seg] - kk) + 7) >> 3;
                        const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
                        const uint32_t a_lo = a_hi + UM_A_BYTES;
                        const uint32_t b_hi = a_hi + 2 * UM_A_BYTES;
                        const uint32_t b_lo = b_hi + bbytes;
                        if (p.dbg_print && blockIdx.x == 0 && kb == w.kb0) {
                            const float* fa = (const float*)(smem + s * stage_bytes);
                            const float* fb = (const float*)(smem + s * stage_bytes + 2 * UM_A_BYTES);
                            printf("idesc %08x nks %d A: %g %g %g %g | %g %g  B: %g %g %g %g | row1 %g %g | box1 %g %g\n", idesc, nks, fa[0], fa[1], fa[2], fa[3],
                                   fa[32], fa[33], fb[0], fb[1], fb[2], fb[3], fb[32], fb[33], fb[1024], fb[1025]);
                        }
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint32_t aoff = ks * p.a_kstep, boff = ks * p.b_kstep;
                            // layout type: 1 = 128-byte swizzle of 32-byte chunks (MN-major tf32), 4 = 64-byte swizzle (K-major rows)
                            const uint64_t dAh = umma_desc_sw128(a_hi + aoff, p.a_lbo, p.a_sbo, A_MN ? 1u : 4u);
                            const uint64_t dAl = umma_desc_sw128(a_lo + aoff, p.a_lbo, p.a_sbo, A_MN ? 1u : 4u);
                            const uint64_t dBh = umma_desc_sw128(b_hi + boff, p.b_lbo, p.b_sbo, B_MN ? 1u : 4u);
                            const uint64_t dBl = umma_desc_sw128(b_lo + boff, p.b_lbo, p.b_sbo, B_MN ? 1u : 4u);
                            const uint32_t acc = (kb == r0 && ks == 0) ? 0u : 1u;
                            umma_tf32(d_main, dAh, dBh, idesc, acc);
                            umma_tf32(d_corr, dAh, dBl, idesc, acc);
                            umma_tf32(d_corr, dAl, dBh, idesc, 1u);
                        }
                        umma_commit(&empty[s]);
                        if (++s == UM_STAGES) { s = 0; ph ^= 1; }
                    }
                    umma_commit(&tfull[as]);
                    if (++as == p.nacc) { as = 0; aph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp < 6 || warp >= 14) {
        // ------------------------------------------------ hi / lo split -------------------------------------------------
        const int t = (warp < 6) ? threadIdx.x - 64 : threadIdx.x - 448 + 128;
        int s = 0;
        uint32_t ph = 0;
        const int nvec = (int)((2 * UM_A_BYTES + 2 * bbytes) / 32);   // float4 count of the A tile + the B tile
        const int nvecA = UM_A_BYTES / 16;
        const uint32_t smem_s = smem_u32(smem);   // 32-bit shared-window addresses: LDS / STS instead of generic loads and stores
        // fp32 -> tf32, round to nearest with ties away (what cvt.rna.tf32.f32 returns for every finite input; the PTX
        // instruction is emulated with an extra Inf / NaN test and select per value, ~40 % of the split's instructions)
        auto rna = [](uint32_t b) { return (b + 0x1000u) & 0xffffe000u; };
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const WorkItem w = decode_item(p, item);
            for (int kb = w.kb0; kb < w.kb1; ++kb) {
                mbar_wait(&full[s], ph);
                const uint32_t st = smem_s + s * stage_bytes;
#pragma unroll 4
                for (int i = t; i < nvec; i += um_split_threads<MODE>()) {
                    // the A tile is followed by its lo tile, then the B tile and its lo tile
                    const uint32_t src = (i < nvecA) ? st + 16u * i : st + 2 * UM_A_BYTES + 16u * (i - nvecA);
                    const uint32_t dst = (i < nvecA) ? st + UM_A_BYTES + 16u * i : st + 2 * UM_A_BYTES + bbytes + 16u * (i - nvecA);
                    uint32_t v[4], hi[4], lo[4];
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(src) : "memory");
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        hi[c] = rna(v[c]);
                        lo[c] = rna(__float_as_uint(__uint_as_float(v[c]) - __uint_as_float(hi[c])));
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(src), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                }
                fence_proxy_async_smem();
                mbar_arrive(&xfrm[s]);
                if (++s == UM_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------ epilogue -------------------------------------------------------
        const int e = warp - 6;
        const int q = warp & 3;          // TMEM lane quarter this warp may read
        const int h = e >> 2;            // column-chunk parity served by this group of 4 warps
        const int row = q * 32 + lane;
        uint8_t* stg = staging + h * UM_STG_BYTES;
        const bool issuer = ((e & 3) == 0) && lane == 0;
        const int nchunks = (p.BN + 31) >> 5;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t aph = 0;
        int as = 0;
        float acc[2][32];
        (void)acc;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const WorkItem w = decode_item(p, item);
            if (w.kb0 >= w.kb1) continue;
            const int m0 = w.mt * UM_BM, n0 = w.nt * p.BN;
            if (MODE == 2) {
#pragma unroll
                for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[ci][i] = 0.f;
            }
            for (int r0 = w.kb0; r0 < w.kb1; r0 += round_len) {
                mbar_wait(&tfull[as], aph);
                tc_fence_after();
                const uint32_t lane_addr = lane_base + as * UM_ACC_STAGE_COLS;
                uint64_t* drained = &tempty[as];
                if (++as == p.nacc) { as = 0; aph ^= 1; }
                const bool first_round = (r0 == w.kb0);
                if (MODE == 2) {
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci) {
                        const int c = h + 2 * ci;
                        if (c < nchunks) {
                            uint32_t v[32], u[32];
                            tmem_ld32(lane_addr + c * 32, v);
                            tmem_ld32(lane_addr + p.corr_col + c * 32, u);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc[ci][i] += __uint_as_float(v[i]) + __uint_as_float(u[i]);
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(drained);
                } else {
                    for (int c = h; c < nchunks; c += 2) {
                        uint32_t v[32], u[32];
                        tmem_ld32(lane_addr + c * 32, v);
                        tmem_ld32(lane_addr + p.corr_col + c * 32, u);
                        tmem_ld_wait();
                        float x[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]) + __uint_as_float(u[i]);
                        if (p.bias != nullptr && first_round) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int col = n0 + c * 32 + i;
                                if (col < p.N) x[i] += __ldg(p.bias + col);
                            }
                        }
                        if (issuer) tma_wait_group_read0();
                        named_bar_sync(1 + h, 128);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *(float4*)(stg + row * 128 + ((j ^ (row & 7)) << 4)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + h, 128);
                        if (issuer) {
                            // later rounds of a long contraction add onto the first round's tile
                            if (p.accumulate || !first_round) tma_reduce_add_2d(&tC, stg, n0 + c * 32, m0);
                            else tma_store_2d(&tC, stg, n0 + c * 32, m0);
                            tma_commit_group();
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(drained);
                }
            }
            if (MODE == 2) {
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    const int c = h + 2 * ci;
                    if (c < nchunks) {
                        if (issuer) tma_wait_group_read0();
                        named_bar_sync(1 + h, 128);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *(float4*)(stg + row * 128 + ((j ^ (row & 7)) << 4)) =
                                make_float4(acc[ci][4 * j], acc[ci][4 * j + 1], acc[ci][4 * j + 2], acc[ci][4 * j + 3]);
                        fence_proxy_async_smem();
                        named_bar_sync(1 + h, 128);
                        if (issuer) {
                            tma_reduce_add_2d(&tC, stg, n0 + c * 32, m0);
                            tma_commit_group();
                        }
                    }
                }
            }
        }
        if (issuer) tma_wait_group0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem, UM_TMEM_COLS);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 row-major tensor [outer, inner] with row pitch `ld` elements; box [box_outer, box_inner], 128-byte swizzle:
// 16-byte chunks (K-major operands, the output), or 32-byte chunks for operands the tensor core reads MN-major
enum MapKind { MAP_OUT = 0, MAP_KMAJOR = 1, MAP_MNMAJOR = 2 };
bool make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
              MapKind kind) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               kind == MAP_MNMAJOR ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (kind == MAP_KMAJOR ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// C[M, N] (+)= sum over segments of A_s . B_s (+ bias), fp32, on the tensor cores.  Returns FSW_UMMA_NA when the shapes / alignments do
// not qualify (the caller then runs the SIMT kernels), FSW_OK or an error code otherwise.
//   op 0: A_s [M, Kd_s] (lda_s), B_s [N, Kd_s] (ldb_s)      op 1: B_s [Kd_s, N]      op 2: A [Kd, M], B [Kd, N] (one segment)
int fsw_umma_gemm(int op, int64_t M, int64_t N, int nseg, const int64_t* Kd, const float* const* A, const int64_t* lda,
                  const float* const* B, const int64_t* ldb, float* C, int64_t ldc, const float* bias, int accumulate,
                  cudaStream_t st) {
    if (op < 0 || op > 2 || nseg < 1 || nseg > 2 || (op == 2 && nseg != 1)) return FSW_UMMA_NA;
    if (M >= (1LL << 31) || N >= (1LL << 31)) return FSW_UMMA_NA;
    int64_t kd_total = 0;
    for (int s = 0; s < nseg; ++s) {
        if (!aligned16(A[s]) || !aligned16(B[s]) || (lda[s] & 3) || (ldb[s] & 3) || Kd[s] < 1 || Kd[s] >= (1LL << 31)) return FSW_UMMA_NA;
        kd_total += Kd[s];
    }
    if (!aligned16(C) || (ldc & 3) || (bias && op == 2)) return FSW_UMMA_NA;
    if (op == 2) {
        if (kd_total < 4096 || M < 16 || N < 16) return FSW_UMMA_NA;
    } else {
        if (M < 2048 || N < 32 || kd_total < 32) return FSW_UMMA_NA;
    }
    if (get_encode() == nullptr) return FSW_UMMA_NA;

    UmmaParams p{};
    p.M = (int)M;
    p.N = (int)N;
    p.nseg = nseg;
    for (int s = 0; s < 2; ++s) {
        p.kd[s] = s < nseg ? (int)Kd[s] : 0;
        p.nkb[s] = s < nseg ? (int)fsw_cdiv(Kd[s], UM_BK) : 0;
    }
    const int bn_max = (op == 2) ? 128 : 256;
    if (N <= bn_max) {
        p.BN = (int)((N + 15) / 16 * 16);
        p.ntiles = 1;
    } else {
        p.BN = bn_max;
        p.ntiles = (int)fsw_cdiv(N, bn_max);
    }
    p.BNL = (op >= 1) ? (p.BN + 31) / 32 * 32 : p.BN;
    p.mtiles = (int)fsw_cdiv(M, UM_BM);
    const int total_kb = p.nkb[0] + p.nkb[1];
    const int sms = num_sms();
    p.splits = 1;
    p.kb_per_split = total_kb;
    if (op == 2) {
        const int tiles = p.mtiles * p.ntiles;
        int splits = (2 * sms + tiles - 1) / tiles;
        const int max_splits = (total_kb + UM_FLUSH_TN - 1) / UM_FLUSH_TN;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        int per = (total_kb + splits - 1) / splits;
        per = (per + UM_FLUSH_TN - 1) / UM_FLUSH_TN * UM_FLUSH_TN;
        p.kb_per_split = per;
        p.splits = (total_kb + per - 1) / per;
    }
    p.accumulate = (accumulate || op == 2) ? 1 : 0;
    p.bias = bias;
    p.round_len = (op == 2) ? UM_FLUSH_TN : UM_FLUSH;
    p.nacc = (p.BN <= 128) ? 2 : 1;
    p.corr_col = (p.BN <= 128) ? 128u : 256u;
    // K-major tile: rows of 64 bytes (64-byte swizzle), 8-row atoms 512 bytes apart, one k-step = 32 bytes inside the row.
    // MN-major tile: TMA boxes of [16 k-rows][128 bytes] (32-byte-chunk swizzle, atoms of 4 k-rows = 512 bytes); 32-element
    // MN groups one box apart, one k-step = 8 rows = two atoms.
    p.a_lbo = (op == 2) ? (uint32_t)UM_BOXB : 16u;  p.a_sbo = 512u;  p.a_kstep = (op == 2) ? 1024u : 32u;
    p.b_lbo = (op >= 1) ? (uint32_t)UM_BOXB : 16u;  p.b_sbo = 512u;  p.b_kstep = (op >= 1) ? 1024u : 32u;
    if (const char* dbg = getenv("FSW_UMMA_DBG")) {   // debug: "a_lbo a_sbo a_kstep b_lbo b_sbo b_kstep"
        unsigned v[6];
        unsigned x = 0, pr = 0;
        if (sscanf(dbg, "%u %u %u %u %u %u %x %u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &x, &pr) >= 6) {
            p.a_lbo = v[0]; p.a_sbo = v[1]; p.a_kstep = v[2]; p.b_lbo = v[3]; p.b_sbo = v[4]; p.b_kstep = v[5];
            p.dbg_idesc_xor = x; p.dbg_print = pr;
        }
    }

    CUtensorMap tA[2], tB[2], tC;
    for (int s = 0; s < 2; ++s) {
        const int u = s < nseg ? s : 0;
        bool ok;
        if (op == 2) ok = make_map(&tA[s], A[u], (uint64_t)M, (uint64_t)Kd[u], (uint64_t)lda[u], 32, UM_BK, MAP_MNMAJOR);
        else ok = make_map(&tA[s], A[u], (uint64_t)Kd[u], (uint64_t)M, (uint64_t)lda[u], UM_BK, UM_BM, MAP_KMAJOR);
        if (op >= 1) ok = ok && make_map(&tB[s], B[u], (uint64_t)N, (uint64_t)Kd[u], (uint64_t)ldb[u], 32, UM_BK, MAP_MNMAJOR);
        else ok = ok && make_map(&tB[s], B[u], (uint64_t)Kd[u], (uint64_t)N, (uint64_t)ldb[u], UM_BK, (uint32_t)p.BNL, MAP_KMAJOR);
        if (!ok) return FSW_UMMA_NA;
    }
    if (!make_map(&tC, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, UM_BM, MAP_OUT)) return FSW_UMMA_NA;

    const size_t stage_bytes = 2 * UM_A_BYTES + 2 * (size_t)p.BNL * UM_ROWB;
    const size_t fixed = 1024 + UM_STAGING + 512;   // alignment slack, epilogue staging, barriers
    int nst = (int)((232448 - fixed) / stage_bytes);
    if (const char* e = getenv("FSW_UMMA_STAGES")) nst = atoi(e) < nst ? atoi(e) : nst;   // tuning knob
    p.nstages = nst > UM_MAX_STAGES ? UM_MAX_STAGES : (nst < 2 ? 2 : nst);
    const size_t smem = fixed + (size_t)p.nstages * stage_bytes;
    const int nitems = p.mtiles * p.ntiles * p.splits;
    const int grid = nitems < sms ? nitems : sms;
    static const char* labels[3] = {"umma_nt", "umma_nn", "umma_tn"};
#define FSW_UMMA_LAUNCH(MODE)                                                                                              \
    do {                                                                                                                   \
        FSW_CUDA(cudaFuncSetAttribute(fsw_umma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        fsw_prof_begin(labels[MODE], st);                                                                                  \
        fsw_umma_kernel<MODE><<<grid, um_threads<MODE>(), smem, st>>>(tA[0], tB[0], tA[1], tB[1], tC, p);                          \
        fsw_prof_end(st);                                                                                                  \
    } while (0)
    if (op == 0) FSW_UMMA_LAUNCH(0);
    else if (op == 1) FSW_UMMA_LAUNCH(1);
    else FSW_UMMA_LAUNCH(2);
#undef FSW_UMMA_LAUNCH
    FSW_CHECK_LAUNCH("fsw_umma_kernel");
    return FSW_OK;
}
