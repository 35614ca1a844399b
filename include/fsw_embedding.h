/*
 * libfsw_embedding.so - B200 (sm_100a) native replacement for the reference's libfsw_embedding.so.
 *
 * Plain C ABI: raw device pointers, sizes and a cudaStream_t (passed as void*).  No torch types.
 * The caller (torch, in the Python host modules) owns every buffer; the library allocates nothing
 * and never synchronises the device.  Every new entry point returns 0 on success or a negative
 * FSW_ERR_* code; `fsw_last_error()` gives the message (the reference's lib printf()s and exit(1)s,
 * fsw_embedding.cu:20-27).
 *
 * Section 1 keeps the 7 symbols the reference's Python binds with ctypes
 * (fsw_embedding.py:2952-2977, :3033-3034) - same names, argument order and types - so the
 * reference's own `segcumsum_cuda` (fsw_embedding.py:2878-3012) runs unchanged on top of this lib.
 * Sections 2-6 are the fused entry points that replace the torch-op graph of
 * FSW_embedding.forward / forward_helper (fsw_embedding.py:778-1112) and its autograd
 * (class ag, fsw_embedding.py:1232-2258), and FSW_conv.edge_index_to_adj (fsw_conv.py:384-447).
 *
 * dtype codes follow the reference's `enum class torch_dtype` (fsw_embedding.cu:14-17):
 *   0 = float32, 1 = float64.
 */
#ifndef FSW_EMBEDDING_H
#define FSW_EMBEDDING_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSW_OK 0
#define FSW_ERR_INVALID (-1)   /* bad argument                                  */
#define FSW_ERR_CUDA (-2)      /* a CUDA runtime call / launch failed           */
#define FSW_ERR_WORKSPACE (-3) /* workspace too small                           */
#define FSW_ERR_UNSUPPORTED (-4)

#define FSW_F32 0
#define FSW_F64 1

/* Number of buckets of the segment plan (section 3). */
#define FSW_PLAN_EXACT 513                        /* buckets 0..512: exact n_eff                   */
#define FSW_PLAN_BUCKETS_PER_KIND 519             /* 513..518: n_eff <=1024, <=2048, <=4096, <=8192, <=32768, beyond */
#define FSW_PLAN_BUCKETS (2 * FSW_PLAN_BUCKETS_PER_KIND) /* kind 0 = uniform weights, 1 = general   */

/* ------------------------------------------------------------------------------------------------
 * 0. Library info / errors
 * ---------------------------------------------------------------------------------------------- */
int fsw_version(void);                /* 100 * major + minor                                     */
const char* fsw_last_error(void);     /* message of the last failing call on this host thread    */
int fsw_built_for_sm(void);           /* 100: the cubin is sm_100a only                          */

/* ------------------------------------------------------------------------------------------------
 * 1. Legacy ABI of the reference (fsw_embedding.cu:125-183, :194, :212, :231)
 *    Semantics per call are those of the reference kernels (fsw_embedding.cu:35-98, :103-117):
 *    block-local segmented inclusive scan in place + per-block tail sum / last id, and the
 *    down-sweep add.  They run on the legacy default stream and synchronise like the reference
 *    (fsw_embedding.cu:197, :208, :215, :227) because the reference's Python relies on that.
 * ---------------------------------------------------------------------------------------------- */
void segcumsum_wrapper(int64_t dtype, void* values, const int64_t* segment_ids, int64_t size,
                       int64_t max_seg_size, void* block_sums_out, int64_t* block_last_ids_out,
                       bool return_next_level, int64_t num_blocks, int64_t threads_per_block,
                       size_t shared_memory_size);
void add_block_sums_wrapper(int64_t dtype, void* output, const void* block_sums, const int64_t* segment_ids,
                            const int64_t* block_last_id, int64_t size, int64_t num_blocks,
                            int64_t threads_per_block);
int get_max_threads_per_block(int device_index);
void launch_segcumsum_kernel_float(float* values, const int64_t* segment_ids, int64_t size, int64_t max_seg_size,
                                   float* block_sums_out, int64_t* block_last_ids_out, bool return_next_level,
                                   int64_t num_blocks, int64_t threads_per_block, int64_t shared_memory_size);
void launch_segcumsum_kernel_double(double* values, const int64_t* segment_ids, int64_t size, int64_t max_seg_size,
                                    double* block_sums_out, int64_t* block_last_ids_out, bool return_next_level,
                                    int64_t num_blocks, int64_t threads_per_block, int64_t shared_memory_size);
void launch_add_block_sums_kernel_float(float* output, const float* block_sums, const int64_t* segment_ids,
                                        const int64_t* block_last_id, int64_t size, int64_t num_blocks,
                                        int64_t threads_per_block);
void launch_add_block_sums_kernel_double(double* output, const double* block_sums, const int64_t* segment_ids,
                                         const int64_t* block_last_id, int64_t size, int64_t num_blocks,
                                         int64_t threads_per_block);

/* ------------------------------------------------------------------------------------------------
 * 2. Kseg: single-pass segmented inclusive cumulative sum (decoupled look-back)
 *    Replaces the whole hierarchy loop of segcumsum_cuda (fsw_embedding.py:2878-3012).
 *    values_in/values_out may alias (in place).  id_bytes = 4 or 8.  A segment is a maximal run of
 *    equal consecutive ids (segcumsum_slow, fsw_embedding.py:3016-3027).
 * ---------------------------------------------------------------------------------------------- */
size_t fsw_segcumsum_workspace_bytes(int64_t n);
int fsw_segcumsum(int dtype, const void* values_in, void* values_out, const void* segment_ids, int id_bytes,
                  int64_t n, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 3. K0: graph preparation and segment plan
 * ---------------------------------------------------------------------------------------------- */
/* edge_index [2, E] int64 (row 0 = source, row 1 = destination; fsw_conv.py:387 flips it so that
 * rows of the adjacency are destinations) -> destination-major CSR.
 *   rowptr [N+1] int32, col [E'] int32 (source vertex), eid [E'] int32 (original edge id, or
 *   E + v for the self loop of vertex v), E' = E + (self_loops ? N : 0).
 * Duplicate (dst, src) pairs are kept as separate elements: the embedding of a multiset is
 * unchanged by splitting a weight over two copies of the same point, so this equals the
 * reference's `coalesce()` sum (fsw_conv.py:397-398) in value and in gradient.
 * workspace: fsw_csr_workspace_bytes(N, E).  The elements of a segment keep the order of the edge list (stable sort by
 * destination), so the CSR of a graph - and with it the order in which exact key ties are resolved - is reproducible. */
size_t fsw_csr_workspace_bytes(int64_t N, int64_t E);
int fsw_csr_from_edge_index(const int64_t* edge_index, int64_t E, int64_t N, int self_loops, int32_t* rowptr,
                            int32_t* col, int32_t* eid, void* workspace, size_t workspace_bytes, void* stream);

/* Coalescing variant (graphs with edge features: the reference's coalesce() sums the feature vectors of duplicate edges,
 * fsw_conv.py:397-398, :438-439): duplicate (dst, src) pairs become ONE element with the summed base weight (1 per edge,
 * self_loop_weight per self loop).  col / W have room for E (+ N) elements, the first *nslots (device int32) are written;
 * slot_of_elem [E (+ N)] = CSR slot of every input edge, then of every self loop; deg [N] = in-degrees (`dtype`);
 * gcn != 0: W = base / sqrt(deg[dst]) / sqrt(deg[src]).  Stable 64-bit radix sort: deterministic. */
size_t fsw_csr_coalesce_workspace_bytes(int64_t N, int64_t E);
int fsw_csr_coalesce(int dtype, const int64_t* edge_index, int64_t E, int64_t N, int self_loops, double self_loop_weight, int gcn,
                     int32_t* rowptr, int32_t* col, void* W, int32_t* slot_of_elem, void* deg, int32_t* nslots, void* workspace,
                     size_t workspace_bytes, void* stream);

/* rows [nnz] int64 sorted ascending (a coalesced COO tensor, fsw_embedding.py:664-668) -> rowptr [S+1] int32 */
int fsw_rowptr_from_sorted_rows(const int64_t* rows, int64_t nnz, int64_t S, int32_t* rowptr, void* stream);

/* Per-CSR-slot weights of FSW_conv.edge_index_to_adj (fsw_conv.py:388-409):
 *   base weight 1 (edges) / self_loop_weight (self loops); in-degree deg[v] = sum of base weights of
 *   row v (written as `dtype`);  gcn != 0: w = base / sqrt(deg[dst]) / sqrt(deg[src]).
 * w may be NULL when gcn == 0 and self_loops == 0 (unit weights need no array). */
int fsw_edge_weights(int dtype, const int32_t* rowptr, const int32_t* col, const int32_t* eid, int64_t N, int64_t E,
                     int self_loops, double self_loop_weight, int gcn, void* deg_out, void* w_out, void* stream);

/* Transpose of a CSR segment structure with explicit columns: for every point row j (0 <= j < Nrows) the
 * (segment, slot) pairs that reference it: tptr [Nrows+1], tseg [E], tslot [E], tn [E] = number of elements of
 * that segment when it is eligible for the source-major backward (uniform weights, n <= nmax_eligible), else 0.
 * `info` comes from fsw_segment_plan.  workspace: fsw_transpose_workspace_bytes(Nrows, E).  The pairs of a source row are
 * listed in element order (stable sort), so the source-major backward sums them in the same order in every run.
 * A transposition handed to fsw_embed_backward must be built with nmax_eligible = FSW_RANKT_ELIGIBLE(max n_eff of
 * the plan): the backward serves exactly those segments through it - every uniform segment of up to FSW_RANKT_NMAX
 * elements (the plan has a closed size bucket that ends there); hubs beyond are re-sorted and added with atomics. */
#define FSW_RANKT_NMAX 32768
#define FSW_RANKT_ELIGIBLE(max_n_eff) (FSW_RANKT_NMAX)
size_t fsw_transpose_workspace_bytes(int64_t Nrows, int64_t E);
int fsw_csr_transpose(const int32_t* rowptr, const int32_t* col, const int32_t* info, int64_t S, int64_t Nrows, int64_t E,
                      int nmax_eligible, int32_t* tptr, int32_t* tseg, int32_t* tslot, int32_t* tn, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Segment statistics + plan.
 *   rowptr [S+1] (or NULL: S segments of n_fixed elements each), W [E] raw weights or NULL (unit).
 *   mass [S] float64 = total mass T_s (fsw_embedding.py:778-784);
 *   info [S] int32   = n_eff | (uniform << 30), n_eff = n + (T < thresh) (deficit pad, :787-815),
 *                      uniform = all weights of the segment equal and T >= thresh;
 *   order [S] int32  = segments sorted by plan bucket (kind * FSW_PLAN_BUCKETS_PER_KIND + size bucket);
 *   bucket_offsets [FSW_PLAN_BUCKETS + 2] int32 (device) = start of each bucket inside `order`,
 *                      then S, then max n_eff;
 *   bucket_elems [FSW_PLAN_BUCKETS] int64 (device, may be NULL) = sum of n_eff over each bucket
 *                      (used by the benchmark to turn per-class kernel times into bytes/s).
 * workspace: fsw_plan_workspace_bytes(S). */
size_t fsw_plan_workspace_bytes(int64_t S);
int fsw_segment_plan(int dtype, const int32_t* rowptr, int64_t n_fixed, const void* W, int64_t S, double thresh,
                     double* mass, int32_t* info, int32_t* order, int32_t* bucket_offsets, int64_t* bucket_elems,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4. K1: dense contractions at full input precision (fp32 FMA / fp64), row-major operands
 *    C[M, N] (ldc) (+)= A . B with
 *      op = 0 (NT): A [M, Kd] (lda), B [N, Kd] (ldb)   -> projection Xp = X . theta^T (fsw_embedding.py:911)
 *      op = 1 (NN): A [M, Kd] (lda), B [Kd, N] (ldb)   -> dX = dXp . theta
 *      op = 2 (TN): A [Kd, M] (lda), B [Kd, N] (ldb)   -> dtheta = dXp^T . X   (split over Kd, atomics)
 *    accumulate != 0 adds into C (C must be initialised); op 2 always accumulates.
 * ---------------------------------------------------------------------------------------------- */
int fsw_gemm(int dtype, int op, int64_t M, int64_t N, int64_t Kd, const void* A, int64_t lda, const void* B,
             int64_t ldb, void* C, int64_t ldc, int accumulate, void* stream);
/* Large fp32 contractions (M >= 2048 rows, or a reduction over >= 4096 rows for op 2; 16-byte aligned operands, leading
 * dimensions multiples of 4) run on the tensor cores: tcgen05.mma kind::tf32 on TMA-staged tiles with a hi/lo split of
 * both operands (3 products per k-step, two TMEM accumulators) that keeps fp32 accuracy - csrc/fsw_umma.cu.  Everything
 * else, and fp64, runs the FMA kernels.  fsw_set_tensor_cores(0) forces the FMA kernels (tests, A/B timing).
 *
 * fsw_gemm_fused: C[M, N] (+)= sum_{s < nseg} A_s . B_s^T (+ bias[N]), nseg in {1, 2}: the contraction axis is the
 * concatenation of the segments, so FSW_conv's cat(emb, x) . W^T (fsw_conv.py:357-361) needs no concatenated copy:
 * A_0 = emb [M, Kd_0], A_1 = x [M, Kd_1], B_0 = W[:, :Kd_0], B_1 = W[:, Kd_0:] (views, ldb = row pitch of W).  fp32. */
int fsw_set_tensor_cores(int on);
int fsw_gemm_fused(int dtype, int64_t M, int64_t N, int nseg, const int64_t* Kd, const void* const* A, const int64_t* lda,
                   const void* const* B, const int64_t* ldb, void* C, int64_t ldc, const void* bias, int accumulate,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * 5. K2: fused gather -> per-(segment, slice) sort -> cumulative weights -> Fourier -> reduce
 *    (fsw_embedding.py:917-1109 and every ag/sp forward it calls)
 *
 *  Xp   [Nrows, ldp]   projected points (K1); element e of segment s reads row col[e] (or e if col==NULL)
 *  Ep   [E, ldp] or NULL  per-element additive projection (edge features, fsw_embedding.py:934-947)
 *  rowptr/n_fixed/col/W/mass/info/order/bucket_offsets_host: the segment plan (section 3);
 *       bucket_offsets_host is the HOST copy of bucket_offsets (FSW_PLAN_BUCKETS + 1 ints)
 *  freqs [K]
 *  out  [S, ld_out]: out[s, out_col0 + k] = (1 + xi_k) * sum_j p_(j) D_j  (+ bias[k] if bias != NULL)
 *  scratch: only needed when a segment does not fit shared memory; size from fsw_embed_scratch_bytes.
 *  ranks_out [E, ldr] uint16 or NULL: when given (training), the forward also records the sorted position
 *       of every element per slice for the uniform-weight segments of up to 512 elements - the analogue of
 *       the compressed permutation the reference saves for its backward (fsw_embedding.py:2041-2050).
 *  dxi_out [S, ld_dxi] or NULL (only with ranks_out): d out[s, k] / d xi_k for the same segments; the caller
 *       zero-initialises it and forms dL/dxi_k = sum_s g[s, k] * dxi_out[s, k] (+ what fsw_embed_backward adds).
 * ---------------------------------------------------------------------------------------------- */
size_t fsw_embed_scratch_bytes(int dtype, const int32_t* bucket_offsets_host, int64_t K, int64_t max_n_eff,
                               int backward);
/* extra bytes to append to the BACKWARD scratch when the transposed structure is passed (pre-scaled gradient) */
size_t fsw_embed_backward_extra_bytes(int dtype, int64_t S, int64_t K);
int fsw_embed_forward(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr,
                      int64_t n_fixed, const int32_t* col, const void* W, const double* mass, const int32_t* info,
                      const int32_t* order, const int32_t* bucket_offsets_host, int64_t S, int64_t K,
                      const void* freqs, double thresh, void* out, int64_t ld_out, int64_t out_col0,
                      const void* bias, int64_t max_n_eff, void* scratch, size_t scratch_bytes, void* ranks_out,
                      int64_t ldr, void* dxi_out, int64_t ld_dxi, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 6. K3: fused backward of section 5 (ag.*.backward fsw_embedding.py:1286-2258)
 *  g    [S, ld_g]: upstream gradient of out[:, g_col0 : g_col0 + K]
 *  dXp  [Nrows, ldp]: += G_k D_rank scattered to the element's row (atomics when col != NULL;
 *        must be zero-initialised by the caller)
 *  dEp  [E, ldp] or NULL: same value per element (no atomics)
 *  dfreqs_acc [K] float64 or NULL: += dL/dxi_k (atomics; caller zero-initialises)
 *  dW   must be NULL: the gradient w.r.t. the weights has its own entry points (section 6b)
 *  ranks [E, ldr] uint16 or NULL: positions recorded by fsw_embed_forward; with them the backward of the
 *        covered segments is a streaming pass without any sort, otherwise everything is re-sorted.
 *  dxi_from_forward != 0: the forward was given dxi_out, so the covered segments skip the frequency gradient.
 *  tptr/tseg/tslot/tn (fsw_csr_transpose) or NULL: with them (fp32 graphs, ranks, dxi_from_forward) the segments
 *        of up to FSW_RANKT_ELIGIBLE(max_n_eff) elements run SOURCE-major: every row of dXp is written once with a plain
 *        store instead of scattered atomics.  dXp [nrows, ldp] then need NOT be initialised by the caller (the library
 *        clears it itself in the cases where that kernel does not run, e.g. an empty shard).
 * ---------------------------------------------------------------------------------------------- */
int fsw_embed_backward(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr,
                       int64_t n_fixed, const int32_t* col, const void* W, const double* mass,
                       const int32_t* info, const int32_t* order, const int32_t* bucket_offsets_host, int64_t S,
                       int64_t K, const void* freqs, double thresh, const void* g, int64_t ld_g, int64_t g_col0,
                       void* dXp, void* dEp, double* dfreqs_acc, void* dW, int64_t max_n_eff, void* scratch,
                       size_t scratch_bytes, const void* ranks, int64_t ldr, int dxi_from_forward, const int32_t* tptr,
                       const int32_t* tseg, const int32_t* tslot, const int32_t* tn, int64_t nrows, void* stream);

/* 5b. Point-cloud mode (BASELINE.json configs[2]; reference: the dense path of forward_helper, fsw_embedding.py:925, :989-1004):
 * dense batch of S unit-weight multisets of n points each (33 <= n <= 1024, n >= thresh), d <= 4, fp32.  The keys
 * <x_e, theta_k> are formed inside the sort kernel (no projected matrix), the ranks are recorded slice-major
 * ranksT [S][K][n] uint16, and the backward evaluates dL/dp where it is consumed (no projected gradient):
 *   fsw_embed_forward_cloud:  X [S*n, d], theta [K, ldt]; mass/info/order/bucket_offsets_host: the plan of the dense batch;
 *                             scratch: fsw_embed_scratch_bytes of that plan; ranksT_out / dxi_out may be NULL (inference).
 *   fsw_embed_backward_cloud: dX [S*n, d] overwritten, dtheta [K, ld_dt] added to (either may be NULL). */
int fsw_embed_forward_cloud(int dtype, const void* X, int64_t d, const void* theta, int64_t ldt, int64_t n, const double* mass,
                            const int32_t* info, const int32_t* order, const int32_t* bucket_offsets_host, int64_t S, int64_t K,
                            const void* freqs, double thresh, void* out, int64_t ld_out, int64_t out_col0, const void* bias,
                            void* scratch, size_t scratch_bytes, void* ranksT_out, void* dxi_out, int64_t ld_dxi, void* stream);
int fsw_embed_backward_cloud(int dtype, const void* X, int64_t d, const void* theta, int64_t ldt, int64_t n, int64_t S, int64_t K,
                             const void* freqs, const void* g, int64_t ld_g, int64_t g_col0, const void* ranksT, void* dX,
                             void* dtheta, int64_t ld_dt, void* stream);


/* ------------------------------------------------------------------------------------------------
 * 6b. K3w: gradient with respect to the WEIGHTS of the multisets (ag.cumsum_sparse.backward fsw_embedding.py:2160-2172,
 *     ag.div_sparse_dense :1656, deficit padding with custom_lowclamp :787-829, :1735-1744)
 *  fsw_embed_backward_weights accumulates d L / d w (NORMALISED weights) of the slices [0, K) of this call into
 *     dwn_acc [E] float64 and, for the pad element of a segment, dwn_pad_acc [S] float64 (both zero-initialised by the
 *     caller; several calls - column chunks - may accumulate into the same buffers);
 *  fsw_embed_weight_grad_finish pushes them through the normalisation and writes dW [E] (raw weights, `dtype`).
 *  any_deficient != 0: some segment of the batch has total mass < thresh (the reference then pads every row, which
 *     changes the gradient of the segments whose mass equals thresh exactly).
 *  max_n: the largest number of elements of a segment.  scratch: fsw_embed_weight_grad_scratch_bytes(dtype, max_n)
 *     (0 when every tile fits shared memory).  Xp / Ep / rowptr / n_fixed / col / W / mass / freqs / g as in section 6.
 * ---------------------------------------------------------------------------------------------- */
/* acc[k] += sum_s g[s, k] * d[s, k] for k < K, float64 accumulators: the frequency gradient of the segments whose d out / d xi
 * the forward wrote (dxi_out of fsw_embed_forward; replaces the reference's autograd through the frequency multiply,
 * fsw_embedding.py:1037-1045), in one pass over the two [S, K] matrices. */
int fsw_column_dot(int dtype, const void* g, int64_t ld_g, const void* d, int64_t ld_d, int64_t S, int64_t K, double* acc,
                   void* stream);
size_t fsw_embed_weight_grad_scratch_bytes(int dtype, int64_t max_n);
int fsw_embed_backward_weights(int dtype, const void* Xp, int64_t ldp, const void* Ep, const int32_t* rowptr, int64_t n_fixed,
                               const int32_t* col, const void* W, const double* mass, int64_t S, int64_t K, const void* freqs,
                               double thresh, int any_deficient, const void* g, int64_t ld_g, int64_t g_col0, double* dwn_acc,
                               double* dwn_pad_acc, int64_t max_n, void* scratch, size_t scratch_bytes, void* stream);
int fsw_embed_weight_grad_finish(int dtype, const int32_t* rowptr, int64_t n_fixed, const void* W, const double* mass, int64_t S,
                                 double thresh, int any_deficient, const double* dwn_acc, const double* dwn_pad_acc, void* dW,
                                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * 7. Counters and per-kernel timers (the reference has only unused wall-clock globals,
 *    fsw_embedding.py:118-119, :1150-1160)
 * ---------------------------------------------------------------------------------------------- */
/* number of kernel launches issued by this library since load */
int64_t fsw_launch_count(void);
/* When enabled, every launch of the embed / gemm kernels is bracketed by CUDA events on the launching
 * stream.  fsw_profile_read synchronises those events and writes one line per label
 * "label count total_ms\n" into buf (returns the number of bytes needed), then clears the records. */
int fsw_profile_enable(int on);
int64_t fsw_profile_read(char* buf, int64_t buf_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FSW_EMBEDDING_H */
