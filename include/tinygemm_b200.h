/*
 * tinygemm_b200 - C ABI of the B200-native (sm_100a) tinygemm replacement.
 *
 * This header is the drop-in boundary of the repo: plain device pointers, sizes and a
 * cudaStream_t (passed as void*) - no torch types.  Every entry point below replaces
 * one operator of the reference's torch extension (reference = facebookresearch/any4,
 * directory tinygemm_lib/; schemas at TinyGemm.cpp:19-121, declarations with the layout
 * comments at TinyGemm.h:19-216).  The torch custom-op layer that re-creates
 * `torch.ops.tinygemm.*` on top of this ABI lives in any4_b200/csrc/torch_ops.cpp.
 *
 * Conventions
 *   - all pointers are device pointers on the current CUDA device unless stated otherwise
 *   - all tensors are dense / contiguous in the layouts written next to each function
 *   - every call is asynchronous on `stream` and never synchronises the device
 *   - return value: TG_OK (0) or a negative TG_ERR_* code; tg_last_error() returns a
 *     human-readable message for the last failure on the calling thread
 *   - there is NO CPU fallback: a call on a machine without a usable sm_100 device
 *     fails with TG_ERR_CUDA
 */
#ifndef TINYGEMM_B200_H
#define TINYGEMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_OK 0
#define TG_ERR_INVALID_ARGUMENT (-1) /* shape / enum / alignment check failed (reference: TORCH_CHECK) */
#define TG_ERR_CUDA (-2)             /* CUDA runtime error at launch */
#define TG_ERR_UNSUPPORTED (-3)      /* combination the reference rejects as well */

/* activation / output element type ("f16" in the reference op names) */
typedef enum { TG_BF16 = 0, TG_FP16 = 1 } tg_dtype;

/* 4-bit weight formats; reference: Int4_QType, TinyGemmUtils.cuh:21-32 */
typedef enum {
  TG_W4_INT4 = 0,         /* uniform int4: w = (code - 8) * scale + zero                */
  TG_W4_ANY4_GLOBAL = 1,  /* one 16-entry LUT for the matrix (nf4/fp4/af4 run this way)  */
  TG_W4_ANY4_ROWWISE = 2, /* one 16-entry LUT per (padded) weight row                    */
  TG_W4_MX4 = 3           /* fp4 e2m1 codes + one e8m0 exponent per (row, group)         */
} tg_w4_format;

/* which mma operand the packed weight was laid out for (reference: `weightOnRight`) */
typedef enum {
  TG_WEIGHT_B = 1, /* weightOnRight = true : y = x W^T, W in "B" layout (8-row tiles)  */
  TG_WEIGHT_A = 0  /* weightOnRight = false: y = (W x^T)^T, W in "A" layout (16-row tiles) */
} tg_weight_side;

/* process-wide tuning options (no counterpart in the reference; defaults reproduce its semantics exactly) */
typedef enum {
  /* launch the GEMV kernels with programmatic stream serialization so that a kernel's prologue overlaps the
   * tail of the previous kernel in the stream (default 1).  Ordering is preserved: the kernel waits for the
   * previous one before it reads activations / writes its output. */
  TG_OPT_PDL = 0,
  /* promise that packed weights, LUTs, scales/zeros and exponents passed to the GEMV entry points are never
   * produced by the kernel launched just before them on the same stream (true for a loaded model; NOT true for
   * convert-then-GEMM sequences such as functional.linear_*(reshape_weight=True)).  With the promise (and PDL)
   * the weight stream starts before the previous kernel has finished.  Default 0. */
  TG_OPT_STATIC_WEIGHTS = 1,
  /* B-layout 4-bit GEMM kernel choice: 0 = automatic (default), 1 = always the tcgen05 / TMEM kernel (one pass for up
   * to 16 activation rows), 2 = the lane-per-row mma.sync decode kernel wherever it applies (<= 4 rows per pass).  All
   * of them produce the same dequantised weights and exact products; only the order of the fp32 partial sums differs. */
  TG_OPT_W4_KERNEL = 2
} tg_option;
int tg_set_option(tg_option option, int value);

const char* tg_last_error(void);
/* library / build identification, e.g. "tinygemm_b200 0.1 sm_100a" */
const char* tg_version(void);
/* number of kernels launched by this library on the calling thread since the last reset
 * (used by bench.py for its `gpu_launches` claim) */
uint64_t tg_launch_count(void);
void tg_reset_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Layout conversion ("convert_matrix_*" ops).  Bit-exact with the reference kernels.
 * Lane geometry: lane t of a warp, g = t / 4, q = t % 4, k0 = kTile * 16 + 2 * q.
 * ---------------------------------------------------------------------------------- */

/* [m][k] 16-bit -> [ceil(m/16)][ceil(k/16)][32][8], zero padded.
 * replaces convert_matrix_to_m16n8k16_A_layout (TinyGemmConvertA.cu:150-223; kernel :19-141) */
int tg_convert_to_A(const void* in, void* out, int64_t m, int64_t k, void* stream);

/* inverse of the above. replaces convert_matrix_from_m16n8k16_A_layout
 * (TinyGemmConvertA.cu:554-626; kernel :442-546) */
int tg_convert_from_A(const void* in, void* out, int64_t m, int64_t k, void* stream);

/* [n][k] 16-bit -> [ceil(n/8)][ceil(k/(ik*16))][32][ik*4], ik in {1,2}.
 * replaces convert_matrix_to_m16n8k16_B_layout (TinyGemmConvertB.cu:76-133; kernel :20-66) */
int tg_convert_to_B(const void* in, void* out, int64_t n, int64_t k, int inner_k_tiles, void* stream);

/* inverse. replaces convert_matrix_from_m16n8k16_B_layout (TinyGemmConvertB.cu:186-249; kernel :136-176) */
int tg_convert_from_B(const void* in, void* out, int64_t n, int64_t k, int inner_k_tiles, void* stream);

/* [m][k] int32 codes (0..15) -> [ceil(m/16)][ceil(k/(ik*16))][32][ik] int32, ik in {1,2,4}.
 * replaces convert_matrix_to_m16n8k16_Aint4_layout (TinyGemmConvertA.cu:289-333; kernel :226-285) */
int tg_convert_to_Aint4(const int32_t* in, int32_t* out, int64_t m, int64_t k, int inner_k_tiles, void* stream);

/* [m][k] int32 codes (0..255) -> [ceil(m/16)][ceil(ceil(k/16)/ik)][32][2*ik] int32, ik in {1,2}.
 * replaces convert_matrix_to_m16n8k16_Aint8_layout (TinyGemmConvertA.cu:400-440; kernel :340-396) */
int tg_convert_to_Aint8(const int32_t* in, int32_t* out, int64_t m, int64_t k, int inner_k_tiles, void* stream);

/* [n][k] int32 codes (0..15) -> [ceil(n/8)][k/(ik*16)][32][ik/2] int32, ik in {2,4,8}, k % (ik*16) == 0.
 * replaces convert_matrix_to_m16n8k16_Bint4_layout (TinyGemmConvertB.cu:312-364; kernel :252-308) */
int tg_convert_to_Bint4(const int32_t* in, int32_t* out, int64_t n, int64_t k, int inner_k_tiles, void* stream);

/* [n][k] int32 codes (0..255) -> [ceil(n/8)][k/(ik*16)][32][ik] int32, ik in {1,2,4}, k % (ik*16) == 0.
 * replaces convert_matrix_to_m16n8k16_Bint8_layout (TinyGemmConvertB.cu:415-464; kernel :367-410) */
int tg_convert_to_Bint8(const int32_t* in, int32_t* out, int64_t n, int64_t k, int inner_k_tiles, void* stream);

/* ------------------------------------------------------------------------------------
 * Weight-only small-batch GEMM, row-major activations and output
 * ("tinygemm_y_f16RM_x_f16RM_w_*TC" ops).
 *
 *   x        [rows_x][k]            dtype           (any rows_x >= 1; tuned for <= 16)
 *   y        [rows_x][w_rows]       dtype           (w_rows = padded weight rows, see below)
 *   w        packed weight in the A- or B- tensor-core layout produced by tg_convert_to_*
 *   w_rows   PADDED number of weight rows: 8 * nTiles (TG_WEIGHT_B) or 16 * mTiles (TG_WEIGHT_A)
 *   k        reduction length, k % 32 == 0 (reference: TinyGemmImpl.cuh:372-376)
 *
 * Numerics contract (reference: Dequantization.cuh, MatrixLayoutB.cuh:1005-1088,
 * FloatDefs.cuh:87-119): every weight is v = LUT[code] in `dtype`, then
 * fma.rn(v, scale, zero) with ONE rounding to `dtype` (mx4: v * 2^(e-127), e==255 -> NaN);
 * products with x are exact and accumulated in fp32; one round-to-nearest-even at the end.
 * ---------------------------------------------------------------------------------- */

/* int4 / any4 / mx4.  replaces tinygemm_y_f16RM_x_f16RM_w_{int4,any4,mx4}TC
 * (TinyGemm_int4.cu:294-548 dispatch, :550-794 public ops).
 *   scales_zeros  [k/group][w_rows][2] dtype       (int4 / any4; ignored for mx4)
 *   lut           [16] (global) or [w_rows][16] dtype (any4 only)
 *   exponents     [w_rows][k/group] uint8 e8m0     (mx4 only; dtype must be TG_BF16)
 *   group         32, 64, 128 or 256
 *   inner_k_tiles B layout: 2, 4, 8;  A layout: 1, 2, 4
 * Concurrency: calls on ONE stream (the normal case, CUDA-graph capture and replay included) are always safe.  The
 * B-layout tensor-core kernel keeps two kinds of device-global scratch, each rotated over 4 pools per launch - the
 * tagged partial sums of row blocks shared by several CTAs and, from 5 activation rows on, the permuted activations:
 * launches that may run CONCURRENTLY on different streams are safe as long as no more than 4 of them are in flight. */
int tg_gemm_w4_rm(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                  const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                  int inner_k_tiles, tg_w4_format format, tg_weight_side side, tg_dtype dtype, void* stream);

/* tg_gemm_w4_rm with scratch memory.  The A-layout kernel takes ONE activation row per launch; with a workspace of
 * tg_gemm_w4_rm_workspace_bytes(...) bytes (0 = no use for one) several rows against an A-layout weight are computed by
 * repacking the weight into the B layout (tg_repack_Aint4_to_Bint4, the same nibbles in another order) and running the
 * one-pass tcgen05 kernel on it.  Same results as the B-layout op on the B-packed weight.  The workspace may be reused
 * as soon as the call's work on `stream` has completed.  Any other case: identical to tg_gemm_w4_rm. */
size_t tg_gemm_w4_rm_workspace_bytes(int64_t rows_x, int64_t w_rows, int64_t k, tg_weight_side side);
int tg_gemm_w4_rm_ws(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                     const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group, int inner_k_tiles,
                     tg_w4_format format, tg_weight_side side, tg_dtype dtype, void* workspace, size_t workspace_bytes,
                     void* stream);
/* packed A int4 layout [rows/16][ceil(k/16/ik_a)][32][ik_a] -> packed B int4 layout [rows/8][k/(16 ik_b)][32][ik_b/2] of the
 * same code matrix (rows = the padded A row count; k % (16 ik_b) == 0) */
int tg_repack_Aint4_to_Bint4(const int32_t* in, int32_t* out, int64_t rows, int64_t k, int ik_a, int ik_b, void* stream);

/* The same GEMM for a decode step whose activations live in PINNED HOST memory and whose result is wanted there
 * (no counterpart in the reference, whose ops take device tensors): one tiny kernel pulls `x_host` (device-accessible
 * pinned memory, unified addressing) into the device buffer `x_staging` [rows_x][k], the GEMV - ordered behind it by
 * programmatic dependent launch, its weight stream already running - writes its outputs straight to `y_host`.  Two
 * kernel launches from ONE call, no copy-engine transfers.  (x is staged because every CTA reads all of it.)
 * `x_host` is an input the HOST wrote before the call: with TG_OPT_PDL the staging kernel fetches it while the previous
 * kernel of the stream may still be running (only `x_staging` and `y_host` are ordered behind that kernel), so it must
 * not be the output of earlier device work on the same stream - stage such data with a device-side copy instead. */
int tg_gemm_w4_rm_hostio(void* y_host, const void* x_host, void* x_staging, const int32_t* w, const void* scales_zeros,
                         const void* lut, const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                         int inner_k_tiles, tg_w4_format format, tg_weight_side side, tg_dtype dtype, void* stream);

/* Row-sharded multi-GPU variant of the B-layout 4-bit GEMV with the exchange FUSED into the epilogue (no
 * counterpart in the reference, which is single-GPU).  This rank holds `w_rows` consecutive weight
 * rows of an n-row layer; `y_peers[r]` is the address, in rank r's copy of a symmetric (peer-mapped) m x n output
 * buffer, of THIS shard's first column, and `y_row_stride` = n.  The kernel's epilogue stores the shard's outputs
 * straight into every rank's buffer over NVLink - there is no separate all-gather / all-reduce kernel.  The caller
 * orders the consumers of the buffer behind all ranks' stores (e.g. a symmetric-memory barrier on the stream) and
 * must not reuse a buffer while a peer may still read it (alternate two buffers).  n_peers <= 8. */
int tg_gemm_w4_rm_sharded(void* const* y_peers, int n_peers, int64_t y_row_stride, const void* x, const int32_t* w,
                          const void* scales_zeros, const void* lut, const uint8_t* exponents, int64_t rows_x,
                          int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_w4_format format, tg_dtype dtype,
                          void* stream);

/* Row-sharded GEMV with the exchange INSIDE the kernel (no collective, no barrier, no flag).
 *   y            this rank's plain output [rows_x][y_row_stride]; on completion columns [0, n_peers * w_rows) hold the
 *                FULL output (every rank's shard), 4-byte aligned, even row stride
 *   xchg_peers   xchg_peers[r] = address, in rank r's memory, of a symmetric exchange buffer of 8-byte words
 *                [rows_x][n_peers * w_rows / 2] (the same buffer for all ranks of one call; zero before its first use)
 *   tag          non-zero.  A consumed word is cleared, so a call may reuse the tag of the buffer's previous use (a CUDA
 *                graph replay re-issues every call with the tag it was captured with)
 * The epilogue stores the shard as words (tag << 32 | two adjacent outputs) into EVERY rank's exchange buffer with one
 * 8-byte store each - whoever sees the tag sees the values - and every CTA, before it exits, collects its slice of all
 * ranks' words from the local buffer (spinning on the tag) into `y`.  When the kernel has completed on a rank the full
 * output is in that rank's `y`: consumers are ordered by plain stream order, programmatic dependent launch chains stay
 * intact.  All ranks must issue the same sequence of calls; at least two exchange buffers must alternate.
 * Values are bit-identical to the single-GPU output of the same kernel.  rows_x: any (4 per pass). */
int tg_gemm_w4_rm_exchange(void* y, void* const* xchg_peers, int self_rank, uint32_t tag, int n_peers, int64_t y_row_stride,
                           const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                           const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                           int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream);

/* tg_gemm_w4_rm_exchange over a shard of ROW-INTERLEAVED (gate, up) weights (as tg_gemm_w4_rm_silu_pairs): the epilogue
 * applies silu(gate) * up and exchanges the w_rows / 2 activated outputs of the shard; `y` receives the full
 * [rows_x][n_peers * w_rows / 2] result, the exchange buffers hold [rows_x][n_peers * w_rows / 4] words.  w_rows % 4 == 0. */
int tg_gemm_w4_rm_exchange_silu_pairs(void* y, void* const* xchg_peers, int self_rank, uint32_t tag, int n_peers,
                                      int64_t y_row_stride, const void* x, const int32_t* w, const void* scales_zeros,
                                      const void* lut, const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k,
                                      int group, int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream);

/* int8.  replaces tinygemm_y_f16RM_x_f16RM_w_int8TC (TinyGemm_int8.cu:215-399, :430-457).
 *   inner_k_tiles B layout: 1, 2, 4;  A layout: 1, 2 */
int tg_gemm_w8_rm(void* y, const void* x, const int32_t* w, const void* scales_zeros, int64_t rows_x,
                  int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_weight_side side,
                  tg_dtype dtype, void* stream);

/* 16-bit weights.  replaces tinygemm_y_f16RM_x_f16RM_w_f16TC (TinyGemm_bf16.cu:163-293, :312-327).
 *   w in the 16-bit A layout (ik 1) or B layout (ik 1, 2); k_padded = 16 * kTiles of the weight */
int tg_gemm_w16_rm(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k,
                   int inner_k_tiles, tg_weight_side side, tg_dtype dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Tensor-core-layout activations and output ("tinygemm_y_f16TC_x_f16TC_w_*TC" ops,
 * TinyGemm_int4.cu:28-292, TinyGemm_int8.cu:22-213, TinyGemm_bf16.cu:35-150).
 *
 * TG_WEIGHT_B: x is in the 16-bit A layout [mT][kT][32][8]; y is written in the 16-bit A
 *              layout [mT][ceil(nT/2)][32][8] (n plays the role of k).
 * TG_WEIGHT_A: x is in the 16-bit B layout [nT][ceil(kT/x_ik)][32][x_ik*4]; y is written in
 *              the B layout [nT][ceil(mT/x_ik)][32][x_ik*4] (weight rows play the role of k).
 * rows_x here is the PADDED activation row count (16*mT or 8*nT).  `workspace` must hold
 * tg_gemm_tc_workspace_bytes(rows_x, w_rows, k) bytes of device memory; it is used for the
 * row-major staging of x and y and may be reused as soon as the call's work on `stream`
 * has completed.
 * ---------------------------------------------------------------------------------- */
size_t tg_gemm_tc_workspace_bytes(int64_t rows_x, int64_t w_rows, int64_t k);

int tg_gemm_w4_tc(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                  const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                  int inner_k_tiles, int x_inner_k_tiles, tg_w4_format format, tg_weight_side side,
                  tg_dtype dtype, void* workspace, void* stream);

int tg_gemm_w8_tc(void* y, const void* x, const int32_t* w, const void* scales_zeros, int64_t rows_x,
                  int64_t w_rows, int64_t k, int group, int inner_k_tiles, int x_inner_k_tiles,
                  tg_weight_side side, tg_dtype dtype, void* workspace, void* stream);

int tg_gemm_w16_tc(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k,
                   int inner_k_tiles, int x_inner_k_tiles, tg_weight_side side, tg_dtype dtype,
                   void* workspace, void* stream);

/* debug op: 8 x int4 -> 8 x bf16 per int32 word, value = code - 8, element order
 * v0 v1 ... v7 of the packed word.  replaces tinygemm_dequant_int4 (TinyGemmDequantize.cu:19-58).
 *   in [n_words] int32 -> out [n_words][8] bf16 */
int tg_dequant_int4(const int32_t* in, void* out, int64_t n_words, void* stream);

/* ------------------------------------------------------------------------------------
 * any4 quantizer front-end: one kernel, one CTA per weight row.  Replaces the reference's CPU
 * pipeline group_q (quantize.py:106-149) -> cluster_matrix / cluster_row (quantize.py:433-521) -> kmeans.run_kmeans
 * (kmeans.py:200-262, init "int" = kmeans.py:41-46) -> lut = any4 - 8 (quantize.py:893) ->
 * convert_matrix_to_m16n8k16_Bint4_layout, for 4 bit, per-row LUT, asymmetric groups with zero point.
 *   w              [n][k] dtype, row-major              sample_weight  [k] fp32 or NULL (weighted k-means)
 *   codes          [n][k] int32 or NULL                 packed         B int4 layout [n/8][k/(ik*16)][32][ik/2] or NULL
 *   scales_zeros   [k/group][n][2] dtype                any4           [n][16] dtype, centroids in [0, 15] code space
 *   lut            [n][16] dtype = any4 - 8 (what the GEMV takes)       iters  [n] int32 Lloyd iterations run, or NULL
 * Group statistics and the stored scale / zero are the reference's fp32 operations (bit-identical); Lloyd's sums are
 * fp32 in a fixed order (deterministic, not numpy's: a value within an ulp of a boundary may take the other code). */
int tg_quantize_any4_rows(const void* w, const float* sample_weight, int64_t n, int64_t k, int group, int inner_k_tiles,
                          int max_iter, float tol, int32_t* codes, int32_t* packed, void* scales_zeros, void* any4, void* lut,
                          int32_t* iters, tg_dtype dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Decode-step plumbing around the GEMVs (no counterpart in tinygemm_lib - the
 * reference leaves these to the HF model code its benchmark.py:145-146 times as a whole).  Single token,
 * dtype = activation dtype, all kernels use programmatic dependent launch like the GEMVs (TG_OPT_PDL).
 * ------------------------------------------------------------------------------------ */
/* h[n] <- h + delta (delta may be NULL), rounded to dtype;  out[n] <- rmsnorm(h, eps) * weight[n]
 * (fp32 statistics, one rounding).  n % 8 == 0, n <= 8192. */
int tg_decode_add_rmsnorm(void* h, const void* delta, const void* weight, void* out, int64_t n, float eps,
                          tg_dtype dtype, void* stream);
/* out[n] <- silu(gate_up[0..n)) * gate_up[n..2n)  (output of a fused gate|up GEMV).  n % 8 == 0. */
int tg_decode_silu_mul(const void* gate_up, void* out, int64_t n, tg_dtype dtype, void* stream);
/* tg_gemm_w4_rm (weight on the right, B layout) with the activation fused into the epilogue: the weight holds
 * gate and up projections ROW-INTERLEAVED (row 2j = gate_j, row 2j+1 = up_j; w_rows even, padded rows come in
 * pairs) and y [rows_x][w_rows/2] = silu(gate) * up.  Every intermediate is rounded exactly as by tg_gemm_w4_rm
 * followed by tg_decode_silu_mul, so the result equals that two-kernel sequence. */
int tg_gemm_w4_rm_silu_pairs(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                             const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                             int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream);
/* qkv = [n_heads*128 | n_kv_heads*128 | n_kv_heads*128] (output of a fused q|k|v GEMV): rotary embedding
 * (half-rotation, cos/sin [128]) of q and k, append k, v to the caches [n_kv_heads][cache_len][128] at `pos`,
 * attention of the token over positions 0..pos with GQA, out [n_heads*128].  head_dim must be 128, pos <= 512.
 * With TG_OPT_PDL the kernel reads cos / sin and the cache rows of positions < pos BEFORE the previous kernel of the
 * stream has finished (they were written by earlier tokens; only qkv is that kernel's output): do not launch it
 * directly behind a kernel that writes those cache rows or the tables.  Likewise tg_decode_add_rmsnorm reads `weight`
 * before the previous kernel has finished. */
int tg_decode_rope_attention(const void* qkv, const void* cos, const void* sin, void* k_cache, void* v_cache,
                             void* out, int n_heads, int n_kv_heads, int head_dim, int pos, int cache_len,
                             float scale, tg_dtype dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYGEMM_B200_H */
