// extern "C" entry points of libtinygemm_b200.so: argument validation (the reference's
// TORCH_CHECKs restated on plain sizes), format dispatch, error strings.
// Declarations and the reference interface each one replaces: include/tinygemm_b200.h.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace tg {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += (uint64_t)n; }

void set_w4_options(int pdl, int static_weights);
// B-layout 4-bit GEMM kernel choice (tg_set_option(TG_OPT_W4_KERNEL, v)): 0 = automatic, 1 = always the tcgen05 kernel
// (gemv_w4_tc.cu), 2 = the lane-per-row mma.sync kernel (gemv_w4_b.cu) wherever it applies (<= 4 activation rows per pass).
int g_w4_kernel = 0;
int launch_gemm_w4_rm_B(void* y, const void* x, const int32_t* w, const void* sz, const void* lut,
                        const uint8_t* exps, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik,
                        tg_w4_format fmt, tg_dtype dt, const uint16_t* const_lut, cudaStream_t st,
                        void* const* y_peers = nullptr, int n_peers = 0, int64_t y_row_stride = 0, int silu_pairs = 0);
int launch_gemm_w4_tc_B(void* y, const void* x, const int32_t* w, const void* sz, const void* lut, const uint8_t* exps,
                        int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik, tg_w4_format fmt, tg_dtype dt,
                        const uint16_t* const_lut, cudaStream_t st, void* const* y_peers = nullptr, int n_peers = 0,
                        int64_t y_row_stride = 0, int silu_pairs = 0, int self_rank = 0, uint32_t exchange_tag = 0);
void set_tc_ctas_per_sm(int v);
// Automatic choice, from measurements on a B200 (profiles/r2/kernel_choice.md): the tcgen05 kernel handles 16 rows per
// pass and wins wherever a CTA has more than a handful of ring stages to amortise its prologue; for a decode GEMV so
// small that every SM gets at most one 32-row block of <= 4096 k (4096 x 4096: 64 KB per SM), the mma.sync kernel's 16
// consumer warps per CTA finish the block faster than the tcgen05 kernel's 8 dequant warps.
static bool use_mma_sync_b(int64_t rows_x, int64_t w_rows, int64_t k, int group) {
  static const int env = getenv("TG_W4_KERNEL") ? atoi(getenv("TG_W4_KERNEL")) : 0;  // tuning override
  const int mode = env ? env : g_w4_kernel;
  if (mode == 1) return false;
  if (rows_x > 4) return false;
  if (mode == 2) return true;
  // (groups of 32 / 64 - mx4, any4 g32: 4 group words per 128 k - are 4-5 % faster on the tcgen05 kernel even there)
  return rows_x == 1 && k <= 4096 && div_up(w_rows, 32) <= 148 && group >= 128;
}
int launch_quantize_any4_rows(const void* w, const float* sample_weight, int64_t n, int64_t k, int group, int inner_k_tiles,
                              int max_iter, float tol, int32_t* codes, int32_t* packed, void* sz, void* any4, void* lut,
                              int* iters, tg_dtype dt, cudaStream_t st);
int launch_gemm_w4_rm_A(void* y, const void* x, const int32_t* w, const void* sz, const void* lut,
                        const uint8_t* exps, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik,
                        tg_w4_format fmt, tg_dtype dt, const uint16_t* const_lut, cudaStream_t st);

namespace {

// int4 and mx4 run through the LUT kernels with a constant table per dtype
// (reference: Dequantization.cuh:136-260 "code - 8", FloatDefs.cuh:18-34 kMX4_Values).  Compile-time bit patterns:
// no init kernel, no stream ordering to worry about, usable inside CUDA graph capture from the first call on.
__device__ __align__(16) const uint16_t g_const_luts[4][16] = {
    // int4, bf16: -8 .. 7
    {0xC100, 0xC0E0, 0xC0C0, 0xC0A0, 0xC080, 0xC040, 0xC000, 0xBF80, 0x0000, 0x3F80, 0x4000, 0x4040, 0x4080, 0x40A0, 0x40C0, 0x40E0},
    // int4, fp16: -8 .. 7
    {0xC800, 0xC700, 0xC600, 0xC500, 0xC400, 0xC200, 0xC000, 0xBC00, 0x0000, 0x3C00, 0x4000, 0x4200, 0x4400, 0x4500, 0x4600, 0x4700},
    // mx4 (fp4 e2m1), bf16: 0, .5, 1, 1.5, 2, 3, 4, 6, -0, ...
    {0x0000, 0x3F00, 0x3F80, 0x3FC0, 0x4000, 0x4040, 0x4080, 0x40C0, 0x8000, 0xBF00, 0xBF80, 0xBFC0, 0xC000, 0xC040, 0xC080, 0xC0C0},
    // mx4, fp16
    {0x0000, 0x3800, 0x3C00, 0x3E00, 0x4000, 0x4200, 0x4400, 0x4600, 0x8000, 0xB800, 0xBC00, 0xBE00, 0xC000, 0xC200, 0xC400, 0xC600},
};

// device address of the constant LUT for (fmt, dt) on the current device
int const_lut_for(tg_w4_format fmt, tg_dtype dt, cudaStream_t, const uint16_t** out) {
  static thread_local const uint16_t* base[kMaxDevices] = {nullptr};
  const int dev = current_device_slot();
  if (base[dev] == nullptr) {
    const uint16_t* p = nullptr;
    if (cudaGetSymbolAddress((void**)&p, g_const_luts) != cudaSuccess) {
      set_error("cudaGetSymbolAddress failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    base[dev] = p;
  }
  *out = base[dev] + ((fmt == TG_W4_MX4 ? 2 : 0) + (dt == TG_FP16 ? 1 : 0)) * 16;
  return TG_OK;
}

bool valid_group(int g) { return g == 32 || g == 64 || g == 128 || g == 256; }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_common(const char* fn, const void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows,
                 int64_t k, tg_weight_side side, tg_dtype dt) {
  TG_REQUIRE(y && x && w, "%s: null tensor pointer", fn);
  TG_REQUIRE(dt == TG_BF16 || dt == TG_FP16, "%s: dtype must be bf16 or fp16", fn);
  TG_REQUIRE(side == TG_WEIGHT_A || side == TG_WEIGHT_B, "%s: bad weight side", fn);
  TG_REQUIRE(rows_x >= 0 && w_rows > 0 && k > 0, "%s: bad sizes rows_x=%lld w_rows=%lld k=%lld", fn, (long long)rows_x,
             (long long)w_rows, (long long)k);
  TG_REQUIRE(k % 32 == 0, "%s: k (%lld) must be a multiple of 32", fn, (long long)k);
  const int rt = side == TG_WEIGHT_A ? 16 : 8;
  TG_REQUIRE(w_rows % rt == 0, "%s: padded weight rows (%lld) must be a multiple of %d", fn, (long long)w_rows, rt);
  TG_REQUIRE(k < (1ll << 30) && w_rows < (1ll << 30) && rows_x < (1ll << 30), "%s: sizes exceed 32-bit indexing", fn);
  TG_REQUIRE(aligned16(x) && aligned16(w) && (reinterpret_cast<uintptr_t>(y) & 1u) == 0,
             "%s: x and w must be 16-byte aligned", fn);
  return TG_OK;
}

}  // namespace
}  // namespace tg

using namespace tg;

extern "C" {

const char* tg_last_error(void) { return g_err; }
const char* tg_version(void) { return "tinygemm_b200 0.1 sm_100a"; }
uint64_t tg_launch_count(void) { return g_launches; }
void tg_reset_launch_count(void) { g_launches = 0; }

int tg_set_option(tg_option option, int value) {
  switch (option) {
    case TG_OPT_PDL: set_w4_options(value != 0, -1); return TG_OK;
    case TG_OPT_STATIC_WEIGHTS: set_w4_options(-1, value != 0); return TG_OK;
    case TG_OPT_W4_KERNEL:
      if (value < 0 || value > 2) break;
      g_w4_kernel = value;
      return TG_OK;
  }
  set_error("tg_set_option: unknown option %d", (int)option);
  return TG_ERR_INVALID_ARGUMENT;
}

int tg_gemm_w4_rm(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                  const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                  int inner_k_tiles, tg_w4_format format, tg_weight_side side, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w4_rm";
  int rc = check_common(fn, y, x, w, rows_x, w_rows, k, side, dtype);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(format >= TG_W4_INT4 && format <= TG_W4_MX4, "%s: bad format", fn);
  TG_REQUIRE(valid_group(group), "%s: qGroupSize must be 32, 64, 128 or 256 (got %d)", fn, group);
  TG_REQUIRE(k % group == 0, "%s: k (%lld) must be a multiple of qGroupSize (%d)", fn, (long long)k, group);
  const int ik = inner_k_tiles;
  if (side == TG_WEIGHT_B) {
    TG_REQUIRE(ik == 2 || ik == 4 || ik == 8, "%s: B-layout innerKTiles must be 2, 4 or 8 (got %d)", fn, ik);
  } else {
    TG_REQUIRE(ik == 1 || ik == 2 || ik == 4, "%s: A-layout innerKTiles must be 1, 2 or 4 (got %d)", fn, ik);
  }
  TG_REQUIRE((k / 16) % ik == 0, "%s: k/16 (%lld) must be a multiple of innerKTiles (%d)", fn, (long long)(k / 16), ik);
  if (format == TG_W4_MX4) {
    TG_REQUIRE(exponents != nullptr, "%s: mx4 needs exponents", fn);
    if (dtype != TG_BF16) {
      set_error("%s: mx4 supports bf16 activations only", fn);
      return TG_ERR_UNSUPPORTED;
    }
  } else {
    TG_REQUIRE(scales_zeros != nullptr && aligned16(scales_zeros), "%s: scales_zeros missing or not 16-byte aligned", fn);
  }
  if (format == TG_W4_ANY4_GLOBAL || format == TG_W4_ANY4_ROWWISE) {
    TG_REQUIRE(lut != nullptr && aligned16(lut), "%s: any4 LUT missing or not 16-byte aligned", fn);
  }
  if (rows_x == 0) return TG_OK;
  const uint16_t* clut = nullptr;
  if (format == TG_W4_INT4 || format == TG_W4_MX4) {
    rc = const_lut_for(format, dtype, (cudaStream_t)stream, &clut);
    if (rc != TG_OK) return rc;
  }
  if (side == TG_WEIGHT_B) {
    if (use_mma_sync_b(rows_x, w_rows, k, group)) {
      rc = launch_gemm_w4_rm_B(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype, clut,
                               (cudaStream_t)stream);
      if (rc != TG_ERR_UNSUPPORTED) return rc;  // (k beyond that kernel's staging areas: the tcgen05 kernel takes any k)
    }
    return launch_gemm_w4_tc_B(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype, clut,
                               (cudaStream_t)stream);
  }
  return launch_gemm_w4_rm_A(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype, clut,
                             (cudaStream_t)stream);
}

// Several activation rows against an A-layout weight: the A kernel takes one row per launch, so from kRepackMinRows
// rows on the weight is repacked into the B layout (mostly into L2) and the one-pass tcgen05 kernel runs on that.
constexpr int64_t kRepackMinRows = 3;
size_t tg_gemm_w4_rm_workspace_bytes(int64_t rows_x, int64_t w_rows, int64_t k, tg_weight_side side) {
  if (side != TG_WEIGHT_A || rows_x < kRepackMinRows || k % 64 != 0 || w_rows % 16 != 0) return 0;
  return (size_t)w_rows * (size_t)k / 2;
}

int tg_gemm_w4_rm_ws(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                     const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group, int inner_k_tiles,
                     tg_w4_format format, tg_weight_side side, tg_dtype dtype, void* workspace, size_t workspace_bytes,
                     void* stream) {
  const size_t need = tg_gemm_w4_rm_workspace_bytes(rows_x, w_rows, k, side);
  if (need == 0 || workspace == nullptr || workspace_bytes < need)
    return tg_gemm_w4_rm(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, inner_k_tiles, format, side, dtype,
                         stream);
  const char* fn = "tg_gemm_w4_rm_ws";
  const int ik = inner_k_tiles == 0 ? 4 : inner_k_tiles;
  TG_REQUIRE(ik == 1 || ik == 2 || ik == 4, "%s: A-layout innerKTiles must be 1, 2 or 4 (got %d)", fn, ik);
  TG_REQUIRE(w != nullptr && aligned16(workspace), "%s: null weight or misaligned workspace", fn);
  int rc = tg_repack_Aint4_to_Bint4(w, static_cast<int32_t*>(workspace), w_rows, k, ik, 4, stream);
  if (rc != TG_OK) return rc;
  return tg_gemm_w4_rm(y, x, static_cast<const int32_t*>(workspace), scales_zeros, lut, exponents, rows_x, w_rows, k, group, 4,
                       format, TG_WEIGHT_B, dtype, stream);
}

// shared by the two row-sharded entry points; exchange_tag != 0: y_peers are the ranks' exchange buffers and y_local is
// this rank's plain output
static int gemm_w4_rm_sharded_impl(const char* fn, void* y_local, void* const* y_peers, int self_rank, uint32_t exchange_tag,
                                   int silu_pairs, int n_peers, int64_t y_row_stride, const void* x, const int32_t* w,
                                   const void* scales_zeros, const void* lut, const uint8_t* exponents, int64_t rows_x,
                                   int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_w4_format format,
                                   tg_dtype dtype, void* stream) {
  TG_REQUIRE(y_peers != nullptr && n_peers >= 1 && n_peers <= 8, "%s: 1..8 peer pointers are required", fn);
  for (int r = 0; r < n_peers; ++r) TG_REQUIRE(y_peers[r] != nullptr, "%s: null peer pointer %d", fn, r);
  const bool exchange = exchange_tag != 0u;
  if (exchange) {
    TG_REQUIRE(self_rank >= 0 && self_rank < n_peers, "%s: bad self_rank %d", fn, self_rank);
    TG_REQUIRE(y_local != nullptr && (reinterpret_cast<uintptr_t>(y_local) & 3u) == 0 && y_row_stride % 2 == 0,
               "%s: the local output must be 4-byte aligned with an even row stride", fn);
    for (int r = 0; r < n_peers; ++r)
      TG_REQUIRE((reinterpret_cast<uintptr_t>(y_peers[r]) & 7u) == 0, "%s: exchange buffer %d is not 8-byte aligned", fn, r);
    TG_REQUIRE(y_row_stride >= (w_rows * n_peers >> (silu_pairs ? 1 : 0)),
               "%s: output row stride (%lld) smaller than the full row (%lld)", fn, (long long)y_row_stride,
               (long long)(w_rows * n_peers >> (silu_pairs ? 1 : 0)));
    TG_REQUIRE(!silu_pairs || w_rows % 4 == 0, "%s: interleaved (gate, up) shards need w_rows %% 4 == 0", fn);
  } else {
    y_local = y_peers[0];
    TG_REQUIRE(y_row_stride >= w_rows, "%s: output row stride (%lld) smaller than the shard (%lld)", fn,
               (long long)y_row_stride, (long long)w_rows);
  }
  int rc = check_common(fn, y_local, x, w, rows_x, w_rows, k, TG_WEIGHT_B, dtype);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(valid_group(group), "%s: group size must be 32, 64, 128 or 256 (got %d)", fn, group);
  TG_REQUIRE(k % group == 0, "%s: k (%lld) must be a multiple of the group size (%d)", fn, (long long)k, group);
  TG_REQUIRE(format >= TG_W4_INT4 && format <= TG_W4_MX4, "%s: bad format", fn);
  if (format == TG_W4_MX4) {
    TG_REQUIRE(dtype == TG_BF16, "%s: mx4 supports bf16 activations only", fn);
    TG_REQUIRE(exponents != nullptr, "%s: mx4 needs the exponent tensor", fn);
  } else {
    TG_REQUIRE(scales_zeros != nullptr, "%s: scales_and_zeros is required", fn);
    TG_REQUIRE((reinterpret_cast<uintptr_t>(scales_zeros) & 3u) == 0, "%s: scales_and_zeros must be 4-byte aligned", fn);
  }
  if (format == TG_W4_ANY4_GLOBAL || format == TG_W4_ANY4_ROWWISE)
    TG_REQUIRE(lut != nullptr && aligned16(lut), "%s: any4 needs a 16-byte aligned LUT", fn);
  const int ik = inner_k_tiles == 0 ? 4 : inner_k_tiles;
  if (rows_x == 0) return TG_OK;
  const uint16_t* clut = nullptr;
  if (format == TG_W4_INT4 || format == TG_W4_MX4) {
    rc = const_lut_for(format, dtype, (cudaStream_t)stream, &clut);
    if (rc != TG_OK) return rc;
  }
  // the in-kernel exchange lives in the tcgen05 kernel; the mma.sync kernel keeps a plain peer-store variant
  if (exchange || !use_mma_sync_b(rows_x, w_rows, k, group))
    return launch_gemm_w4_tc_B(y_local, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype,
                               clut, (cudaStream_t)stream, y_peers, n_peers, y_row_stride, silu_pairs, self_rank,
                               exchange_tag);
  return launch_gemm_w4_rm_B(y_local, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype,
                             clut, (cudaStream_t)stream, y_peers, n_peers, y_row_stride);
}

int tg_gemm_w4_rm_sharded(void* const* y_peers, int n_peers, int64_t y_row_stride, const void* x, const int32_t* w,
                          const void* scales_zeros, const void* lut, const uint8_t* exponents, int64_t rows_x,
                          int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_w4_format format, tg_dtype dtype,
                          void* stream) {
  return gemm_w4_rm_sharded_impl("tg_gemm_w4_rm_sharded", nullptr, y_peers, 0, 0u, 0, n_peers, y_row_stride, x, w,
                                 scales_zeros, lut, exponents, rows_x, w_rows, k, group, inner_k_tiles, format, dtype, stream);
}

int tg_gemm_w4_rm_exchange(void* y, void* const* xchg_peers, int self_rank, uint32_t tag, int n_peers, int64_t y_row_stride,
                           const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                           const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                           int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w4_rm_exchange";
  TG_REQUIRE(tag != 0u, "%s: the call tag must not be 0", fn);
  return gemm_w4_rm_sharded_impl(fn, y, xchg_peers, self_rank, tag, 0, n_peers, y_row_stride, x, w, scales_zeros, lut,
                                 exponents, rows_x, w_rows, k, group, inner_k_tiles, format, dtype, stream);
}

int tg_gemm_w4_rm_exchange_silu_pairs(void* y, void* const* xchg_peers, int self_rank, uint32_t tag, int n_peers,
                                      int64_t y_row_stride, const void* x, const int32_t* w, const void* scales_zeros,
                                      const void* lut, const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k,
                                      int group, int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w4_rm_exchange_silu_pairs";
  TG_REQUIRE(tag != 0u, "%s: the call tag must not be 0", fn);
  return gemm_w4_rm_sharded_impl(fn, y, xchg_peers, self_rank, tag, 1, n_peers, y_row_stride, x, w, scales_zeros, lut,
                                 exponents, rows_x, w_rows, k, group, inner_k_tiles, format, dtype, stream);
}

int tg_quantize_any4_rows(const void* w, const float* sample_weight, int64_t n, int64_t k, int group, int inner_k_tiles,
                          int max_iter, float tol, int32_t* codes, int32_t* packed, void* scales_zeros, void* any4, void* lut,
                          int32_t* iters, tg_dtype dtype, void* stream) {
  const char* fn = "tg_quantize_any4_rows";
  TG_REQUIRE(w && scales_zeros && any4 && lut, "%s: null tensor pointer", fn);
  TG_REQUIRE(dtype == TG_BF16 || dtype == TG_FP16, "%s: dtype must be bf16 or fp16", fn);
  TG_REQUIRE(n > 0 && k > 0 && n < (1ll << 30), "%s: bad sizes n=%lld k=%lld", fn, (long long)n, (long long)k);
  TG_REQUIRE(valid_group(group) && k % group == 0, "%s: group size must be 32, 64, 128 or 256 and divide k", fn);
  TG_REQUIRE(k * 5 <= 227 * 1024, "%s: k = %lld does not fit the per-row shared-memory working set (k <= 46489)", fn,
             (long long)k);
  TG_REQUIRE(max_iter >= 1 && tol >= 0.f, "%s: bad max_iter / tol", fn);
  if (packed != nullptr) {
    TG_REQUIRE(inner_k_tiles == 2 || inner_k_tiles == 4 || inner_k_tiles == 8, "%s: B-layout innerKTiles must be 2, 4 or 8", fn);
    TG_REQUIRE(n % 8 == 0 && k % 32 == 0, "%s: the packed layout needs n %% 8 == 0 and k %% 32 == 0", fn);
  }
  return launch_quantize_any4_rows(w, sample_weight, n, k, group, inner_k_tiles, max_iter, tol, codes, packed, scales_zeros,
                                   any4, lut, iters, dtype, (cudaStream_t)stream);
}

int tg_gemm_w4_rm_silu_pairs(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                             const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                             int inner_k_tiles, tg_w4_format format, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w4_rm_silu_pairs";
  int rc = check_common(fn, y, x, w, rows_x, w_rows, k, TG_WEIGHT_B, dtype);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(format >= TG_W4_INT4 && format <= TG_W4_MX4, "%s: bad format", fn);
  TG_REQUIRE(valid_group(group), "%s: qGroupSize must be 32, 64, 128 or 256 (got %d)", fn, group);
  TG_REQUIRE(k % group == 0, "%s: k must be a multiple of qGroupSize", fn);
  const int ik = inner_k_tiles;
  TG_REQUIRE(ik == 2 || ik == 4 || ik == 8, "%s: B-layout innerKTiles must be 2, 4 or 8 (got %d)", fn, ik);
  TG_REQUIRE((k / 16) % ik == 0, "%s: k/16 must be a multiple of innerKTiles", fn);
  if (format == TG_W4_MX4) {
    TG_REQUIRE(exponents != nullptr && dtype == TG_BF16, "%s: mx4 needs exponents and bf16", fn);
  } else {
    TG_REQUIRE(scales_zeros != nullptr && aligned16(scales_zeros), "%s: scales_zeros missing or misaligned", fn);
  }
  if (format == TG_W4_ANY4_GLOBAL || format == TG_W4_ANY4_ROWWISE)
    TG_REQUIRE(lut != nullptr && aligned16(lut), "%s: any4 LUT missing or misaligned", fn);
  if (rows_x == 0) return TG_OK;
  const uint16_t* clut = nullptr;
  if (format == TG_W4_INT4 || format == TG_W4_MX4) {
    rc = const_lut_for(format, dtype, (cudaStream_t)stream, &clut);
    if (rc != TG_OK) return rc;
  }
  if (!use_mma_sync_b(rows_x, w_rows, k, group))
    return launch_gemm_w4_tc_B(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype, clut,
                               (cudaStream_t)stream, nullptr, 0, 0, /*silu_pairs=*/1);
  return launch_gemm_w4_rm_B(y, x, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, ik, format, dtype, clut,
                             (cudaStream_t)stream, nullptr, 0, 0, /*silu_pairs=*/1);
}

int tg_gemm_w8_rm(void* y, const void* x, const int32_t* w, const void* scales_zeros, int64_t rows_x,
                  int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_weight_side side, tg_dtype dtype,
                  void* stream) {
  const char* fn = "tg_gemm_w8_rm";
  int rc = check_common(fn, y, x, w, rows_x, w_rows, k, side, dtype);
  if (rc != TG_OK) return rc;
  TG_REQUIRE(valid_group(group), "%s: qGroupSize must be 32, 64, 128 or 256 (got %d)", fn, group);
  TG_REQUIRE(k % group == 0, "%s: k (%lld) must be a multiple of qGroupSize (%d)", fn, (long long)k, group);
  TG_REQUIRE(scales_zeros != nullptr, "%s: scales_zeros missing", fn);
  const int ik = inner_k_tiles;
  if (side == TG_WEIGHT_B) {
    TG_REQUIRE(ik == 1 || ik == 2 || ik == 4, "%s: B-layout innerKTiles must be 1, 2 or 4 (got %d)", fn, ik);
  } else {
    TG_REQUIRE(ik == 1 || ik == 2, "%s: A-layout innerKTiles must be 1 or 2 (got %d)", fn, ik);
  }
  TG_REQUIRE((k / 16) % ik == 0, "%s: k/16 must be a multiple of innerKTiles (%d)", fn, ik);
  if (rows_x == 0) return TG_OK;
  return launch_gemm_w8_rm(y, x, w, scales_zeros, rows_x, w_rows, k, group, ik, side, dtype, (cudaStream_t)stream);
}

int tg_gemm_w16_rm(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k,
                   int inner_k_tiles, tg_weight_side side, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w16_rm";
  int rc = check_common(fn, y, x, w, rows_x, w_rows, k, side, dtype);
  if (rc != TG_OK) return rc;
  const int ik = inner_k_tiles;
  if (side == TG_WEIGHT_B) {
    TG_REQUIRE(ik == 1 || ik == 2, "%s: B-layout innerKTiles must be 1 or 2 (got %d)", fn, ik);
  } else {
    TG_REQUIRE(ik == 1, "%s: A-layout innerKTiles must be 1 (got %d)", fn, ik);
  }
  TG_REQUIRE((k / 16) % ik == 0, "%s: k/16 must be a multiple of innerKTiles (%d)", fn, ik);
  if (rows_x == 0) return TG_OK;
  return launch_gemm_w16_rm(y, x, w, rows_x, w_rows, k, ik, side, dtype, (cudaStream_t)stream);
}

// ---- tensor-core-layout activations / outputs: RM staging in the caller's workspace ----
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t tg_gemm_tc_workspace_bytes(int64_t rows_x, int64_t w_rows, int64_t k) {
  return align256((size_t)rows_x * (size_t)k * 2) + align256((size_t)rows_x * (size_t)w_rows * 2);
}

// unpack x (A layout when the weight is on the right, B layout otherwise) into the workspace
static int tc_unpack_x(const void* x, void* xs, int64_t rows_x, int64_t k, int x_ik, tg_weight_side side, void* stream) {
  if (side == TG_WEIGHT_B) return tg_convert_from_A(x, xs, rows_x, k, stream);
  return tg_convert_from_B(x, xs, rows_x, k, x_ik, stream);
}
// pack y [rows_x][w_rows] into the layout of the activation operand
static int tc_pack_y(const void* ys, void* y, int64_t rows_x, int64_t w_rows, int x_ik, tg_weight_side side, void* stream) {
  if (side == TG_WEIGHT_B) return tg_convert_to_A(ys, y, rows_x, w_rows, stream);
  return tg_convert_to_B(ys, y, rows_x, w_rows, x_ik, stream);
}

#define TG_TC_PROLOGUE(fn)                                                                         \
  TG_REQUIRE(workspace != nullptr && aligned16(workspace), "%s: workspace missing or misaligned", fn); \
  TG_REQUIRE(x_inner_k_tiles == 1 || (side == TG_WEIGHT_A && x_inner_k_tiles == 2),                \
             "%s: bad activation innerKTiles %d", fn, x_inner_k_tiles);                            \
  TG_REQUIRE(rows_x % (side == TG_WEIGHT_B ? 16 : 8) == 0, "%s: rows_x must be padded to the tile size", fn); \
  char* xs = (char*)workspace;                                                                     \
  char* ys = xs + align256((size_t)rows_x * (size_t)k * 2);                                        \
  if (rows_x == 0) return TG_OK;                                                                   \
  int rc = tc_unpack_x(x, xs, rows_x, k, x_inner_k_tiles, side, stream);                           \
  if (rc != TG_OK) return rc;

int tg_gemm_w4_tc(void* y, const void* x, const int32_t* w, const void* scales_zeros, const void* lut,
                  const uint8_t* exponents, int64_t rows_x, int64_t w_rows, int64_t k, int group,
                  int inner_k_tiles, int x_inner_k_tiles, tg_w4_format format, tg_weight_side side,
                  tg_dtype dtype, void* workspace, void* stream) {
  TG_TC_PROLOGUE("tg_gemm_w4_tc");
  rc = tg_gemm_w4_rm(ys, xs, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, inner_k_tiles, format, side,
                     dtype, stream);
  if (rc != TG_OK) return rc;
  return tc_pack_y(ys, y, rows_x, w_rows, x_inner_k_tiles, side, stream);
}

int tg_gemm_w8_tc(void* y, const void* x, const int32_t* w, const void* scales_zeros, int64_t rows_x,
                  int64_t w_rows, int64_t k, int group, int inner_k_tiles, int x_inner_k_tiles,
                  tg_weight_side side, tg_dtype dtype, void* workspace, void* stream) {
  TG_TC_PROLOGUE("tg_gemm_w8_tc");
  rc = tg_gemm_w8_rm(ys, xs, w, scales_zeros, rows_x, w_rows, k, group, inner_k_tiles, side, dtype, stream);
  if (rc != TG_OK) return rc;
  return tc_pack_y(ys, y, rows_x, w_rows, x_inner_k_tiles, side, stream);
}

int tg_gemm_w16_tc(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k,
                   int inner_k_tiles, int x_inner_k_tiles, tg_weight_side side, tg_dtype dtype,
                   void* workspace, void* stream) {
  TG_TC_PROLOGUE("tg_gemm_w16_tc");
  rc = tg_gemm_w16_rm(ys, xs, w, rows_x, w_rows, k, inner_k_tiles, side, dtype, stream);
  if (rc != TG_OK) return rc;
  return tc_pack_y(ys, y, rows_x, w_rows, x_inner_k_tiles, side, stream);
}

}  // extern "C"
