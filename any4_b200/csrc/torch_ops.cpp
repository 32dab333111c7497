// torch custom-op layer: re-creates the reference's `torch.ops.tinygemm.*` surface (19 ops,
// schemas identical to tinygemm_lib/TinyGemm.cpp:19-121) on top of the C ABI in
// include/tinygemm_b200.h.  Host-side behaviour mirrors the reference dispatch functions
// (TinyGemm_int4.cu:28-292, :294-548, :550-794; TinyGemm_int8.cu; TinyGemm_bf16.cu;
// TinyGemmConvertA.cu / TinyGemmConvertB.cu host halves): same shape / dtype checks raised as
// c10::Error (-> RuntimeError), output freshly allocated on the input's device through the
// caching allocator (CUDA-graph safe), work enqueued on the current stream, no host sync.
// PyTorch is plumbing here; all compute is in libtinygemm_b200.so.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "tinygemm_b200.h"

namespace {

constexpr int64_t kWarp = 32;

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

void check_rc(int rc, const char* op) {
  TORCH_CHECK(rc == TG_OK, "tinygemm::", op, " failed: ", tg_last_error());
}

void* cur_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

tg_dtype dtype_of(const at::Tensor& t) {
  TORCH_CHECK(t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kHalf,
              "tinygemm: activations must be bfloat16 or float16");
  return t.scalar_type() == at::kBFloat16 ? TG_BF16 : TG_FP16;
}

void check_cuda(const at::Tensor& t) { TORCH_CHECK(t.is_cuda(), "tinygemm: tensor must live on a CUDA device"); }

// ---------------------------------------------------------------------------------------
// layout conversion ops
// ---------------------------------------------------------------------------------------
at::Tensor to_A(const at::Tensor& in, int64_t innerKTiles) {
  check_cuda(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(innerKTiles == 1);
  TORCH_CHECK(in.dim() == 2 && in.is_contiguous());
  dtype_of(in);
  auto out = at::empty({div_up(in.size(0), 16), div_up(in.size(1), 16), kWarp, 8}, in.options());
  check_rc(tg_convert_to_A(in.data_ptr(), out.data_ptr(), in.size(0), in.size(1), cur_stream()), "to_A");
  return out;
}

at::Tensor from_A(const at::Tensor& in, int64_t m, int64_t k) {
  check_cuda(in);
  c10::cuda::CUDAGuard g(in.device());
  dtype_of(in);
  TORCH_CHECK(in.is_contiguous() && in.dim() == 4);
  TORCH_CHECK(div_up(m, 16) == in.size(0));
  TORCH_CHECK(div_up(k, 16) == in.size(1));
  TORCH_CHECK(in.size(2) == kWarp && in.size(3) == 8);
  auto out = at::empty({m, k}, in.options());
  check_rc(tg_convert_from_A(in.data_ptr(), out.data_ptr(), m, k, cur_stream()), "from_A");
  return out;
}

at::Tensor to_B(const at::Tensor& in, int64_t innerKTiles) {
  check_cuda(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(innerKTiles == 1 || innerKTiles == 2);
  TORCH_CHECK(in.dim() == 2 && in.is_contiguous());
  dtype_of(in);
  auto out = at::empty({div_up(in.size(0), 8), div_up(in.size(1), innerKTiles * 16), kWarp, innerKTiles * 4},
                       in.options());
  check_rc(tg_convert_to_B(in.data_ptr(), out.data_ptr(), in.size(0), in.size(1), (int)innerKTiles, cur_stream()),
           "to_B");
  return out;
}

at::Tensor from_B(const at::Tensor& in, int64_t n, int64_t k) {
  check_cuda(in);
  c10::cuda::CUDAGuard g(in.device());
  dtype_of(in);
  TORCH_CHECK(in.is_contiguous() && in.dim() == 4);
  TORCH_CHECK(in.size(3) % 4 == 0);
  const int64_t ik = in.size(3) / 4;
  TORCH_CHECK(ik == 1 || ik == 2);
  TORCH_CHECK(div_up(n, 8) == in.size(0));
  TORCH_CHECK(div_up(k, 16 * ik) == in.size(1));
  TORCH_CHECK(in.size(2) == kWarp);
  auto out = at::empty({n, k}, in.options());
  check_rc(tg_convert_from_B(in.data_ptr(), out.data_ptr(), n, k, (int)ik, cur_stream()), "from_B");
  return out;
}

void check_codes(const at::Tensor& in) {
  check_cuda(in);
  TORCH_CHECK(in.dim() == 2);
  TORCH_CHECK(in.scalar_type() == at::kInt);
  TORCH_CHECK(in.is_contiguous());
}

at::Tensor to_Aint4(const at::Tensor& in, int64_t ik) {
  check_codes(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(ik == 1 || ik == 2 || ik == 4);
  auto out = at::empty({div_up(in.size(0), 16), div_up(in.size(1), ik * 16), kWarp, ik}, in.options());
  check_rc(tg_convert_to_Aint4(in.data_ptr<int32_t>(), out.data_ptr<int32_t>(), in.size(0), in.size(1), (int)ik,
                               cur_stream()),
           "to_Aint4");
  return out;
}

at::Tensor to_Aint8(const at::Tensor& in, int64_t ik) {
  check_codes(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(ik == 1 || ik == 2);
  auto out = at::empty({div_up(in.size(0), 16), div_up(div_up(in.size(1), 16), ik), kWarp, 2 * ik}, in.options());
  check_rc(tg_convert_to_Aint8(in.data_ptr<int32_t>(), out.data_ptr<int32_t>(), in.size(0), in.size(1), (int)ik,
                               cur_stream()),
           "to_Aint8");
  return out;
}

at::Tensor to_Bint4(const at::Tensor& in, int64_t ik) {
  check_codes(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(ik == 2 || ik == 4 || ik == 8);
  TORCH_CHECK(in.size(1) % (ik * 16) == 0);
  auto out = at::empty({div_up(in.size(0), 8), in.size(1) / (ik * 16), kWarp, ik / 2}, in.options());
  check_rc(tg_convert_to_Bint4(in.data_ptr<int32_t>(), out.data_ptr<int32_t>(), in.size(0), in.size(1), (int)ik,
                               cur_stream()),
           "to_Bint4");
  return out;
}

at::Tensor to_Bint8(const at::Tensor& in, int64_t ik) {
  check_codes(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(ik == 1 || ik == 2 || ik == 4);
  TORCH_CHECK(in.size(1) % (ik * 16) == 0);
  auto out = at::empty({div_up(in.size(0), 8), in.size(1) / (ik * 16), kWarp, ik}, in.options());
  check_rc(tg_convert_to_Bint8(in.data_ptr<int32_t>(), out.data_ptr<int32_t>(), in.size(0), in.size(1), (int)ik,
                               cur_stream()),
           "to_Bint8");
  return out;
}

// ---------------------------------------------------------------------------------------
// GEMM ops.  `A` is the left operand, `B` the right one; weightOnRight says which is the weight.
// ---------------------------------------------------------------------------------------
enum class WKind { W4, W8, W16 };

struct Problem {
  const at::Tensor* x = nullptr;  // activations
  const at::Tensor* w = nullptr;  // packed weight
  int64_t rows_x = 0;             // activation rows (padded for TC layouts)
  int64_t w_rows = 0;             // padded weight rows
  int64_t k = 0;
  int w_ik = 1;
  int x_ik = 1;
  tg_weight_side side = TG_WEIGHT_B;
  tg_dtype dt = TG_BF16;
};

// Decode innerKTiles of the packed weight from its innermost size and validate it
int weight_ik(WKind kind, bool right, int64_t last) {
  switch (kind) {
    case WKind::W4:
      if (right) {
        TORCH_CHECK(last == 1 || last == 2 || last == 4);
        return (int)last * 2;
      }
      TORCH_CHECK(last == 1 || last == 2 || last == 4);
      return (int)last;
    case WKind::W8:
      if (right) {
        TORCH_CHECK(last == 1 || last == 2 || last == 4);
        return (int)last;
      }
      TORCH_CHECK(last == 2 || last == 4);
      return (int)last / 2;
    case WKind::W16:
      if (right) {
        TORCH_CHECK(last == 4 || last == 8);
        return (int)last / 4;
      }
      TORCH_CHECK(last == 8);
      return 1;
  }
  return 1;
}

Problem parse_rm(const at::Tensor& A, const at::Tensor& B, bool weightOnRight, WKind kind) {
  check_cuda(A);
  check_cuda(B);
  TORCH_CHECK(A.device() == B.device());
  Problem p;
  p.side = weightOnRight ? TG_WEIGHT_B : TG_WEIGHT_A;
  p.x = weightOnRight ? &A : &B;
  p.w = weightOnRight ? &B : &A;
  TORCH_CHECK(p.x->dim() == 2 && p.x->is_contiguous());
  TORCH_CHECK(p.w->dim() == 4 && p.w->is_contiguous());
  if (kind != WKind::W16) TORCH_CHECK(p.w->scalar_type() == at::kInt);
  p.dt = dtype_of(*p.x);
  if (kind == WKind::W16) TORCH_CHECK(p.w->scalar_type() == p.x->scalar_type());
  p.rows_x = p.x->size(0);
  p.k = p.x->size(1);
  const int64_t kTiles = div_up(p.k, 16);
  p.w_rows = p.w->size(0) * (weightOnRight ? 8 : 16);
  p.w_ik = weight_ik(kind, weightOnRight, p.w->size(3));
  TORCH_CHECK(p.w->size(1) == div_up(kTiles, p.w_ik));
  TORCH_CHECK(p.w->size(2) == kWarp);
  TORCH_CHECK(p.k % 32 == 0);
  TORCH_CHECK(kTiles % p.w_ik == 0);
  return p;
}

Problem parse_tc(const at::Tensor& A, const at::Tensor& B, bool weightOnRight, WKind kind) {
  check_cuda(A);
  check_cuda(B);
  TORCH_CHECK(A.device() == B.device());
  TORCH_CHECK(A.is_contiguous() && A.dim() == 4 && A.size(2) == kWarp);
  TORCH_CHECK(B.is_contiguous() && B.dim() == 4 && B.size(2) == kWarp);
  Problem p;
  p.side = weightOnRight ? TG_WEIGHT_B : TG_WEIGHT_A;
  p.x = weightOnRight ? &A : &B;
  p.w = weightOnRight ? &B : &A;
  if (kind != WKind::W16) TORCH_CHECK(p.w->scalar_type() == at::kInt);
  p.dt = dtype_of(*p.x);
  if (kind == WKind::W16) TORCH_CHECK(p.w->scalar_type() == p.x->scalar_type());
  int64_t kTilesX;
  if (weightOnRight) {
    TORCH_CHECK(A.size(3) == 8);  // activations in the A layout
    p.x_ik = 1;
    p.rows_x = A.size(0) * 16;
    kTilesX = A.size(1);
  } else {
    TORCH_CHECK(B.size(3) == 4 || B.size(3) == 8);  // activations in the B layout
    p.x_ik = (int)B.size(3) / 4;
    p.rows_x = B.size(0) * 8;
    kTilesX = B.size(1) * p.x_ik;
  }
  p.w_rows = p.w->size(0) * (weightOnRight ? 8 : 16);
  p.w_ik = weight_ik(kind, weightOnRight, p.w->size(3));
  const int64_t kTilesW = p.w->size(1) * p.w_ik;
  TORCH_CHECK(kTilesX == kTilesW);
  p.k = kTilesW * 16;
  TORCH_CHECK(p.k % 32 == 0);
  return p;
}

at::Tensor alloc_tc_out(const Problem& p) {
  if (p.side == TG_WEIGHT_B) {
    // A layout with the padded weight rows as the k dimension: [mT][ceil(nT/2)][32][8]
    return at::empty({p.rows_x / 16, div_up(p.w_rows, 16), kWarp, 8}, p.x->options());
  }
  // B layout: [nT][ceil(mT / x_ik)][32][x_ik * 4]
  return at::empty({p.rows_x / 8, div_up(p.w_rows / 16, p.x_ik), kWarp, p.x_ik * 4}, p.x->options());
}

struct QuantArgs {
  int64_t group = 32;
  const void* sz = nullptr;
  const void* lut = nullptr;
  const uint8_t* exps = nullptr;
  tg_w4_format fmt = TG_W4_INT4;
};

// group scale / zero checks shared by int4, any4, int8 (TinyGemm_int4.cu:103-121, :373-391)
void check_scales(const Problem& p, int64_t qGroupSize, const at::Tensor& sz, bool strict_group) {
  TORCH_CHECK(qGroupSize == 32 || qGroupSize == 64 || qGroupSize == 128 || qGroupSize == 256);
  TORCH_CHECK(sz.device() == p.x->device());
  TORCH_CHECK(sz.dim() == 3);
  if (strict_group) {
    TORCH_CHECK(p.k % qGroupSize == 0);
    TORCH_CHECK(sz.size(0) == p.k / qGroupSize);
  } else {
    TORCH_CHECK(sz.size(0) > 0 && p.k % sz.size(0) == 0);
    TORCH_CHECK(p.k / sz.size(0) == qGroupSize, "tinygemm: qScaleAndZeros has ", sz.size(0), " groups but k / qGroupSize = ",
                p.k / qGroupSize);
  }
  TORCH_CHECK(sz.size(1) == p.w_rows);
  TORCH_CHECK(sz.size(2) == 2);
  TORCH_CHECK(sz.scalar_type() == p.x->scalar_type(), "tinygemm: qScaleAndZeros must have the activation dtype");
  TORCH_CHECK(sz.is_contiguous());
}

void check_lut(const Problem& p, const at::Tensor& lut, QuantArgs& q) {
  TORCH_CHECK(lut.device() == p.x->device());
  TORCH_CHECK(lut.scalar_type() == p.x->scalar_type());
  TORCH_CHECK(lut.is_contiguous());
  if (lut.dim() == 1) {
    TORCH_CHECK(lut.size(0) == 16);
    q.fmt = TG_W4_ANY4_GLOBAL;
  } else if (lut.dim() == 2) {
    TORCH_CHECK(lut.size(0) == p.w_rows && lut.size(1) == 16);
    q.fmt = TG_W4_ANY4_ROWWISE;
  } else {
    TORCH_CHECK(false, "invalid any4 dequantization tensor");
  }
  q.lut = lut.data_ptr();
}

void check_mx4(const Problem& p, int64_t qGroupSize, const at::Tensor& e, QuantArgs& q) {
  TORCH_CHECK(qGroupSize == 32 || qGroupSize == 64 || qGroupSize == 128 || qGroupSize == 256);
  TORCH_CHECK(p.k % qGroupSize == 0);
  TORCH_CHECK(e.device() == p.x->device());
  TORCH_CHECK(e.scalar_type() == at::kByte);
  TORCH_CHECK(e.dim() == 2);
  TORCH_CHECK(e.size(0) == p.w_rows);
  TORCH_CHECK(e.size(1) == p.k / qGroupSize);
  TORCH_CHECK(e.is_contiguous());
  TORCH_CHECK(p.x->scalar_type() == at::kBFloat16, "tinygemm: mx4 supports bfloat16 activations only");
  q.fmt = TG_W4_MX4;
  q.exps = e.data_ptr<uint8_t>();
}

at::Tensor run_rm(const Problem& p, WKind kind, const QuantArgs& q, const char* op) {
  c10::cuda::CUDAGuard g(p.x->device());
  auto y = at::empty({p.rows_x, p.w_rows}, p.x->options());
  int rc = TG_OK;
  switch (kind) {
    case WKind::W4: {
      // (several activation rows against an A-layout weight: scratch for the B-layout repack, from the caching allocator)
      const size_t ws_bytes = tg_gemm_w4_rm_workspace_bytes(p.rows_x, p.w_rows, p.k, p.side);
      at::Tensor ws;
      if (ws_bytes) ws = at::empty({(int64_t)ws_bytes}, p.x->options().dtype(at::kByte));
      rc = tg_gemm_w4_rm_ws(y.data_ptr(), p.x->data_ptr(), (const int32_t*)p.w->data_ptr(), q.sz, q.lut, q.exps, p.rows_x,
                            p.w_rows, p.k, (int)q.group, p.w_ik, q.fmt, p.side, p.dt, ws_bytes ? ws.data_ptr() : nullptr,
                            ws_bytes, cur_stream());
      break;
    }
    case WKind::W8:
      rc = tg_gemm_w8_rm(y.data_ptr(), p.x->data_ptr(), (const int32_t*)p.w->data_ptr(), q.sz, p.rows_x, p.w_rows, p.k,
                         (int)q.group, p.w_ik, p.side, p.dt, cur_stream());
      break;
    case WKind::W16:
      rc = tg_gemm_w16_rm(y.data_ptr(), p.x->data_ptr(), p.w->data_ptr(), p.rows_x, p.w_rows, p.k, p.w_ik, p.side, p.dt,
                          cur_stream());
      break;
  }
  check_rc(rc, op);
  return y;
}

at::Tensor run_tc(const Problem& p, WKind kind, const QuantArgs& q, const char* op) {
  c10::cuda::CUDAGuard g(p.x->device());
  auto y = alloc_tc_out(p);
  auto ws = at::empty({(int64_t)tg_gemm_tc_workspace_bytes(p.rows_x, p.w_rows, p.k)},
                      p.x->options().dtype(at::kByte));
  int rc = TG_OK;
  switch (kind) {
    case WKind::W4:
      rc = tg_gemm_w4_tc(y.data_ptr(), p.x->data_ptr(), (const int32_t*)p.w->data_ptr(), q.sz, q.lut, q.exps, p.rows_x,
                         p.w_rows, p.k, (int)q.group, p.w_ik, p.x_ik, q.fmt, p.side, p.dt, ws.data_ptr(), cur_stream());
      break;
    case WKind::W8:
      rc = tg_gemm_w8_tc(y.data_ptr(), p.x->data_ptr(), (const int32_t*)p.w->data_ptr(), q.sz, p.rows_x, p.w_rows, p.k,
                         (int)q.group, p.w_ik, p.x_ik, p.side, p.dt, ws.data_ptr(), cur_stream());
      break;
    case WKind::W16:
      rc = tg_gemm_w16_tc(y.data_ptr(), p.x->data_ptr(), p.w->data_ptr(), p.rows_x, p.w_rows, p.k, p.w_ik, p.x_ik,
                          p.side, p.dt, ws.data_ptr(), cur_stream());
      break;
  }
  check_rc(rc, op);
  return y;
}

template <bool TC>
at::Tensor gemm(const at::Tensor& A, const at::Tensor& B, bool right, WKind kind, QuantArgs& q, int64_t group,
                const at::Tensor* sz, const at::Tensor* lut, const at::Tensor* exps, const char* op) {
  Problem p = TC ? parse_tc(A, B, right, kind) : parse_rm(A, B, right, kind);
  q.group = group;
  if (sz) {
    check_scales(p, group, *sz, /*strict_group=*/TC);
    q.sz = sz->data_ptr();
  }
  if (lut) check_lut(p, *lut, q);
  if (exps) check_mx4(p, group, *exps, q);
  return TC ? run_tc(p, kind, q, op) : run_rm(p, kind, q, op);
}

// ---- the public ops (names and argument order: TinyGemm.cpp:47-118) ----
at::Tensor y_TC_int4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, bool right) {
  QuantArgs q;
  return gemm<true>(A, B, right, WKind::W4, q, g, &sz, nullptr, nullptr, "tinygemm_y_f16TC_x_f16TC_w_int4TC");
}
at::Tensor y_RM_int4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, bool right) {
  QuantArgs q;
  return gemm<false>(A, B, right, WKind::W4, q, g, &sz, nullptr, nullptr, "tinygemm_y_f16RM_x_f16RM_w_int4TC");
}
at::Tensor y_TC_any4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, at::Tensor lut, bool right) {
  QuantArgs q;
  return gemm<true>(A, B, right, WKind::W4, q, g, &sz, &lut, nullptr, "tinygemm_y_f16TC_x_f16TC_w_any4TC");
}
at::Tensor y_RM_any4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, at::Tensor lut, bool right) {
  QuantArgs q;
  return gemm<false>(A, B, right, WKind::W4, q, g, &sz, &lut, nullptr, "tinygemm_y_f16RM_x_f16RM_w_any4TC");
}
at::Tensor y_TC_mx4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor e, bool right) {
  QuantArgs q;
  return gemm<true>(A, B, right, WKind::W4, q, g, nullptr, nullptr, &e, "tinygemm_y_f16TC_x_f16TC_w_mx4TC");
}
at::Tensor y_RM_mx4(at::Tensor A, at::Tensor B, int64_t g, at::Tensor e, bool right) {
  QuantArgs q;
  return gemm<false>(A, B, right, WKind::W4, q, g, nullptr, nullptr, &e, "tinygemm_y_f16RM_x_f16RM_w_mx4TC");
}
at::Tensor y_TC_int8(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, bool right) {
  QuantArgs q;
  return gemm<true>(A, B, right, WKind::W8, q, g, &sz, nullptr, nullptr, "tinygemm_y_f16TC_x_f16TC_w_int8TC");
}
at::Tensor y_RM_int8(at::Tensor A, at::Tensor B, int64_t g, at::Tensor sz, bool right) {
  QuantArgs q;
  return gemm<false>(A, B, right, WKind::W8, q, g, &sz, nullptr, nullptr, "tinygemm_y_f16RM_x_f16RM_w_int8TC");
}
at::Tensor y_TC_f16(at::Tensor A, at::Tensor B, bool right) {
  QuantArgs q;
  return gemm<true>(A, B, right, WKind::W16, q, 32, nullptr, nullptr, nullptr, "tinygemm_y_f16TC_x_f16TC_w_f16TC");
}
at::Tensor y_RM_f16(at::Tensor A, at::Tensor B, bool right) {
  QuantArgs q;
  return gemm<false>(A, B, right, WKind::W16, q, 32, nullptr, nullptr, nullptr, "tinygemm_y_f16RM_x_f16RM_w_f16TC");
}

at::Tensor dequant_int4(at::Tensor in) {
  check_cuda(in);
  c10::cuda::CUDAGuard g(in.device());
  TORCH_CHECK(in.scalar_type() == at::kInt);
  TORCH_CHECK(in.dim() == 1 && in.is_contiguous());
  auto out = at::empty({in.numel() * 8}, in.options().dtype(at::kBFloat16));
  check_rc(tg_dequant_int4(in.data_ptr<int32_t>(), out.data_ptr(), in.numel(), cur_stream()), "tinygemm_dequant_int4");
  return out;
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(tinygemm, m) {
  m.def("convert_matrix_to_m16n8k16_A_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_to_m16n8k16_Aint4_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_to_m16n8k16_Aint8_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_from_m16n8k16_A_layout(Tensor t, int m, int k) -> Tensor");
  m.def("convert_matrix_to_m16n8k16_B_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_to_m16n8k16_Bint4_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_to_m16n8k16_Bint8_layout(Tensor t, int innerKTiles) -> Tensor");
  m.def("convert_matrix_from_m16n8k16_B_layout(Tensor t, int n, int k) -> Tensor");
  m.def(
      "tinygemm_y_f16TC_x_f16TC_w_int4TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, bool "
      "weightOnRight) -> Tensor");
  m.def(
      "tinygemm_y_f16RM_x_f16RM_w_int4TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, bool "
      "weightOnRight) -> Tensor");
  m.def(
      "tinygemm_y_f16TC_x_f16TC_w_any4TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, Tensor "
      "int4DequantValues, bool weightOnRight) -> Tensor");
  m.def(
      "tinygemm_y_f16RM_x_f16RM_w_any4TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, Tensor "
      "int4DequantValues, bool weightOnRight) -> Tensor");
  m.def(
      "tinygemm_y_f16TC_x_f16TC_w_mx4TC(Tensor A, Tensor B, int qGroupSize, Tensor mx4Exponents, bool weightOnRight) "
      "-> Tensor");
  m.def(
      "tinygemm_y_f16RM_x_f16RM_w_mx4TC(Tensor A, Tensor B, int qGroupSize, Tensor mx4Exponents, bool weightOnRight) "
      "-> Tensor");
  m.def(
      "tinygemm_y_f16TC_x_f16TC_w_int8TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, bool "
      "weightOnRight) -> Tensor");
  m.def(
      "tinygemm_y_f16RM_x_f16RM_w_int8TC(Tensor A, Tensor B, int qGroupSize, Tensor qScaleAndZeros, bool "
      "weightOnRight) -> Tensor");
  m.def("tinygemm_y_f16TC_x_f16TC_w_f16TC(Tensor A, Tensor B, bool weightOnRight) -> Tensor");
  m.def("tinygemm_y_f16RM_x_f16RM_w_f16TC(Tensor A, Tensor B, bool weightOnRight) -> Tensor");
  m.def("tinygemm_dequant_int4(Tensor t) -> Tensor");
}

// Like the reference (TinyGemm.cpp:124-200) the implementations are registered without a
// dispatch key: the ops validate device placement themselves and fail loudly on CPU tensors.
TORCH_LIBRARY_IMPL(tinygemm, CompositeExplicitAutograd, m) {
  m.impl("convert_matrix_to_m16n8k16_A_layout", to_A);
  m.impl("convert_matrix_to_m16n8k16_Aint4_layout", to_Aint4);
  m.impl("convert_matrix_to_m16n8k16_Aint8_layout", to_Aint8);
  m.impl("convert_matrix_from_m16n8k16_A_layout", from_A);
  m.impl("convert_matrix_to_m16n8k16_B_layout", to_B);
  m.impl("convert_matrix_to_m16n8k16_Bint4_layout", to_Bint4);
  m.impl("convert_matrix_to_m16n8k16_Bint8_layout", to_Bint8);
  m.impl("convert_matrix_from_m16n8k16_B_layout", from_B);
  m.impl("tinygemm_y_f16TC_x_f16TC_w_int4TC", y_TC_int4);
  m.impl("tinygemm_y_f16RM_x_f16RM_w_int4TC", y_RM_int4);
  m.impl("tinygemm_y_f16TC_x_f16TC_w_any4TC", y_TC_any4);
  m.impl("tinygemm_y_f16RM_x_f16RM_w_any4TC", y_RM_any4);
  m.impl("tinygemm_y_f16TC_x_f16TC_w_mx4TC", y_TC_mx4);
  m.impl("tinygemm_y_f16RM_x_f16RM_w_mx4TC", y_RM_mx4);
  m.impl("tinygemm_y_f16TC_x_f16TC_w_int8TC", y_TC_int8);
  m.impl("tinygemm_y_f16RM_x_f16RM_w_int8TC", y_RM_int8);
  m.impl("tinygemm_y_f16TC_x_f16TC_w_f16TC", y_TC_f16);
  m.impl("tinygemm_y_f16RM_x_f16RM_w_f16TC", y_RM_f16);
  m.impl("tinygemm_dequant_int4", dequant_int4);
}
