// Layout conversion kernels: row-major <-> m16n8k16 tensor-core fragment order.
//
// These replace the reference's eight `convert_matrix_*` ops and must be BIT-EXACT with
// them (reference: tinygemm_lib/TinyGemmConvertA.cu, TinyGemmConvertB.cu).  They run once
// per layer at load time (modules.py:197-205), i.e. off the decode path, and are pure
// HBM-bound gather/pack work.  Unlike the reference (one 32-thread CTA per tile through
// PackedTensorAccessor32) every kernel here is output-indexed: one thread produces one
// 16-byte vector (16-bit layouts) or one packed 32-bit word, consecutive threads write
// consecutive output addresses, and the grid is a grid-stride multiple of the SM count.
//
// Fragment geometry (lane t: g = t / 4, q = t % 4, k0 = kTile * 16 + 2 * q):
//   A fragment v0..v7: (g,k0) (g,k0+1) (g+8,k0) (g+8,k0+1) (g,k0+8) (g,k0+9) (g+8,k0+8) (g+8,k0+9)
//   B fragment v0..v3: (g,k0) (g,k0+1) (g,k0+8) (g,k0+9)
//   4-bit word: v7<<28 | v5<<24 | v3<<20 | v1<<16 | v6<<12 | v4<<8 | v2<<4 | v0
//   8-bit word: v3<<24 | v1<<16 | v2<<8 | v0
#include "common.cuh"

namespace tg {
namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work_items) {
  int64_t blocks = div_up(work_items, kThreads);
  const int64_t cap = 148 * 16;  // 16 resident 256-thread CTAs per SM is plenty for a gather
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ uint32_t ld_or_zero(const uint16_t* __restrict__ p, int64_t rows, int64_t cols,
                                               int64_t r, int64_t c) {
  return (r < rows && c < cols) ? (uint32_t)p[r * cols + c] : 0u;
}
__device__ __forceinline__ uint32_t ld_or_zero(const int32_t* __restrict__ p, int64_t rows, int64_t cols,
                                               int64_t r, int64_t c) {
  return (r < rows && c < cols) ? (uint32_t)p[r * cols + c] : 0u;
}

// value v (0..7) of the A fragment of lane t
__device__ __forceinline__ void a_frag_pos(int t, int v, int& dr, int& dc) {
  const int g = t >> 2, q = t & 3;
  dr = g + ((v >> 1) & 1) * 8;
  dc = 2 * q + (v & 1) + (v >> 2) * 8;
}
// value v (0..3) of the B fragment of lane t
__device__ __forceinline__ void b_frag_pos(int t, int v, int& dr, int& dc) {
  const int g = t >> 2, q = t & 3;
  dr = g;
  dc = 2 * q + (v & 1) + (v >> 1) * 8;
}

__device__ __forceinline__ uint32_t nib_shift(int v) {
  // v0->0 v1->16 v2->4 v3->20 v4->8 v5->24 v6->12 v7->28
  return (uint32_t)((v >> 1) * 4 + (v & 1) * 16);
}
__device__ __forceinline__ uint32_t byte_shift(int v) {
  // v0->0 v1->16 v2->8 v3->24
  return (uint32_t)((v >> 1) * 8 + (v & 1) * 16);
}

// ------------------------------------------------------------------------------------
// 16-bit layouts
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) to_A_kernel(const uint16_t* __restrict__ in, uint4* __restrict__ out,
                                                        int64_t m, int64_t k, int64_t mT, int64_t kT) {
  const int64_t total = mT * kT * 32;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int t = (int)(i & 31);
    const int64_t kt = (i >> 5) % kT;
    const int64_t mt = (i >> 5) / kT;
    uint32_t w[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int dr0, dc0, dr1, dc1;
      a_frag_pos(t, 2 * p, dr0, dc0);
      a_frag_pos(t, 2 * p + 1, dr1, dc1);
      const uint32_t lo = ld_or_zero(in, m, k, mt * 16 + dr0, kt * 16 + dc0);
      const uint32_t hi = ld_or_zero(in, m, k, mt * 16 + dr1, kt * 16 + dc1);
      w[p] = lo | (hi << 16);
    }
    out[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__global__ void __launch_bounds__(kThreads) from_A_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                          int64_t m, int64_t k, int64_t kT) {
  const int64_t total = m * k;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / k, c = i % k;
    const int64_t mt = r >> 4, kt = c >> 4;
    const int rr = (int)(r & 15), cc = (int)(c & 15);
    const int g = rr & 7, q = (cc & 7) >> 1;
    const int v = (cc >> 3) * 4 + (rr >> 3) * 2 + (cc & 1);
    out[i] = in[((mt * kT + kt) * 32 + (g * 4 + q)) * 8 + v];
  }
}

__global__ void __launch_bounds__(kThreads) to_B_kernel(const uint16_t* __restrict__ in, uint2* __restrict__ out,
                                                        int64_t n, int64_t k, int64_t nT, int64_t kO, int ik) {
  // one thread per (nt, ko, t, ki): writes 4 values = 8 bytes
  const int64_t total = nT * kO * 32 * ik;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int ki = (int)(i % ik);
    const int t = (int)((i / ik) & 31);
    const int64_t ko = (i / ik / 32) % kO;
    const int64_t nt = (i / ik / 32) / kO;
    const int64_t kt = ko * ik + ki;
    uint32_t w[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      int dr0, dc0, dr1, dc1;
      b_frag_pos(t, 2 * p, dr0, dc0);
      b_frag_pos(t, 2 * p + 1, dr1, dc1);
      const uint32_t lo = ld_or_zero(in, n, k, nt * 8 + dr0, kt * 16 + dc0);
      const uint32_t hi = ld_or_zero(in, n, k, nt * 8 + dr1, kt * 16 + dc1);
      w[p] = lo | (hi << 16);
    }
    out[i] = make_uint2(w[0], w[1]);
  }
}

__global__ void __launch_bounds__(kThreads) from_B_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out,
                                                          int64_t n, int64_t k, int64_t kO, int ik) {
  const int64_t total = n * k;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / k, c = i % k;
    const int64_t nt = r >> 3, kt = c >> 4;
    const int g = (int)(r & 7), cc = (int)(c & 15);
    const int q = (cc & 7) >> 1;
    const int v = (cc >> 3) * 2 + (cc & 1);
    const int64_t ko = kt / ik;
    const int ki = (int)(kt % ik);
    out[i] = in[(((nt * kO + ko) * 32 + (g * 4 + q)) * ik + ki) * 4 + v];
  }
}

// ------------------------------------------------------------------------------------
// packed integer layouts: one thread per output int32 word
// ------------------------------------------------------------------------------------
// A int4: out[mt][ks][t][i] packs the 8 A-fragment codes of k-tile ks*ik+i
__global__ void __launch_bounds__(kThreads) to_Aint4_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            int64_t m, int64_t k, int64_t mT, int64_t kS, int ik) {
  const int64_t total = mT * kS * 32 * ik;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int ii = (int)(i % ik);
    const int t = (int)((i / ik) & 31);
    const int64_t ks = (i / ik / 32) % kS;
    const int64_t mt = (i / ik / 32) / kS;
    const int64_t kt = ks * ik + ii;
    uint32_t w = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      int dr, dc;
      a_frag_pos(t, v, dr, dc);
      w |= ld_or_zero(in, m, k, mt * 16 + dr, kt * 16 + dc) << nib_shift(v);
    }
    out[i] = w;
  }
}

// A int8: out[mt][ko][t][i*2+j] packs A-fragment codes v[4j..4j+3] of k-tile ko*ik+i
__global__ void __launch_bounds__(kThreads) to_Aint8_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            int64_t m, int64_t k, int64_t mT, int64_t kO, int ik) {
  const int64_t total = mT * kO * 32 * ik * 2;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(i & 1);
    const int ii = (int)((i >> 1) % ik);
    const int t = (int)(((i >> 1) / ik) & 31);
    const int64_t ko = ((i >> 1) / ik / 32) % kO;
    const int64_t mt = ((i >> 1) / ik / 32) / kO;
    const int64_t kt = ko * ik + ii;
    uint32_t w = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int dr, dc;
      a_frag_pos(t, 4 * j + v, dr, dc);
      w |= ld_or_zero(in, m, k, mt * 16 + dr, kt * 16 + dc) << byte_shift(v);
    }
    out[i] = w;
  }
}

// B int4: out[nt][ks][t][j] packs B-fragment codes of k-tiles ks*ik+2j (v0-3) and ks*ik+2j+1 (v4-7)
__global__ void __launch_bounds__(kThreads) to_Bint4_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            int64_t n, int64_t k, int64_t nT, int64_t kS, int ik) {
  const int half = ik / 2;
  const int64_t total = nT * kS * 32 * half;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(i % half);
    const int t = (int)((i / half) & 31);
    const int64_t ks = (i / half / 32) % kS;
    const int64_t nt = (i / half / 32) / kS;
    uint32_t w = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      int dr, dc;
      b_frag_pos(t, v & 3, dr, dc);
      const int64_t kt = ks * ik + 2 * j + (v >> 2);
      w |= ld_or_zero(in, n, k, nt * 8 + dr, kt * 16 + dc) << nib_shift(v);
    }
    out[i] = w;
  }
}

// B int8: out[nt][ks][t][i] packs the 4 B-fragment codes of k-tile ks*ik+i
__global__ void __launch_bounds__(kThreads) to_Bint8_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            int64_t n, int64_t k, int64_t nT, int64_t kS, int ik) {
  const int64_t total = nT * kS * 32 * ik;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int ii = (int)(i % ik);
    const int t = (int)((i / ik) & 31);
    const int64_t ks = (i / ik / 32) % kS;
    const int64_t nt = (i / ik / 32) / kS;
    const int64_t kt = ks * ik + ii;
    uint32_t w = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int dr, dc;
      b_frag_pos(t, v, dr, dc);
      w |= ld_or_zero(in, n, k, nt * 8 + dr, kt * 16 + dc) << byte_shift(v);
    }
    out[i] = w;
  }
}

// debug: 8 x int4 -> 8 x bf16 (code - 8), order v0..v7
__global__ void __launch_bounds__(kThreads) dequant_int4_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out,
                                                                int64_t n_words) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * kThreads) {
    const uint32_t w = in[i];
    uint32_t o[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      // (v_{2p}, v_{2p+1}) sit at bits 4p and 16+4p; 0x4300 | c is the bf16 value 128 + c
      const uint32_t pair = ((w >> (4 * p)) & 0x000f000fu) | 0x43004300u;
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&pair);
      h = __hsub2(h, __float2bfloat162_rn(136.0f));  // exact: integers < 256
      o[p] = *reinterpret_cast<uint32_t*>(&h);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace
}  // namespace tg

using namespace tg;

extern "C" {

int tg_convert_to_A(const void* in, void* out, int64_t m, int64_t k, void* stream) {
  TG_REQUIRE(in && out && m >= 0 && k >= 0, "tg_convert_to_A: bad arguments");
  const int64_t mT = div_up(m, 16), kT = div_up(k, 16);
  if (mT * kT == 0) return TG_OK;
  to_A_kernel<<<grid_for(mT * kT * 32), kThreads, 0, (cudaStream_t)stream>>>((const uint16_t*)in, (uint4*)out, m, k, mT, kT);
  TG_CHECK_LAUNCH("tg_convert_to_A");
  return TG_OK;
}

int tg_convert_from_A(const void* in, void* out, int64_t m, int64_t k, void* stream) {
  TG_REQUIRE(in && out && m >= 0 && k >= 0, "tg_convert_from_A: bad arguments");
  if (m * k == 0) return TG_OK;
  from_A_kernel<<<grid_for(m * k), kThreads, 0, (cudaStream_t)stream>>>((const uint16_t*)in, (uint16_t*)out, m, k, div_up(k, 16));
  TG_CHECK_LAUNCH("tg_convert_from_A");
  return TG_OK;
}

int tg_convert_to_B(const void* in, void* out, int64_t n, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && n >= 0 && k >= 0, "tg_convert_to_B: bad arguments");
  TG_REQUIRE(ik == 1 || ik == 2, "tg_convert_to_B: innerKTiles must be 1 or 2 (got %d)", ik);
  const int64_t nT = div_up(n, 8), kO = div_up(k, 16 * ik);
  if (nT * kO == 0) return TG_OK;
  to_B_kernel<<<grid_for(nT * kO * 32 * ik), kThreads, 0, (cudaStream_t)stream>>>((const uint16_t*)in, (uint2*)out, n, k, nT, kO, ik);
  TG_CHECK_LAUNCH("tg_convert_to_B");
  return TG_OK;
}

int tg_convert_from_B(const void* in, void* out, int64_t n, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && n >= 0 && k >= 0, "tg_convert_from_B: bad arguments");
  TG_REQUIRE(ik == 1 || ik == 2, "tg_convert_from_B: innerKTiles must be 1 or 2 (got %d)", ik);
  if (n * k == 0) return TG_OK;
  from_B_kernel<<<grid_for(n * k), kThreads, 0, (cudaStream_t)stream>>>((const uint16_t*)in, (uint16_t*)out, n, k, div_up(k, 16 * ik), ik);
  TG_CHECK_LAUNCH("tg_convert_from_B");
  return TG_OK;
}

int tg_convert_to_Aint4(const int32_t* in, int32_t* out, int64_t m, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && m >= 0 && k >= 0, "tg_convert_to_Aint4: bad arguments");
  TG_REQUIRE(ik == 1 || ik == 2 || ik == 4, "tg_convert_to_Aint4: innerKTiles must be 1, 2 or 4 (got %d)", ik);
  const int64_t mT = div_up(m, 16), kS = div_up(k, 16 * ik);
  if (mT * kS == 0) return TG_OK;
  to_Aint4_kernel<<<grid_for(mT * kS * 32 * ik), kThreads, 0, (cudaStream_t)stream>>>(in, (uint32_t*)out, m, k, mT, kS, ik);
  TG_CHECK_LAUNCH("tg_convert_to_Aint4");
  return TG_OK;
}

int tg_convert_to_Aint8(const int32_t* in, int32_t* out, int64_t m, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && m >= 0 && k >= 0, "tg_convert_to_Aint8: bad arguments");
  TG_REQUIRE(ik == 1 || ik == 2, "tg_convert_to_Aint8: innerKTiles must be 1 or 2 (got %d)", ik);
  const int64_t mT = div_up(m, 16), kO = div_up(div_up(k, 16), ik);
  if (mT * kO == 0) return TG_OK;
  to_Aint8_kernel<<<grid_for(mT * kO * 32 * ik * 2), kThreads, 0, (cudaStream_t)stream>>>(in, (uint32_t*)out, m, k, mT, kO, ik);
  TG_CHECK_LAUNCH("tg_convert_to_Aint8");
  return TG_OK;
}

int tg_convert_to_Bint4(const int32_t* in, int32_t* out, int64_t n, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && n >= 0 && k >= 0, "tg_convert_to_Bint4: bad arguments");
  TG_REQUIRE(ik == 2 || ik == 4 || ik == 8, "tg_convert_to_Bint4: innerKTiles must be 2, 4 or 8 (got %d)", ik);
  TG_REQUIRE(k % (ik * 16) == 0, "tg_convert_to_Bint4: k (%lld) must be a multiple of innerKTiles*16 (%d)", (long long)k, ik * 16);
  const int64_t nT = div_up(n, 8), kS = k / (ik * 16);
  if (nT * kS == 0) return TG_OK;
  to_Bint4_kernel<<<grid_for(nT * kS * 32 * (ik / 2)), kThreads, 0, (cudaStream_t)stream>>>(in, (uint32_t*)out, n, k, nT, kS, ik);
  TG_CHECK_LAUNCH("tg_convert_to_Bint4");
  return TG_OK;
}

int tg_convert_to_Bint8(const int32_t* in, int32_t* out, int64_t n, int64_t k, int ik, void* stream) {
  TG_REQUIRE(in && out && n >= 0 && k >= 0, "tg_convert_to_Bint8: bad arguments");
  TG_REQUIRE(ik == 1 || ik == 2 || ik == 4, "tg_convert_to_Bint8: innerKTiles must be 1, 2 or 4 (got %d)", ik);
  TG_REQUIRE(k % (ik * 16) == 0, "tg_convert_to_Bint8: k (%lld) must be a multiple of innerKTiles*16 (%d)", (long long)k, ik * 16);
  const int64_t nT = div_up(n, 8), kS = k / (ik * 16);
  if (nT * kS == 0) return TG_OK;
  to_Bint8_kernel<<<grid_for(nT * kS * 32 * ik), kThreads, 0, (cudaStream_t)stream>>>(in, (uint32_t*)out, n, k, nT, kS, ik);
  TG_CHECK_LAUNCH("tg_convert_to_Bint8");
  return TG_OK;
}

int tg_dequant_int4(const int32_t* in, void* out, int64_t n_words, void* stream) {
  TG_REQUIRE(in && out && n_words >= 0, "tg_dequant_int4: bad arguments");
  if (n_words == 0) return TG_OK;
  dequant_int4_kernel<<<grid_for(n_words), kThreads, 0, (cudaStream_t)stream>>>((const uint32_t*)in, (uint4*)out, n_words);
  TG_CHECK_LAUNCH("tg_dequant_int4");
  return TG_OK;
}

}  // extern "C"
