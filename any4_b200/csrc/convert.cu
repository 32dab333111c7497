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

// ---------------------------------------------------------------------------------------
// Packed A int4 layout -> packed B int4 layout of the same matrix (no counterpart in the reference).  The two layouts
// hold the same nibbles in another order: an A word = one k-tile of the row PAIR (g, g+8),
//   v0..v7 = (g,k0) (g,k0+1) (g+8,k0) (g+8,k0+1) (g,k0+8) (g,k0+9) (g+8,k0+8) (g+8,k0+9)      (TinyGemmConvertA.cu:248-278)
// a B word = two k-tiles (2j, 2j+1) of ONE row, v0..v3 = k0, k0+1, k0+8, k0+9 of tile 2j, v4..v7 of tile 2j+1
// (TinyGemmConvertB.cu:280-303); both pack v7 v5 v3 v1 v6 v4 v2 v0 from the top nibble down.  One thread per B word.
// Used for several activation rows against an A-layout weight (the A kernel takes one row per launch): repack once,
// mostly into L2, and run the one-pass tcgen05 kernel of the B layout.
// ---------------------------------------------------------------------------------------
__global__ void repack_Aint4_to_Bint4_kernel(const uint32_t* __restrict__ a, uint32_t* __restrict__ b, int64_t n_words,
                                             int k_tiles, int outer_a, int ik_a, int outer_b, int ik_b) {
  const int wpl = ik_b >> 1;  // B words per lane
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % wpl);
    const int t = (int)((i / wpl) & 31);
    const int64_t blk = i / (wpl * 32);
    const int ko_b = (int)(blk % outer_b);
    const int64_t nt = blk / outer_b;
    const int g = t >> 2, q = t & 3;
    const int64_t row = nt * 8 + g;
    const int64_t mt = row >> 4;
    const int rr = (int)(row & 15), ga = rr & 7, hi = rr >> 3;
    const int ta = 4 * ga + q;
    uint32_t word = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int T = (ko_b * wpl + j) * 2 + h;
      uint32_t w = 0;
      if (T < k_tiles) w = a[((mt * outer_a + T / ik_a) * 32 + ta) * ik_a + (T % ik_a)];
      w >>= 4 * hi;  // the low row's nibbles sit at bits 0, 8, 16, 24; the high row's 4 bits above
      // (k0, k0+1, k0+8, k0+9) = A nibbles at bits (0, 16, 8, 24) -> B nibble positions (0, 16, 4, 20) + 8 h
      word |= ((w & 0xfu) | (((w >> 16) & 0xfu) << 16) | (((w >> 8) & 0xfu) << 4) | (((w >> 24) & 0xfu) << 20)) << (8 * h);
    }
    b[i] = word;
  }
}

// the common case ik_a = ik_b = 4, k a multiple of 64: a lane's 16-byte A vector (4 k-tiles of its row pair) becomes the
// 8-byte B vectors of the SAME lane in the two n-tiles of the m-tile - one 16-byte load, two 8-byte stores, all coalesced
__global__ void repack_Aint4_to_Bint4_ik4_kernel(const uint4* __restrict__ a, uint2* __restrict__ b, int64_t n_vec, int outer) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (int64_t)gridDim.x * blockDim.x) {
    const uint4 w4 = a[v];
    const int t = (int)(v & 31);
    const int64_t blk = v >> 5;
    const int ko = (int)(blk % outer);
    const int64_t mt = blk / outer;
    auto half = [](uint32_t w0, uint32_t w1, int hi) {  // two k-tiles of one row -> one B word
      auto tile = [](uint32_t w) { return (w & 0xfu) | (((w >> 16) & 0xfu) << 16) | (((w >> 8) & 0xfu) << 4) | (((w >> 24) & 0xfu) << 20); };
      return tile(w0 >> (4 * hi)) | (tile(w1 >> (4 * hi)) << 8);
    };
#pragma unroll
    for (int hi = 0; hi < 2; ++hi)
      b[((2 * mt + hi) * outer + ko) * 32 + t] = make_uint2(half(w4.x, w4.y, hi), half(w4.z, w4.w, hi));
  }
}

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

int tg_repack_Aint4_to_Bint4(const int32_t* in, int32_t* out, int64_t rows, int64_t k, int ik_a, int ik_b, void* stream) {
  TG_REQUIRE(in && out && rows > 0 && k > 0, "tg_repack_Aint4_to_Bint4: bad arguments");
  TG_REQUIRE(rows % 16 == 0, "tg_repack_Aint4_to_Bint4: rows (%lld) must be the padded A-layout row count", (long long)rows);
  TG_REQUIRE(ik_a == 1 || ik_a == 2 || ik_a == 4, "tg_repack_Aint4_to_Bint4: A innerKTiles must be 1, 2 or 4 (got %d)", ik_a);
  TG_REQUIRE(ik_b == 2 || ik_b == 4 || ik_b == 8, "tg_repack_Aint4_to_Bint4: B innerKTiles must be 2, 4 or 8 (got %d)", ik_b);
  TG_REQUIRE(k % (ik_b * 16) == 0, "tg_repack_Aint4_to_Bint4: k (%lld) must be a multiple of %d", (long long)k, ik_b * 16);
  const int k_tiles = (int)(k / 16);
  const int outer_a = (int)div_up(k_tiles, ik_a), outer_b = k_tiles / ik_b;
  const int64_t n_words = (rows / 8) * outer_b * 32 * (ik_b / 2);
  if (ik_a == 4 && ik_b == 4 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const int64_t n_vec = (rows / 16) * outer_b * 32;
    repack_Aint4_to_Bint4_ik4_kernel<<<grid_for(n_vec), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint2*)out, n_vec,
                                                                                            outer_b);
    TG_CHECK_LAUNCH("tg_repack_Aint4_to_Bint4");
    return TG_OK;
  }
  repack_Aint4_to_Bint4_kernel<<<grid_for(n_words), kThreads, 0, (cudaStream_t)stream>>>(
      (const uint32_t*)in, (uint32_t*)out, n_words, k_tiles, outer_a, ik_a, outer_b, ik_b);
  TG_CHECK_LAUNCH("tg_repack_Aint4_to_Bint4");
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
