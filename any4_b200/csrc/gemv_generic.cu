// Weight-only GEMM for int8 and 16-bit packed weights, in either layout (4-bit weights have their own kernels:
// gemv_w4_b.cu, gemv_w4_a.cu).
//
// Replaces tinygemm_m16n8k16_chunk_kernel instantiated with *Layout_TC_int8 / *Layout_TC
// (reference: TinyGemmImpl.cuh:23-345, MatrixLayoutA.cuh:211-373, :818-1062, MatrixLayoutB.cuh:461-684,
// :1103-1328, Dequantization.cuh:265-328).
//
// gemm_w8_ring_kernel (int8): producer-warp bulk-TMA ring, weights as the 16-row mma operand (a PAIR of B-layout tiles
// is one A fragment), group words travelling with the stage, byte-permute decode; see the comment above the kernel.
// gemm_stream_kernel (16-bit weights; int8 shapes the ring kernel does not take): the packed layouts are mma.m16n8k16
// fragment orders, and the words a lane owns for IK consecutive k-tiles are contiguous.  A CTA (8 warps splitting k)
// walks over row tiles (persistent grid), stages the activations once in shared memory, fetches U units per lane with
// 16-byte loads before decoding any of them, decodes int8 with bit patterns + one exact HSUB2 + the reference's
// single-rounded FMA, and feeds the words straight into mma.sync (fp32 accumulation of exact products).  Warps are
// combined through shared memory, one RN at the end.
// gemm_frag_kernel: the simple per-k-tile FFMA version, kept for k too long to stage the activations.
#include <cstdlib>

#include "common.cuh"
#include "w4_common.cuh"  // mma16816, mbarrier / bulk-copy helpers

namespace tg {
namespace w4 {
extern bool g_pdl, g_static_weights;  // tg_set_option (gemv_w4_b.cu)
}
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kActs = 8;  // activation rows per pass

enum Kind { W4 = 0, W8 = 1, W16 = 2 };

struct GParams {
  const uint32_t* w;      // packed weight (int32 words, or 16-bit pairs viewed as words)
  const uint16_t* x;      // [rows_x][k]
  uint16_t* y;            // [rows_x][w_rows]
  const uint16_t* sz;     // [k/g][w_rows][2] dtype
  const uint16_t* lut;    // W4 only
  const uint8_t* exps;    // W4 mx4 only, [w_rows][k/g]
  int lut_stride;
  int rows_x, w_rows, k, k_tiles, outer_k, ik, glog2;
  int is_mx4;
};

template <tg_dtype DT>
__device__ __forceinline__ float to_f32(uint16_t v) {
  if constexpr (DT == TG_BF16) return __uint_as_float((uint32_t)v << 16);
  return __half2float(__ushort_as_half(v));
}
template <tg_dtype DT>
__device__ __forceinline__ uint16_t from_f32(float f) {
  if constexpr (DT == TG_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  return __half_as_ushort(__float2half_rn(f));
}
// single-rounded v * s + z in the activation dtype
template <tg_dtype DT>
__device__ __forceinline__ uint16_t fma_dt(uint16_t v, uint16_t s, uint16_t z) {
  if constexpr (DT == TG_BF16) {
    return __bfloat16_as_ushort(__hfma(__ushort_as_bfloat16(v), __ushort_as_bfloat16(s), __ushort_as_bfloat16(z)));
  } else {
    return __half_as_ushort(__hfma(__ushort_as_half(v), __ushort_as_half(s), __ushort_as_half(z)));
  }
}
template <tg_dtype DT>
__device__ __forceinline__ uint16_t int_to_dt(int v) {  // exact for |v| <= 256
  return from_f32<DT>((float)v);
}
template <tg_dtype DT>
__device__ __forceinline__ uint16_t e8m0_dt(uint32_t e) {
  if constexpr (DT == TG_BF16) {
    if (e == 255u) return 0x7fc0;
    if (e == 0u) return 0x0040;
    return (uint16_t)(e << 7);
  } else {
    return __half_as_ushort(__float2half_rn(e == 255u ? __int_as_float(0x7fc00000) : exp2f((float)e - 127.0f)));
  }
}

// Decode the fragment values lane t holds for (row tile rt, k-tile kt) into `out`, as floats,
// in fragment order.  ALAYOUT: 8 values, rows {g, g, g+8, g+8, g, g, g+8, g+8}; B: 4 values, row g.
template <tg_dtype DT, Kind KIND, bool ALAYOUT>
__device__ __forceinline__ void decode(const GParams& p, int rt, int kt, int t, float* out) {
  constexpr int NV = ALAYOUT ? 8 : 4;
  const int g = t >> 2;
  uint16_t raw[NV];
  if constexpr (KIND == W16) {
    if constexpr (ALAYOUT) {
      const uint4 v = reinterpret_cast<const uint4*>(p.w)[((int64_t)rt * p.k_tiles + kt) * 32 + t];
      const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        raw[2 * i] = (uint16_t)(ww[i] & 0xffff);
        raw[2 * i + 1] = (uint16_t)(ww[i] >> 16);
      }
    } else {
      const int ko = kt / p.ik, ki = kt % p.ik;
      const uint2 v = reinterpret_cast<const uint2*>(p.w)[(((int64_t)rt * p.outer_k + ko) * 32 + t) * p.ik + ki];
      raw[0] = (uint16_t)(v.x & 0xffff);
      raw[1] = (uint16_t)(v.x >> 16);
      raw[2] = (uint16_t)(v.y & 0xffff);
      raw[3] = (uint16_t)(v.y >> 16);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) out[i] = to_f32<DT>(raw[i]);
    return;
  } else {
    int codes[NV];
    const int ko = kt / p.ik, ki = kt % p.ik;
    if constexpr (KIND == W4) {
      static_assert(ALAYOUT, "B-layout 4-bit weights use gemv_w4_b.cu");
      const uint32_t w = p.w[(((int64_t)rt * p.outer_k + ko) * 32 + t) * p.ik + ki];
#pragma unroll
      for (int i = 0; i < 8; ++i) codes[i] = (w >> ((i >> 1) * 4 + (i & 1) * 16)) & 0xf;
    } else {
      if constexpr (ALAYOUT) {
        const uint2 w = reinterpret_cast<const uint2*>(p.w)[(((int64_t)rt * p.outer_k + ko) * 32 + t) * p.ik + ki];
        const uint32_t ww[2] = {w.x, w.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) codes[i] = (ww[i >> 2] >> (((i & 3) >> 1) * 8 + (i & 1) * 16)) & 0xff;
      } else {
        const uint32_t w = p.w[(((int64_t)rt * p.outer_k + ko) * 32 + t) * p.ik + ki];
#pragma unroll
        for (int i = 0; i < 4; ++i) codes[i] = (w >> ((i >> 1) * 8 + (i & 1) * 16)) & 0xff;
      }
    }
    const int grp = (kt * 16) >> p.glog2;
    const int n_groups = p.k >> p.glog2;
#pragma unroll
    for (int h = 0; h < (ALAYOUT ? 2 : 1); ++h) {
      const int row = rt * (ALAYOUT ? 16 : 8) + g + h * 8;
      uint16_t s, z;
      if (KIND == W4 && p.is_mx4) {
        s = e8m0_dt<DT>(p.exps[(int64_t)row * n_groups + grp]);
        z = 0x8000;  // -0 keeps v * s exact including the sign of zero
      } else {
        const uint32_t szv = reinterpret_cast<const uint32_t*>(p.sz)[(int64_t)grp * p.w_rows + row];
        s = (uint16_t)(szv & 0xffff);
        z = (uint16_t)(szv >> 16);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool mine = ALAYOUT ? (((i >> 1) & 1) == h) : true;
        if (!mine) continue;
        uint16_t v;
        if constexpr (KIND == W4) {
          v = p.lut[(int64_t)row * p.lut_stride + codes[i]];
        } else {
          v = int_to_dt<DT>(codes[i] - 128);
        }
        out[i] = to_f32<DT>(fma_dt<DT>(v, s, z));
      }
    }
  }
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT>
__global__ void __launch_bounds__(kThreads) gemm_frag_kernel(const GParams p) {
  constexpr int NV = ALAYOUT ? 8 : 4;
  constexpr int ROWS = ALAYOUT ? 16 : 8;
  constexpr int RH = ALAYOUT ? 2 : 1;  // row halves per lane
  __shared__ float red[kWarps][kActs][ROWS];

  const int rt = blockIdx.x;
  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  const int g = t >> 2, q = t & 3;

  for (int a0 = 0; a0 < p.rows_x; a0 += kActs) {
    const int na = min(kActs, p.rows_x - a0);
    float acc[kActs][RH];
#pragma unroll
    for (int a = 0; a < kActs; ++a)
#pragma unroll
      for (int h = 0; h < RH; ++h) acc[a][h] = 0.f;

    for (int kt = warp; kt < p.k_tiles; kt += kWarps) {
      float w[NV];
      decode<DT, KIND, ALAYOUT>(p, rt, kt, t, w);
      const int kc = kt * 16 + 2 * q;
#pragma unroll
      for (int a = 0; a < kActs; ++a) {
        if (a < na) {
          const uint16_t* xr = p.x + (int64_t)(a0 + a) * p.k;
          float xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = kc + (j & 1) + (j >> 1) * 8;
            xv[j] = kk < p.k ? to_f32<DT>(xr[kk]) : 0.f;
          }
          if constexpr (ALAYOUT) {
            // v0,v1: (g, k0),(g, k0+1); v2,v3: g+8; v4,v5: (g, k0+8),(g, k0+9); v6,v7: g+8
            acc[a][0] = fmaf(w[0], xv[0], acc[a][0]);
            acc[a][0] = fmaf(w[1], xv[1], acc[a][0]);
            acc[a][0] = fmaf(w[4], xv[2], acc[a][0]);
            acc[a][0] = fmaf(w[5], xv[3], acc[a][0]);
            acc[a][1] = fmaf(w[2], xv[0], acc[a][1]);
            acc[a][1] = fmaf(w[3], xv[1], acc[a][1]);
            acc[a][1] = fmaf(w[6], xv[2], acc[a][1]);
            acc[a][1] = fmaf(w[7], xv[3], acc[a][1]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[a][0] = fmaf(w[j], xv[j], acc[a][0]);
          }
        }
      }
    }
    // the four q-lanes of a row
#pragma unroll
    for (int a = 0; a < kActs; ++a)
#pragma unroll
      for (int h = 0; h < RH; ++h) {
        float v = acc[a][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (q == 0) red[warp][a][g + 8 * h] = v;
      }
    __syncthreads();
    for (int i = threadIdx.x; i < na * ROWS; i += kThreads) {
      const int a = i / ROWS, r = i % ROWS;
      float s = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < kWarps; ++w2) s += red[w2][a][r];
      p.y[(int64_t)(a0 + a) * p.w_rows + rt * ROWS + r] = from_f32<DT>(s);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Streaming variant (int8 and 16-bit weights; the default).  Same decode arithmetic as above, restructured for
// memory-level parallelism: a lane's words of one "unit" (IK consecutive k-tiles) are contiguous in every packed
// layout, so they are fetched with 16-byte loads, U units (plus their scale/zero words) are in flight per lane
// before anything is decoded, and the activations are staged once per CTA in shared memory (zero-padded to whole
// units) instead of being re-read from global memory with bounds checks per k-tile.
// ---------------------------------------------------------------------------------------
constexpr int kMaxXSmem = 64 * 1024;

template <int N>
__device__ __forceinline__ void load_words(const uint32_t* __restrict__ src, uint32_t* dst) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      dst[4 * i] = v.x, dst[4 * i + 1] = v.y, dst[4 * i + 2] = v.z, dst[4 * i + 3] = v.w;
    }
  } else if constexpr (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(src) + i);
      dst[2 * i] = v.x, dst[2 * i + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = __ldg(src + i);
  }
}

// two int8 codes (bits 0..7 and 16..23 of `two`) -> (code - 128) * s + z for both, single-rounded in the activation
// dtype like fma_dt(), as a packed pair.  No int->float conversion instructions: bit patterns + one exact HSUB2.
template <tg_dtype DT>
__device__ __forceinline__ uint32_t decode8_pair(uint32_t two, uint32_t s2, uint32_t z2) {
  if constexpr (DT == TG_FP16) {
    const uint32_t h = (two & 0x00ff00ffu) | 0x64006400u;  // 1024 + b
    const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&h), __floats2half2_rn(1152.f, 1152.f));
    const __half2 w = __hfma2(v, *reinterpret_cast<const __half2*>(&s2), *reinterpret_cast<const __half2*>(&z2));
    return *reinterpret_cast<const uint32_t*>(&w);
  } else {
    // bf16 has 8 significant bits: t = 128 + (b & 127) and the offset 256 - (b & 128) are both exact, and so is
    // t - offset = b - 128
    const uint32_t tt = (two & 0x007f007fu) | 0x43004300u;
    const uint32_t off = 0x43804380u ^ (two & 0x00800080u);
    const __nv_bfloat162 v = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&tt), *reinterpret_cast<const __nv_bfloat162*>(&off));
    const __nv_bfloat162 w = __hfma2(v, *reinterpret_cast<const __nv_bfloat162*>(&s2),
                                     *reinterpret_cast<const __nv_bfloat162*>(&z2));
    return *reinterpret_cast<const uint32_t*>(&w);
  }
}

// The packed layouts ARE mma.m16n8k16 fragment orders, so the decoded words go straight into the tensor core:
//   A layout (weight on the left):  A operand = 16 weight rows x 16 k, B operand = activations (k x 8 rows)
//   B layout (weight on the right): B operand = 16 k x 8 weight rows,  A operand = activations (16 rows x k)
// fp32 accumulation of exact products, as in the reference.  HI: activation rows 8..15 of a pass exist (B layout).
template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK, bool HI>
#ifndef TG_STREAM_MINB
#define TG_STREAM_MINB 3  // 78 registers: three CTAs per SM (measured: int8 m = 1 14.4 -> 13.4 us, m = 16 53 -> 37 us at 4096^2)
#endif
__global__ void __launch_bounds__(kThreads, TG_STREAM_MINB) gemm_stream_kernel(const GParams p, int rows_per_pass, int kpad) {
  static_assert(KIND != W4, "4-bit weights have their own kernels");
  static_assert(!(ALAYOUT && HI), "the A layout carries at most 8 activation rows per mma");
  constexpr int ROWS = ALAYOUT ? 16 : 8;
  constexpr int RH = ALAYOUT ? 2 : 1;
  constexpr int NP = ALAYOUT ? 4 : 2;                                        // packed value pairs per lane per k-tile
  constexpr int WPT = KIND == W8 ? (ALAYOUT ? 2 : 1) : NP;                   // words per lane per k-tile
  constexpr int NW = IK * WPT;                                               // ... per unit, contiguous
  constexpr int U0 = (32 / NW) < (16 / (IK * RH)) ? (32 / NW) : (16 / (IK * RH));
  constexpr int U = U0 < 1 ? 1 : (U0 > 8 ? 8 : U0);                          // units in flight per lane
  extern __shared__ __align__(16) uint8_t xs_raw[];
  uint16_t* xs = reinterpret_cast<uint16_t*>(xs_raw);                        // [rows_per_pass][xstride]
  // row stride = kpad + 8 elements = 4 * odd words: the 8 activation rows x 4 k-pairs one operand load touches fall
  // into 32 distinct banks (a stride of kpad alone would put all 8 rows into the same 4 banks)
  const int xstride = kpad + 8;
  __shared__ float red[kWarps][4][32];

  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  const int g = t >> 2, q = t & 3;
  const int n_units = p.outer_k;  // units per row tile
  const int n_tiles = p.w_rows / ROWS;
  const uint32_t* szw = reinterpret_cast<const uint32_t*>(p.sz);
  const bool vec_ok = ((p.k & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  const bool one_group_per_unit = (1 << p.glog2) >= IK * 16;

  uint32_t raw[U][NW];
  uint32_t szv[U][IK][RH];
  // one batch = U units of row tile rt, starting at unit u0 (stride kWarps): all loads are issued before any use
  auto load_batch = [&](int rt, int u0) {
    const uint32_t* wrow = p.w + ((int64_t)rt * n_units * 32 + t) * NW;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int u = u0 + j * kWarps;
      if (u < n_units) {
        load_words<NW>(wrow + (int64_t)u * 32 * NW, raw[j]);
        if constexpr (KIND == W8) {
          const int n_groups = p.k >> p.glog2;
          if (one_group_per_unit) {
            const int grp = min((u * IK * 16) >> p.glog2, n_groups - 1);  // (padding tiles: any group)
#pragma unroll
            for (int h = 0; h < RH; ++h) {
              const uint32_t v = __ldg(szw + (int64_t)grp * p.w_rows + rt * ROWS + g + 8 * h);
#pragma unroll
              for (int ki = 0; ki < IK; ++ki) szv[j][ki][h] = v;
            }
          } else {
#pragma unroll
            for (int ki = 0; ki < IK; ++ki) {
              const int grp = min(((u * IK + ki) * 16) >> p.glog2, n_groups - 1);
#pragma unroll
              for (int h = 0; h < RH; ++h) szv[j][ki][h] = __ldg(szw + (int64_t)grp * p.w_rows + rt * ROWS + g + 8 * h);
            }
          }
        }
      }
    }
  };

  for (int a0 = 0; a0 < p.rows_x; a0 += rows_per_pass) {
    const int na = min(rows_per_pass, p.rows_x - a0);
    int rt = blockIdx.x;
    bool preloaded = false;
    if (rt < n_tiles) {  // the first weight batch is in flight while the activations are staged
      load_batch(rt, warp);
      preloaded = true;
    }
    // stage the activations of this pass: 16-byte pieces, zero beyond k
    for (int i = threadIdx.x; i < na * (kpad >> 3); i += kThreads) {
      const int a = i / (kpad >> 3), c = (i % (kpad >> 3)) * 8;
      const uint16_t* xr = p.x + (int64_t)(a0 + a) * p.k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c + 8 <= p.k && vec_ok) {
        v = *reinterpret_cast<const uint4*>(xr + c);
      } else {
        uint16_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = (c + j < p.k) ? xr[c + j] : (uint16_t)0;
        v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      }
      *reinterpret_cast<uint4*>(xs + (size_t)a * xstride + c) = v;
    }
    __syncthreads();
    // this lane's activation rows (g and, with HI, g + 8); rows beyond the pass contribute zeros
    const bool has_lo = g < na, has_hi = HI && (g + 8 < na);
    const uint16_t* x_lo = xs + (size_t)(has_lo ? g : 0) * xstride + 2 * q;
    const uint16_t* x_hi = xs + (size_t)(has_hi ? g + 8 : 0) * xstride + 2 * q;

    for (; rt < n_tiles; rt += gridDim.x) {  // persistent over row tiles: the staged activations are reused
      float acc[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;

      for (int u0 = warp; u0 < n_units; u0 += kWarps * U) {
        if (!preloaded) load_batch(rt, u0);
        preloaded = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int u = u0 + j * kWarps;
          if (u >= n_units) break;
#pragma unroll
          for (int ki = 0; ki < IK; ++ki) {
            uint32_t wv[NP];
            const uint32_t* rw = &raw[j][ki * WPT];
            if constexpr (KIND == W16) {
#pragma unroll
              for (int i = 0; i < NP; ++i) wv[i] = rw[i];
            } else {
              // fragment value i sits in byte (i>>1) + 2*(i&1) of its word: pairs (0,1) / (2,3) are bytes (0,2) / (1,3)
              uint32_t s2[RH], z2[RH];
#pragma unroll
              for (int h = 0; h < RH; ++h) {
                s2[h] = __byte_perm(szv[j][ki][h], 0, 0x1010);
                z2[h] = __byte_perm(szv[j][ki][h], 0, 0x3232);
              }
#pragma unroll
              for (int pr = 0; pr < NP; ++pr) {
                const uint32_t word = ALAYOUT ? rw[pr >> 1] : rw[0];
                wv[pr] = decode8_pair<DT>(word >> (8 * (pr & 1)), s2[ALAYOUT ? (pr & 1) : 0], z2[ALAYOUT ? (pr & 1) : 0]);
              }
            }
            const int kt = (u * IK + ki) * 16;
            const uint32_t xl0 = has_lo ? *reinterpret_cast<const uint32_t*>(x_lo + kt) : 0u;
            const uint32_t xl1 = has_lo ? *reinterpret_cast<const uint32_t*>(x_lo + kt + 8) : 0u;
            if constexpr (ALAYOUT) {
              // pairs: (g,k0..1) (g+8,k0..1) (g,k0+8..9) (g+8,k0+8..9) = a0..a3
              w4::mma16816<DT>(acc[(ki + j) & 1], wv[0], wv[1], wv[2], wv[3], xl0, xl1);
            } else {
              uint32_t xh0 = 0u, xh1 = 0u;
              if constexpr (HI) {
                xh0 = has_hi ? *reinterpret_cast<const uint32_t*>(x_hi + kt) : 0u;
                xh1 = has_hi ? *reinterpret_cast<const uint32_t*>(x_hi + kt + 8) : 0u;
              }
              w4::mma16816<DT>(acc[(ki + j) & 1], xl0, xh0, xl1, xh1, wv[0], wv[1]);
            }
          }
        }
      }
      // the next row tile's first batch is in flight during the reduction
      if (rt + (int)gridDim.x < n_tiles) {
        load_batch(rt + (int)gridDim.x, warp);
        preloaded = true;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) red[warp][i][t] = acc[0][i] + acc[1][i];
      __syncthreads();
      if (threadIdx.x < 128) {
        const int ci = threadIdx.x >> 5, gl = (threadIdx.x & 31) >> 2, ql = threadIdx.x & 3;
        float sum = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kWarps; ++w2) sum += red[w2][ci][threadIdx.x & 31];
        // C fragment: c0,c1 = (row g, cols 2q, 2q+1), c2,c3 = (row g+8, ...)
        const int crow = gl + 8 * (ci >> 1), ccol = 2 * ql + (ci & 1);
        const int act = ALAYOUT ? ccol : crow, wrow = ALAYOUT ? crow : ccol;
        if (act < na) p.y[(int64_t)(a0 + act) * p.w_rows + rt * ROWS + wrow] = from_f32<DT>(sum);
      }
      __syncthreads();
    }
    __syncthreads();  // everyone is done with the staged activations before the next pass overwrites them
  }
}

// ---------------------------------------------------------------------------------------
// Ring variant for int8 and 16-bit weights, up to 16 activation rows per pass (the decode case; both layouts; HI: rows
// 8..15 ride on a second mma per decoded fragment).
//  * The weight is always the mma's 16-row A operand and the activations its 8-column B operand: the A layout is that
//    fragment already; in the B layout a lane's words of two ADJACENT 8-row tiles are exactly (a0, a2) and (a1, a3) of
//    a 16-row fragment, so a pair of tiles is processed together - one mma and one pair of activation loads per 16
//    rows x 16 k, no zero operands (the stream kernel spends an mma per 8 rows and zero-fills half its A operand).
//  * A row tile's packed words are one contiguous run, so a producer warp streams them as bulk-TMA copies (16 KiB per
//    stage: 1024 k of 16 rows) into a shared-memory ring: bytes in flight do not depend on registers or occupancy, and
//    the ring keeps filling across row tiles while the consumers reduce.  The (scale, zero) words of those 1024 k
//    travel in the same stage: the producer fetches them one item ahead into registers and stores them next to the
//    chunk before it arms the barrier.
//  * Eight consumer warps take the chunk's units round-robin (conflict-free 16-byte shared loads) and decode a word
//    (two fragment pairs) with byte permutes instead of shifts and masks.  One block barrier per row tile
//    (double-buffered partial sums).
// PDL: dependents are released at entry; the weight stream starts before `griddepcontrol.wait` when the caller
// declared the weights static (TG_OPT_STATIC_WEIGHTS), the activations are always read after it.
// ---------------------------------------------------------------------------------------
constexpr int kRingChunk = 16 * 1024;                            // packed-weight bytes per stage
constexpr int kRingSzWords = 512;                                // (scale, zero) words per stage (group 32: 32 groups x 16 rows)
constexpr int kRingStageBytes = kRingChunk + kRingSzWords * 4;
constexpr int kRingThreads = kThreads + 32;                      // + producer warp
constexpr int kRingCtrl = 128;                                   // barriers
constexpr int kRingRedBytes = 2 * kWarps * 4 * 32 * 4;
constexpr int kRingMaxStages = 4;
constexpr int kRingMaxX = 144 * 1024;                            // activation area (one CTA per SM beyond ~30 KiB)
constexpr int ring_fixed_bytes(int stages) { return kRingCtrl + stages * kRingStageBytes + kRingRedBytes; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

template <int N>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t* dst) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const uint4 v = w4::lds128(addr + 16 * i);
      dst[4 * i] = v.x, dst[4 * i + 1] = v.y, dst[4 * i + 2] = v.z, dst[4 * i + 3] = v.w;
    }
  } else if constexpr (N == 2) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(dst[0]), "=r"(dst[1]) : "r"(addr));
  } else {
    static_assert(N == 1, "1, 2 or a multiple of 4 words");
    dst[0] = w4::lds32(addr);
  }
}

// One packed word = two fragment pairs: bytes (0, 2) -> `lo`, bytes (1, 3) -> `hi`; each value (code - 128) * s + z,
// single-rounded in the activation dtype (the arithmetic of decode8_pair, with PRMT doing extraction and exponent
// insertion at once).
template <tg_dtype DT>
__device__ __forceinline__ void decode8_word(uint32_t word, uint32_t s_lo, uint32_t z_lo, uint32_t s_hi, uint32_t z_hi,
                                             uint32_t& lo, uint32_t& hi) {
  if constexpr (DT == TG_FP16) {
    const uint32_t h0 = __byte_perm(word, 0x64646464u, 0x4240), h1 = __byte_perm(word, 0x64646464u, 0x4341);  // 1024 + b
    const __half2 off = __floats2half2_rn(1152.f, 1152.f);
    const __half2 v0 = __hsub2(*reinterpret_cast<const __half2*>(&h0), off);
    const __half2 v1 = __hsub2(*reinterpret_cast<const __half2*>(&h1), off);
    const __half2 r0 = __hfma2(v0, *reinterpret_cast<const __half2*>(&s_lo), *reinterpret_cast<const __half2*>(&z_lo));
    const __half2 r1 = __hfma2(v1, *reinterpret_cast<const __half2*>(&s_hi), *reinterpret_cast<const __half2*>(&z_hi));
    lo = *reinterpret_cast<const uint32_t*>(&r0);
    hi = *reinterpret_cast<const uint32_t*>(&r1);
  } else {
    // t = 128 + (b & 127) and offset = 256 - (b & 128) are exact in bf16, and so is t - offset = b - 128
    const uint32_t m = word & 0x7f7f7f7fu, nh = ~word & 0x80808080u;
    const uint32_t t0 = __byte_perm(m, 0x43434343u, 0x4240), t1 = __byte_perm(m, 0x43434343u, 0x4341);
    const uint32_t o0 = __byte_perm(nh, 0x43434343u, 0x4240), o1 = __byte_perm(nh, 0x43434343u, 0x4341);
    const __nv_bfloat162 v0 = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t0), *reinterpret_cast<const __nv_bfloat162*>(&o0));
    const __nv_bfloat162 v1 = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&t1), *reinterpret_cast<const __nv_bfloat162*>(&o1));
    const __nv_bfloat162 r0 = __hfma2(v0, *reinterpret_cast<const __nv_bfloat162*>(&s_lo), *reinterpret_cast<const __nv_bfloat162*>(&z_lo));
    const __nv_bfloat162 r1 = __hfma2(v1, *reinterpret_cast<const __nv_bfloat162*>(&s_hi), *reinterpret_cast<const __nv_bfloat162*>(&z_hi));
    lo = *reinterpret_cast<const uint32_t*>(&r0);
    hi = *reinterpret_cast<const uint32_t*>(&r1);
  }
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK, bool HI>  // HI: activation rows 8..15 of the pass exist (second mma per fragment)
__global__ void __launch_bounds__(kRingThreads, HI ? 1 : 2) gemm_w8_ring_kernel(const GParams p, int kpad, int static_w, int stages) {
  static_assert(KIND == W8 || KIND == W16, "int8 or 16-bit weights");
  constexpr int ROWS = 16;                       // weight rows per item: one A-layout tile, or two adjacent B-layout tiles
  constexpr int NT = ALAYOUT ? 1 : 2;            // packed tiles per item
  constexpr int WKT = (KIND == W8 ? 1 : 2) * (ALAYOUT ? 2 : 1);  // words per lane per k-tile of one packed tile
  constexpr int NWT = WKT * IK;                  // ... per unit
  constexpr int SUB = kRingChunk / NT;           // bytes of one packed tile per stage
  constexpr int UC = SUB / (128 * NWT);          // units per chunk
  constexpr int UPW = UC / kWarps;               // units per warp per chunk
  constexpr int CK = UC * IK * 16;               // k per chunk: 1024 (int8) / 512 (16-bit)
  static_assert(UPW >= 1 && UPW * NT * NWT == 16 && CK == (KIND == W8 ? 1024 : 512), "a warp owns 16 words per lane per chunk");
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t base = w4::smem_u32(ring_smem);
  const uint32_t bar_full = base, bar_empty = base + 8 * kRingMaxStages, stage0 = base + kRingCtrl;
  float* red = reinterpret_cast<float*>(ring_smem + kRingCtrl + stages * kRingStageBytes);      // [2][kWarps][4][32]
  uint16_t* xs = reinterpret_cast<uint16_t*>(ring_smem + kRingCtrl + stages * kRingStageBytes + kRingRedBytes);  // [rows_x][kpad + 8]

  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      w4::mbar_init(bar_full + 8 * s, 32);        // every producer lane (each stored group words)
      w4::mbar_init(bar_empty + 8 * s, kThreads);  // every consumer thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();

  const int n_units = p.outer_k, n_items = p.w_rows / ROWS;
  const int n_chunks = (n_units + UC - 1) / UC;
  const int n_groups = p.k >> p.glog2;
  const int gc = CK >> p.glog2;  // groups per chunk
  const uint32_t* szw = reinterpret_cast<const uint32_t*>(p.sz);

  if (warp == kWarps) {  // ---- producer ----
    if (!static_w) asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint64_t pol = w4::l2_evict_first_policy();
    // group words per lane per chunk: 16 / 8 / 4 / 2 for groups of 32 / 64 / 128 / 256 (none for 16-bit weights)
    const int nj = KIND == W8 ? (gc * ROWS) >> 5 : 0;
    uint32_t szr[16];
    auto load_sz = [&](int rt, int c) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = t + 32 * j;
        const int grp = c * gc + i / ROWS;
        szr[j] = (j < nj && grp < n_groups) ? __ldg(szw + (int64_t)grp * p.w_rows + rt * ROWS + (i % ROWS)) : 0u;
      }
    };
    int rt = blockIdx.x, it = 0, s = 0;
    uint32_t ph = 1;  // first pass over the ring: the slots are free
    if (rt < n_items) load_sz(rt, 0);
    for (; rt < n_items; rt += gridDim.x) {
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const uint32_t st = stage0 + s * kRingStageBytes;
        w4::mbar_wait(bar_empty + 8 * s, ph);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nj) w4::sts32(st + kRingChunk + (t + 32 * j) * 4, szr[j]);
        if (t == 0) {
          const int u0 = c * UC;
          const int nu = min(UC, n_units - u0);
          const uint32_t bytes = (uint32_t)nu * 128u * NWT;
          w4::mbar_expect_tx(bar_full + 8 * s, bytes * NT);
#pragma unroll
          for (int h = 0; h < NT; ++h)
            w4::bulk_g2s(st + h * SUB, p.w + ((int64_t)(rt * NT + h) * n_units + u0) * 32 * NWT, bytes, bar_full + 8 * s, pol);
        } else {
          mbar_arrive(bar_full + 8 * s);  // release: this lane's group words are visible with the chunk
        }
        int nrt = rt, nc = c + 1;
        if (nc == n_chunks) nc = 0, nrt = rt + (int)gridDim.x;
        if (nrt < n_items) load_sz(nrt, nc);
        if (++s == stages) s = 0, ph ^= 1;
      }
    }
    return;
  }

  // ---- consumers ----
  const int g = t >> 2, q = t & 3;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  {
    const bool vec_ok = ((p.k & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
    const int xstride = kpad + 8;  // 4 * odd words: the 8 rows x 4 k-pairs one operand load touches fall into 32 banks
    for (int i = threadIdx.x; i < p.rows_x * (kpad >> 3); i += kThreads) {
      const int a = i / (kpad >> 3), c = (i % (kpad >> 3)) * 8;
      const uint16_t* xr = p.x + (int64_t)a * p.k;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c + 8 <= p.k && vec_ok) {
        v = *reinterpret_cast<const uint4*>(xr + c);
      } else {
        uint16_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = (c + j < p.k) ? xr[c + j] : (uint16_t)0;
        v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      }
      *reinterpret_cast<uint4*>(xs + (size_t)a * xstride + c) = v;
    }
  }
  consumer_barrier();
  const int na = p.rows_x;
  const bool has_x = g < na;  // this lane's activation row (B operand column g)
  const uint32_t x_lane = w4::smem_u32(xs) + ((has_x ? g : 0) * (kpad + 8) + 2 * q) * 2;
  const bool has_xh = HI && g + 8 < na;  // ... and row g + 8 of a 9..16-row pass
  const uint32_t x_lane_h = w4::smem_u32(xs) + ((has_xh ? g + 8 : 0) * (kpad + 8) + 2 * q) * 2;
  const bool one_group_per_unit = (1 << p.glog2) >= IK * 16;

  int s = 0, buf = 0;
  uint32_t ph = 0;
  for (int rt = blockIdx.x; rt < n_items; rt += gridDim.x, buf ^= 1) {
    float acc[2][4], acch[2][4];  // (acch: rows 8..15 of the pass, HI only)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = 0.f, acch[c][i] = 0.f;

    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t st = stage0 + s * kRingStageBytes;
      const int u0 = c * UC;
      const int nu = min(UC, n_units - u0);
      const int gmax = min(gc, n_groups - c * gc) - 1;
      w4::mbar_wait(bar_full + 8 * s, ph);
      uint32_t raw[UPW][NT * NWT];
#pragma unroll
      for (int j = 0; j < UPW; ++j) {
        const int ul = warp + j * kWarps;
        if (ul < nu) {
#pragma unroll
          for (int h = 0; h < NT; ++h) lds_words<NWT>(st + h * SUB + (uint32_t)(ul * 32 + t) * NWT * 4, &raw[j][h * NWT]);
        }
      }
#pragma unroll
      for (int j = 0; j < UPW; ++j) {
        const int ul = warp + j * kWarps;
        if (ul < nu) {
          uint32_t s2[2], z2[2];
          auto group_words = [&](int k_local) {
            const int grp = min(k_local >> p.glog2, gmax);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t v = w4::lds32(st + kRingChunk + (grp * ROWS + g + 8 * h) * 4);
              s2[h] = __byte_perm(v, 0, 0x1010);
              z2[h] = __byte_perm(v, 0, 0x3232);
            }
          };
          if (KIND == W8 && one_group_per_unit) group_words(ul * IK * 16);
#pragma unroll
          for (int ki = 0; ki < IK; ++ki) {
            if (KIND == W8 && !one_group_per_unit) group_words((ul * IK + ki) * 16);
            uint32_t a0, a1, a2, a3;
            if constexpr (KIND == W16) {
              if constexpr (ALAYOUT) {  // the four words ARE the fragment
                a0 = raw[j][4 * ki], a1 = raw[j][4 * ki + 1], a2 = raw[j][4 * ki + 2], a3 = raw[j][4 * ki + 3];
              } else {                  // (b0, b1) of the first tile: row g at (k, k + 8); of the second tile: row g + 8
                a0 = raw[j][2 * ki], a2 = raw[j][2 * ki + 1];
                a1 = raw[j][NWT + 2 * ki], a3 = raw[j][NWT + 2 * ki + 1];
              }
            } else if constexpr (ALAYOUT) {  // word 0: (row g, row g + 8) at k 2q..2q+1, word 1: the same rows at k + 8
              decode8_word<DT>(raw[j][2 * ki], s2[0], z2[0], s2[1], z2[1], a0, a1);
              decode8_word<DT>(raw[j][2 * ki + 1], s2[0], z2[0], s2[1], z2[1], a2, a3);
            } else {                  // a word of the first tile: row g at (k, k + 8); of the second: row g + 8
              decode8_word<DT>(raw[j][ki], s2[0], z2[0], s2[0], z2[0], a0, a2);
              decode8_word<DT>(raw[j][IK + ki], s2[1], z2[1], s2[1], z2[1], a1, a3);
            }
            const uint32_t xa = x_lane + (uint32_t)((u0 + ul) * IK + ki) * 32;
            const uint32_t x0 = has_x ? w4::lds32(xa) : 0u;
            const uint32_t x1 = has_x ? w4::lds32(xa + 16) : 0u;
            w4::mma16816<DT>(acc[(ki + j) & 1], a0, a1, a2, a3, x0, x1);
            if constexpr (HI) {
              const uint32_t xh = x_lane_h + (uint32_t)((u0 + ul) * IK + ki) * 32;
              const uint32_t x2 = has_xh ? w4::lds32(xh) : 0u;
              const uint32_t x3 = has_xh ? w4::lds32(xh + 16) : 0u;
              w4::mma16816<DT>(acch[(ki + j) & 1], a0, a1, a2, a3, x2, x3);
            }
          }
        }
      }
      mbar_arrive(bar_empty + 8 * s);
      if (++s == stages) s = 0, ph ^= 1;
    }

    // cross-warp sums, rows 0..7 and (HI) 8..15 of the pass one after the other.  The scratch is double-buffered: the
    // buffer written two reductions ago was read before the barrier of the previous one.
#pragma unroll
    for (int half = 0; half < (HI ? 2 : 1); ++half) {
      float* rb = red + buf * (kWarps * 4 * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) rb[(warp * 4 + i) * 32 + t] = half == 0 ? acc[0][i] + acc[1][i] : acch[0][i] + acch[1][i];
      consumer_barrier();
      if (threadIdx.x < 128) {
        // C fragment: c0, c1 = (weight row g, activation rows 2q, 2q + 1), c2, c3 = (weight row g + 8, ...)
        const int ci = threadIdx.x >> 5;
        float sum = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kWarps; ++w2) sum += rb[(w2 * 4 + ci) * 32 + t];
        const int wrow = g + 8 * (ci >> 1), act = 2 * q + (ci & 1) + 8 * half;
        if (act < na) p.y[(int64_t)act * p.w_rows + rt * ROWS + wrow] = from_f32<DT>(sum);
      }
      if (half + 1 < (HI ? 2 : 1)) buf ^= 1;
    }
  }
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT>
int launch_simple(const GParams& p, cudaStream_t st) {
  const int tiles = p.w_rows / (ALAYOUT ? 16 : 8);
  gemm_frag_kernel<DT, KIND, ALAYOUT><<<tiles, kThreads, 0, st>>>(p);
  TG_CHECK_LAUNCH("gemm_frag_kernel");
  return TG_OK;
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK, bool HI>
int launch_stream_a(const GParams& p, int rows_per_pass, int kpad, cudaStream_t st) {
  const size_t smem = (size_t)rows_per_pass * (kpad + 8) * 2;
  auto kern = gemm_stream_kernel<DT, KIND, ALAYOUT, IK, HI>;
  static thread_local int ctas_per_sm_dev[kMaxDevices] = {}, n_sm = 0;
  int& ctas_per_sm = ctas_per_sm_dev[current_device_slot()];
  if (ctas_per_sm == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxXSmem) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_stream_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    int occ = 0;  // with a typical 8-16 KiB of staged activations
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 16 * 1024) != cudaSuccess || occ <= 0) occ = 2;
    ctas_per_sm = occ;
  }
  const int tiles = p.w_rows / (ALAYOUT ? 16 : 8);
  const int slots = ctas_per_sm * n_sm;  // persistent: the resident CTAs walk over the row tiles
  const int grid = tiles < slots ? tiles : slots;
  kern<<<grid, kThreads, smem, st>>>(p, rows_per_pass, kpad);
  TG_CHECK_LAUNCH("gemm_stream_kernel");
  return TG_OK;
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK>
int launch_stream(const GParams& p, cudaStream_t st) {
  const int kpad = p.outer_k * IK * 16;
  int rows_per_pass = kMaxXSmem / ((kpad + 8) * 2);
  if (rows_per_pass < 1) return launch_simple<DT, KIND, ALAYOUT>(p, st);  // very long k: activations stay in global memory
  const int cap = ALAYOUT ? 8 : 16;  // activation rows one mma carries
  if (rows_per_pass > cap) rows_per_pass = cap;
  if (rows_per_pass > p.rows_x) rows_per_pass = p.rows_x;
  if constexpr (!ALAYOUT) {
    if (rows_per_pass > 8) return launch_stream_a<DT, KIND, ALAYOUT, IK, true>(p, rows_per_pass, kpad, st);
  }
  return launch_stream_a<DT, KIND, ALAYOUT, IK, false>(p, rows_per_pass, kpad, st);
}

// int8 / 16-bit weights through the ring kernel, in passes of up to 16 activation rows (the weight of a second pass
// usually comes from L2): TG_OK / TG_ERR_UNSUPPORTED (use the stream kernel)
template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK, bool HI>
int launch_ring_pass(const GParams& p, int kpad, size_t xbytes, int ctas_env, cudaStream_t st) {
  auto kern = gemm_w8_ring_kernel<DT, KIND, ALAYOUT, IK, HI>;
  static thread_local int ready[kMaxDevices] = {}, n_sm = 0;
  int& rdy = ready[current_device_slot()];
  if (!rdy) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_fixed_bytes(kRingMaxStages) + kRingMaxX) !=
        cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_w8_ring_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    rdy = 1;
  }
  const int items = p.w_rows / 16;
  // two CTAs per SM while four stages + the activations fit 113 KiB, one (with the whole 227 KiB) beyond that
  int per_sm = ctas_env > 0 ? ctas_env : (ring_fixed_bytes(kRingMaxStages) + xbytes <= 113 * 1024 ? 2 : 1);
  if (HI) per_sm = 1;
  const int slots = per_sm * n_sm;  // persistent: the resident CTAs walk over the row tiles
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(items < slots ? items : slots), 1, 1);
  cfg.blockDim = dim3(kRingThreads, 1, 1);
  cfg.dynamicSmemBytes = ring_fixed_bytes(kRingMaxStages) + xbytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = w4::g_pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, kpad, (int)(w4::g_static_weights ? 1 : 0), (int)kRingMaxStages);
  if (e != cudaSuccess) {
    set_error("gemm_w8_ring_kernel launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TG_ERR_CUDA;
  }
  count_launch();
  return TG_OK;
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT, int IK>
int launch_ring(const GParams& p0, cudaStream_t st) {
  static const int enabled = [] { const char* e = getenv("TG_W8_RING"); return e ? atoi(e) : 1; }();
  static const int ctas_env = [] { const char* e = getenv("TG_W8_CTAS"); return e ? atoi(e) : 0; }();
  static const int rows_env = [] { const char* e = getenv("TG_W8_ROWS"); return e ? atoi(e) : 16; }();  // rows per pass (tuning)
  const int kpad = p0.outer_k * IK * 16;
  const size_t row_bytes = (size_t)(kpad + 8) * 2;
  int per_pass = (int)((size_t)kRingMaxX / row_bytes);
  const int cap = rows_env >= 1 && rows_env <= 16 ? rows_env : 16;
  if (per_pass > cap) per_pass = cap;
  if (!enabled || per_pass < 1 || (p0.w_rows & 15) != 0 || (reinterpret_cast<uintptr_t>(p0.w) & 15) != 0 ||
      (KIND == W8 && (reinterpret_cast<uintptr_t>(p0.sz) & 3) != 0))
    return TG_ERR_UNSUPPORTED;
  for (int r0 = 0; r0 < p0.rows_x; r0 += per_pass) {
    GParams p = p0;
    p.rows_x = p0.rows_x - r0 < per_pass ? p0.rows_x - r0 : per_pass;
    p.x = p0.x + (int64_t)r0 * p0.k;
    p.y = p0.y + (int64_t)r0 * p0.w_rows;
    const size_t xbytes = (size_t)p.rows_x * row_bytes;
    const int rc = p.rows_x > 8 ? launch_ring_pass<DT, KIND, ALAYOUT, IK, true>(p, kpad, xbytes, ctas_env, st)
                                : launch_ring_pass<DT, KIND, ALAYOUT, IK, false>(p, kpad, xbytes, ctas_env, st);
    if (rc != TG_OK) return rc;
  }
  return TG_OK;
}

template <tg_dtype DT, Kind KIND, bool ALAYOUT>
int launch(const GParams& p, cudaStream_t st) {
  if constexpr (KIND == W8 || KIND == W16) {
    // the inner-k values the layouts exist in (capi.cu): int8 1, 2 and - B layout only - 4; 16-bit 1 and - B layout only - 2
    int rc = TG_ERR_UNSUPPORTED;
    switch (KIND == W16 && ALAYOUT ? 1 : p.ik) {
      case 1: rc = launch_ring<DT, KIND, ALAYOUT, 1>(p, st); break;
      case 2:
        if constexpr (KIND == W8 || !ALAYOUT) rc = launch_ring<DT, KIND, ALAYOUT, 2>(p, st);
        break;
      case 4:
        if constexpr (KIND == W8 && !ALAYOUT) rc = launch_ring<DT, KIND, ALAYOUT, 4>(p, st);
        break;
    }
    if (rc != TG_ERR_UNSUPPORTED) return rc;
  }
  if (KIND == W16 && ALAYOUT) return launch_stream<DT, KIND, ALAYOUT, 1>(p, st);  // no inner k in this layout
  switch (p.ik) {
    case 1: return launch_stream<DT, KIND, ALAYOUT, 1>(p, st);
    case 2: return launch_stream<DT, KIND, ALAYOUT, 2>(p, st);
    case 4: return launch_stream<DT, KIND, ALAYOUT, 4>(p, st);
    case 8: return launch_stream<DT, KIND, ALAYOUT, 8>(p, st);
  }
  return launch_simple<DT, KIND, ALAYOUT>(p, st);
}

template <Kind KIND>
int dispatch(const GParams& p, tg_weight_side side, tg_dtype dt, cudaStream_t st) {
  {
    if (side == TG_WEIGHT_A) return dt == TG_BF16 ? launch<TG_BF16, KIND, true>(p, st) : launch<TG_FP16, KIND, true>(p, st);
    return dt == TG_BF16 ? launch<TG_BF16, KIND, false>(p, st) : launch<TG_FP16, KIND, false>(p, st);
  }
}

int glog2_of(int group) { return group == 32 ? 5 : group == 64 ? 6 : group == 128 ? 7 : 8; }

}  // namespace

int launch_gemm_w8_rm(void* y, const void* x, const int32_t* w, const void* sz, int64_t rows_x, int64_t w_rows,
                      int64_t k, int group, int ik, tg_weight_side side, tg_dtype dt, cudaStream_t st) {
  GParams p{};
  p.w = reinterpret_cast<const uint32_t*>(w);
  p.x = (const uint16_t*)x;
  p.y = (uint16_t*)y;
  p.sz = (const uint16_t*)sz;
  p.rows_x = (int)rows_x;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.k_tiles = (int)div_up(k, 16);
  p.ik = ik;
  p.outer_k = (int)div_up(p.k_tiles, ik);
  p.glog2 = glog2_of(group);
  return dispatch<W8>(p, side, dt, st);
}

// 4-bit weights in the A layout through the simple per-k-tile kernel: the fallback of gemv_w4_a.cu for shapes its
// staging areas cannot hold (very long k at small group sizes); any k, up to 8 activation rows per pass
int launch_gemm_w4_rm_A_generic(void* y, const void* x, const int32_t* w, const void* sz, const void* lut, const uint8_t* exps,
                                int lut_stride, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik, bool mx4,
                                tg_dtype dt, cudaStream_t st) {
  GParams p{};
  p.w = reinterpret_cast<const uint32_t*>(w);
  p.x = (const uint16_t*)x;
  p.y = (uint16_t*)y;
  p.sz = (const uint16_t*)sz;
  p.lut = (const uint16_t*)lut;
  p.exps = exps;
  p.lut_stride = lut_stride;
  p.is_mx4 = mx4 ? 1 : 0;
  p.rows_x = (int)rows_x;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.k_tiles = (int)div_up(k, 16);
  p.ik = ik;
  p.outer_k = (int)div_up(p.k_tiles, ik);
  p.glog2 = glog2_of(group);
  return dt == TG_BF16 ? launch_simple<TG_BF16, W4, true>(p, st) : launch_simple<TG_FP16, W4, true>(p, st);
}

int launch_gemm_w16_rm(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k, int ik,
                       tg_weight_side side, tg_dtype dt, cudaStream_t st) {
  GParams p{};
  p.w = reinterpret_cast<const uint32_t*>(w);
  p.x = (const uint16_t*)x;
  p.y = (uint16_t*)y;
  p.rows_x = (int)rows_x;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.k_tiles = (int)div_up(k, 16);
  p.ik = side == TG_WEIGHT_A ? 1 : ik;
  p.outer_k = (int)div_up(p.k_tiles, p.ik);
  p.glog2 = 5;
  return dispatch<W16>(p, side, dt, st);
}

}  // namespace tg
