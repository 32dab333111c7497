// Fragment-order weight-only GEMM for every packed format that is NOT the tuned B-layout
// 4-bit kernel (gemv_w4_b.cu): int4/any4/mx4 in the "A" layout (weight on the left), int8 and
// 16-bit weights in either layout.
//
// Replaces tinygemm_m16n8k16_chunk_kernel instantiated with ALayout_TC_int4 / *_int8 / *_TC
// (reference: TinyGemmImpl.cuh:23-345, MatrixLayoutA.cuh:375-1062, MatrixLayoutB.cuh:461-684,
// :1103-1328, Dequantization.cuh:265-328).
//
// One CTA owns one weight row tile (16 rows in the A layout, 8 in the B layout); its warps
// split the k-tiles.  A lane reads exactly the words the packed layout assigns to its lane
// id, so every warp load is one fully coalesced 128..512-byte run, decodes them to the
// activation dtype with the reference's single-rounded FMA, and accumulates exact products in
// fp32 (FFMA on up-converted operands: a bf16/fp16 product is exact in fp32, so this is the
// same arithmetic as the tensor core's fp32 accumulate).  The four lanes that share a weight
// row are combined with warp shuffles, the warps through shared memory, one RN at the end.
#include "common.cuh"

namespace tg {
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

template <tg_dtype DT, Kind KIND, bool ALAYOUT>
int launch(const GParams& p, cudaStream_t st) {
  const int tiles = p.w_rows / (ALAYOUT ? 16 : 8);
  gemm_frag_kernel<DT, KIND, ALAYOUT><<<tiles, kThreads, 0, st>>>(p);
  TG_CHECK_LAUNCH("gemm_frag_kernel");
  return TG_OK;
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
