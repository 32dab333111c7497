// 4-bit weights (B layout, weight on the right) for 5..16+ activation rows: fragment-order tensor-core kernel.
//
// gemv_w4_b.cu (lane per weight row) is built for decode: its block-structured mma operand carries at most 4
// activation rows, so m = 16 streams and dequantises the weights four times.  Here the packed words are used as what
// they are - mma.m16n8k16 B fragments (TinyGemmConvertB.cu:252-308) - so ONE pass covers 8 (or 16) activation rows:
//   * a CTA (8 warps splitting k) walks over row tiles (8 weight rows) with a persistent grid; the activations of a pass
//     are staged once in shared memory (padded row stride: conflict-free operand loads);
//   * per row tile the CTA builds the same byte-pair table as the decode kernel, pair[b] = (LUT[b & 15], LUT[b >> 4]),
//     for the tile's 8 rows (8 KiB; built once if the LUT is global); a packed byte (k0 | k0+8) is dequantised by one
//     LDS.32 + one single-rounded HFMA2 (scale, zero), two PRMTs re-pair the halves into the (k0, k0+1) / (k0+8, k0+9)
//     fragment registers;
//   * U units (super-tiles of IK k-tiles) and their group words are fetched per lane before any is decoded.
// Replaces the same reference path as gemv_w4_b.cu (TinyGemm_int4.cu:294-548, MatrixLayoutB.cuh:686-1101) for m > 4.
// Numerics: identical dequantised weights, exact products, fp32 accumulation in the tensor core, one RN.  A non-finite
// weight only reaches its own output column (ordinary mma operand layout), as in the reference.
#include "common.cuh"
#include "w4_common.cuh"

namespace tg {
namespace {

constexpr int kFWarps = 8;
constexpr int kFThreads = kFWarps * 32;
constexpr int kFMaxXSmem = 96 * 1024;

struct FParams {
  const uint32_t* w;     // packed [n/8][k/(ik*16)][32][ik/2]
  const uint16_t* x;     // [rows_x][k]
  uint16_t* y;           // [rows_x][w_rows]
  const uint32_t* sz;    // [k/g][w_rows] (scale, zero) pairs, null for mx4
  const uint8_t* exps;   // [w_rows][k/g] e8m0, mx4 only
  const uint16_t* lut;   // [16] or [w_rows][16]
  int lut_stride;        // 0 or 16
  int rows_x, w_rows, k, glog2, n_units;
};

template <int N>
__device__ __forceinline__ void load_words_f(const uint32_t* __restrict__ src, uint32_t* dst) {
  if constexpr (N == 4) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
    dst[0] = v.x, dst[1] = v.y, dst[2] = v.z, dst[3] = v.w;
  } else if constexpr (N == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(src));
    dst[0] = v.x, dst[1] = v.y;
  } else {
    dst[0] = __ldg(src);
  }
}

template <tg_dtype DT, int IK, bool HI>
__global__ void __launch_bounds__(kFThreads, 2) gemm_w4_frag_kernel(const FParams p, int rows_per_pass, int kpad) {
  constexpr int NW = IK / 2;                 // packed words per lane per unit (one word = two k-tiles)
  constexpr int U = NW == 4 ? 2 : 4;         // units in flight per lane (register budget: 2 CTAs per SM)
  extern __shared__ __align__(16) uint8_t smem_f[];
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem_f);               // [256 entries][8 rows]
  uint16_t* xs = reinterpret_cast<uint16_t*>(smem_f + 8192);         // [rows_per_pass][xstride]
  __shared__ float red[kFWarps][4][32];
  const int xstride = kpad + 8;  // 4 * odd words: 8 activation rows x 4 k-pairs of one operand load -> 32 distinct banks

  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  const int g = t >> 2, q = t & 3;
  const int n_units = p.n_units;
  const int n_tiles = p.w_rows >> 3;
  const int n_groups = p.k >> p.glog2;
  const bool is_mx4 = p.sz == nullptr;
  const bool vec_ok = ((p.k & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  const uint32_t tab_lane = w4::smem_u32(tab) + (uint32_t)g * 4u;

  uint32_t raw[U][NW];
  uint32_t szv[U][NW];
  auto load_batch = [&](int rt, int u0) {
    const uint32_t* wrow = p.w + ((int64_t)rt * n_units * 32 + t) * NW;
    const int row = rt * 8 + g;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int u = u0 + j * kFWarps;
      if (u < n_units) {
        load_words_f<NW>(wrow + (int64_t)u * 32 * NW, raw[j]);
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) {
          const int grp = min(((u * IK + 2 * wi) * 16) >> p.glog2, n_groups - 1);  // (k padding: any group, x is 0 there)
          if (is_mx4) szv[j][wi] = w4::e8m0_to_dt<DT>((uint32_t)__ldg(p.exps + (int64_t)row * n_groups + grp)) | 0x80000000u;
          else szv[j][wi] = __ldg(p.sz + (int64_t)grp * p.w_rows + row);
        }
      }
    }
  };
  // pair table of the 8 rows of tile rt: thread = (row r, high nibble hi, low-nibble half)
  auto build_table = [&](int rt) {
    const int r = threadIdx.x & 7, half = (threadIdx.x >> 3) & 1, hi = threadIdx.x >> 4;
    const uint16_t* lrow = p.lut + (int64_t)(rt * 8 + r) * p.lut_stride;
    const uint4 lo8 = __ldg(reinterpret_cast<const uint4*>(lrow) + half);  // LUT[half*8 .. half*8+7]
    const uint32_t hv = (uint32_t)__ldg(lrow + hi) << 16;
    const uint32_t lw[4] = {lo8.x, lo8.y, lo8.z, lo8.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t lv = (lw[i >> 1] >> (16 * (i & 1))) & 0xffffu;
      tab[(hi * 16 + half * 8 + i) * 8 + r] = lv | hv;
    }
  };

  for (int a0 = 0; a0 < p.rows_x; a0 += rows_per_pass) {
    const int na = min(rows_per_pass, p.rows_x - a0);
    int rt = blockIdx.x;
    bool preloaded = false;
    if (rt < n_tiles) {
      load_batch(rt, warp);
      preloaded = true;
      if (a0 == 0 || p.lut_stride != 0) build_table(rt);
    }
    // stage the activations of this pass (16-byte pieces, zero beyond k)
    for (int i = threadIdx.x; i < na * (kpad >> 3); i += kFThreads) {
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
    const bool has_lo = g < na, has_hi = HI && (g + 8 < na);
    const uint16_t* x_lo = xs + (size_t)(has_lo ? g : 0) * xstride + 2 * q;
    const uint16_t* x_hi = xs + (size_t)(has_hi ? g + 8 : 0) * xstride + 2 * q;

    for (; rt < n_tiles; rt += gridDim.x) {
      float acc[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;

      for (int u0 = warp; u0 < n_units; u0 += kFWarps * U) {
        if (!preloaded) load_batch(rt, u0);
        preloaded = false;
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int u = u0 + j * kFWarps;
          if (u >= n_units) break;
#pragma unroll
          for (int wi = 0; wi < NW; ++wi) {
            const uint32_t w = raw[j][wi];
            const uint32_t s2 = __byte_perm(szv[j][wi], 0, 0x1010), z2 = __byte_perm(szv[j][wi], 0, 0x3232);
            // byte i -> pair (k0 | k0+8) [bytes 0, 1] or (k0+1 | k0+9) [bytes 2, 3] of tile 2wi [0, 2] / 2wi+1 [1, 3]
            uint32_t e0 = w4::lds32(tab_lane + ((w << 5) & 0x1fe0u));
            uint32_t e1 = w4::lds32(tab_lane + ((w >> 3) & 0x1fe0u));
            uint32_t e2 = w4::lds32(tab_lane + ((w >> 11) & 0x1fe0u));
            uint32_t e3 = w4::lds32(tab_lane + ((w >> 19) & 0x1fe0u));
            e0 = w4::fma2<DT>(e0, s2, z2);
            e1 = w4::fma2<DT>(e1, s2, z2);
            e2 = w4::fma2<DT>(e2, s2, z2);
            e3 = w4::fma2<DT>(e3, s2, z2);
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              const uint32_t lo = tt ? e1 : e0, hi2 = tt ? e3 : e2;
              const uint32_t b0 = __byte_perm(lo, hi2, 0x5410);  // (k0, k0+1)
              const uint32_t b1 = __byte_perm(lo, hi2, 0x7632);  // (k0+8, k0+9)
              const int kt = (u * IK + 2 * wi + tt) * 16;
              const uint32_t xl0 = has_lo ? *reinterpret_cast<const uint32_t*>(x_lo + kt) : 0u;
              const uint32_t xl1 = has_lo ? *reinterpret_cast<const uint32_t*>(x_lo + kt + 8) : 0u;
              uint32_t xh0 = 0u, xh1 = 0u;
              if constexpr (HI) {
                xh0 = has_hi ? *reinterpret_cast<const uint32_t*>(x_hi + kt) : 0u;
                xh1 = has_hi ? *reinterpret_cast<const uint32_t*>(x_hi + kt + 8) : 0u;
              }
              w4::mma16816<DT>(acc[tt], xl0, xh0, xl1, xh1, b0, b1);
            }
          }
        }
      }
      const int rt_next = rt + (int)gridDim.x;
      if (rt_next < n_tiles) {  // the next tile's first batch is in flight during the reduction
        load_batch(rt_next, warp);
        preloaded = true;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) red[warp][i][t] = acc[0][i] + acc[1][i];
      __syncthreads();  // partial sums visible; nobody reads this tile's table any more
      if (rt_next < n_tiles && p.lut_stride != 0) build_table(rt_next);
      if (threadIdx.x < 128) {
        const int ci = threadIdx.x >> 5, gl = (threadIdx.x & 31) >> 2, ql = threadIdx.x & 3;
        float sum = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kFWarps; ++w2) sum += red[w2][ci][threadIdx.x & 31];
        // C fragment: c0,c1 = (act g, weight rows 2q, 2q+1), c2,c3 = (act g+8, ...)
        const int act = gl + 8 * (ci >> 1), wrow = 2 * ql + (ci & 1);
        if (act < na) p.y[(int64_t)(a0 + act) * p.w_rows + rt * 8 + wrow] = w4::f32_to_dt<DT>(sum);
      }
      __syncthreads();  // next table complete, red[] free
    }
    __syncthreads();
  }
}

template <tg_dtype DT, int IK, bool HI>
int launch_frag(const FParams& p, int rows_per_pass, int kpad, cudaStream_t st) {
  auto kern = gemm_w4_frag_kernel<DT, IK, HI>;
  static thread_local int ctas_per_sm_dev[kMaxDevices] = {}, n_sm = 0;
  int& ctas_per_sm = ctas_per_sm_dev[current_device_slot()];
  const size_t smem = 8192 + (size_t)rows_per_pass * (kpad + 8) * 2;
  if (ctas_per_sm == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + kFMaxXSmem) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_w4_frag_kernel) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFThreads, 8192 + 72 * 1024) != cudaSuccess || occ <= 0) occ = 2;
    ctas_per_sm = occ;
  }
  const int tiles = p.w_rows >> 3;
  const int slots = ctas_per_sm * n_sm;
  kern<<<tiles < slots ? tiles : slots, kFThreads, smem, st>>>(p, rows_per_pass, kpad);
  TG_CHECK_LAUNCH("gemm_w4_frag_kernel");
  return TG_OK;
}

template <tg_dtype DT, int IK>
int launch_frag_ik(const FParams& p, cudaStream_t st) {
  const int kpad = p.n_units * IK * 16;
  int rows_per_pass = kFMaxXSmem / ((kpad + 8) * 2);
  if (rows_per_pass < 8) return -1;  // long k: fewer than 8 rows per pass would not beat the lane-per-row kernel
  if (rows_per_pass > 16) rows_per_pass = 16;
  if (rows_per_pass > p.rows_x) rows_per_pass = p.rows_x;
  if (rows_per_pass > 8 && rows_per_pass < 16 && rows_per_pass < p.rows_x) rows_per_pass = 8;  // whole 8-row passes
  if (rows_per_pass > 8) return launch_frag<DT, IK, true>(p, rows_per_pass, kpad, st);
  return launch_frag<DT, IK, false>(p, rows_per_pass, kpad, st);
}

}  // namespace

// returns TG_OK / TG_ERR_*, or -1 if the shape is not handled here (caller falls back to the lane-per-row kernel)
int launch_gemm_w4_frag_B(void* y, const void* x, const int32_t* w, const void* sz, const void* lut, const uint8_t* exps,
                          int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik, tg_w4_format fmt, tg_dtype dt,
                          const uint16_t* const_lut, cudaStream_t st) {
  FParams p{};
  p.w = reinterpret_cast<const uint32_t*>(w);
  p.x = static_cast<const uint16_t*>(x);
  p.y = static_cast<uint16_t*>(y);
  p.sz = (fmt == TG_W4_MX4) ? nullptr : static_cast<const uint32_t*>(sz);
  p.exps = (fmt == TG_W4_MX4) ? exps : nullptr;
  if (fmt == TG_W4_ANY4_GLOBAL || fmt == TG_W4_ANY4_ROWWISE) {
    p.lut = static_cast<const uint16_t*>(lut);
    p.lut_stride = fmt == TG_W4_ANY4_ROWWISE ? 16 : 0;
  } else {
    p.lut = const_lut;
    p.lut_stride = 0;
  }
  p.rows_x = (int)rows_x;
  p.w_rows = (int)w_rows;
  p.k = (int)k;
  p.glog2 = group == 32 ? 5 : group == 64 ? 6 : group == 128 ? 7 : 8;
  p.n_units = (int)div_up(div_up(k, 16), ik);
  int rc = -1;
  if (dt == TG_BF16) {
    rc = ik == 2 ? launch_frag_ik<TG_BF16, 2>(p, st) : ik == 4 ? launch_frag_ik<TG_BF16, 4>(p, st)
                                                                 : ik == 8 ? launch_frag_ik<TG_BF16, 8>(p, st) : -1;
  } else {
    rc = ik == 2 ? launch_frag_ik<TG_FP16, 2>(p, st) : ik == 4 ? launch_frag_ik<TG_FP16, 4>(p, st)
                                                                 : ik == 8 ? launch_frag_ik<TG_FP16, 8>(p, st) : -1;
  }
  return rc;
}

}  // namespace tg
