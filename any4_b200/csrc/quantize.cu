// any4 quantizer front-end on the GPU: group scaling + per-row 1-D k-means + direct emission of the
// packed tensor-core layout, one kernel, one CTA per weight row.
//
// Replaces, for the configuration the any4 Linear uses (4 bit, per-row LUT, asymmetric groups with zero point, "int"
// initialisation), the reference's CPU pipeline
//   group_q                      quantize.py:106-149   scale = clamp(max - min, 1e-6) / 15, zero = min + 8 * scale,
//                                                      v = (w - min) / scale in [0, 15]
//   cluster_matrix / cluster_row quantize.py:433-521   one weighted 1-D k-means (16 clusters) per row over v
//   kmeans.run_kmeans            kmeans.py:200-262     Lloyd: nearest centroid (ties: lower index), weighted mean, empty
//                                                      clusters keep their centroid; stops when the labels repeat, when
//                                                      the centroids move less than var(v) * tol, or after max_iter
//   init "int"                   kmeans.py:41-46       torch.linspace(min(v), max(v), 16)
//   lut = any4 - 8               quantize.py:893       in the weight dtype
//   convert_..._Bint4_layout     TinyGemmConvertB.cu:252-308 (a 64 MiB int32 code matrix per 4096^2 layer in between)
// which takes sklearn / joblib minutes per model.  Numerics: group statistics, v and the stored scale / zero are the
// same IEEE fp32 operations (bit-identical); the k-means sums are fp32 in a fixed (deterministic) order that differs
// from numpy's, so a value that sits within an ulp of a cluster boundary may take the neighbouring code.
//
// 1-D structure used: the centroids stay sorted (the mean of an interval lies inside it; an empty cluster keeps a
// centroid between its neighbours'), so "nearest centroid, ties to the lower index" is "number of boundaries
// m_j = (c_j + c_j+1) / 2 with v > m_j": 15 compares instead of 16 distances.
#include "common.cuh"

namespace tg {
namespace {

constexpr int kQThreads = 256;
constexpr int kQWarps = kQThreads / 32;
constexpr int K = 16;

template <tg_dtype DT>
__device__ __forceinline__ float q_load(const void* w, int64_t i) {
  if constexpr (DT == TG_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[i]);
  else return __half2float(reinterpret_cast<const __half*>(w)[i]);
}
template <tg_dtype DT>
__device__ __forceinline__ uint16_t q_round(float f) {
  if constexpr (DT == TG_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(f));
  else return __half_as_ushort(__float2half_rn(f));
}
template <tg_dtype DT>
__device__ __forceinline__ float q_widen(uint16_t v) {
  if constexpr (DT == TG_BF16) return __bfloat162float(__ushort_as_bfloat16(v));
  else return __half2float(__ushort_as_half(v));
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct QParams {
  const void* w;          // [n][k] weight dtype
  const float* sw;        // [k] sample weights or null
  int32_t* codes;         // [n][k] or null
  int32_t* packed;        // B int4 layout [n/8][k/(ik*16)][32][ik/2] or null
  uint16_t* sz;           // [k/g][n][2]
  uint16_t* any4;         // [n][16] centroids in [0, 15] code space, weight dtype
  uint16_t* lut;          // [n][16] = any4 - 8 in the weight dtype
  int* iters;             // [n] Lloyd iterations run, or null
  int n, k, glog2, ik, max_iter;
  float tol;
};

// dynamic shared memory: v[k] float | labels[k] uint8
template <tg_dtype DT>
__global__ void __launch_bounds__(kQThreads) any4_quantize_rows_kernel(const QParams p) {
  extern __shared__ __align__(16) uint8_t q_smem[];
  float* v = reinterpret_cast<float*>(q_smem);
  uint8_t* lab = q_smem + (size_t)p.k * 4;
  __shared__ float red[kQWarps][3 * K];
  __shared__ float cen[K], bnd[K];
  __shared__ float stat[4];
  __shared__ int flag[2];

  const int row = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = 1 << p.glog2;
  const int n_groups = p.k >> p.glog2;

  // ---- group_q: one group per warp at a time ----
  for (int grp = warp; grp < n_groups; grp += kQWarps) {
    float mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < g; i += 32) {
      const float x = q_load<DT>(p.w, (int64_t)row * p.k + grp * g + i);
      v[grp * g + i] = x;
      mn = fminf(mn, x), mx = fmaxf(mx, x);
    }
    mn = warp_min(mn), mx = warp_max(mx);
    const float scale = __fdiv_rn(fmaxf(__fsub_rn(mx, mn), 1e-6f), 15.0f);
    const float zero = __fadd_rn(mn, __fmul_rn(scale, 8.0f));
    if (lane == 0) {
      const int64_t o = ((int64_t)grp * p.n + row) * 2;
      p.sz[o] = q_round<DT>(scale);
      p.sz[o + 1] = q_round<DT>(zero);
    }
    __syncwarp();
    for (int i = lane; i < g; i += 32) v[grp * g + i] = __fdiv_rn(__fsub_rn(v[grp * g + i], mn), scale);
  }
  __syncthreads();

  // ---- row statistics: min, max (init), variance (tolerance) ----
  {
    float mn = INFINITY, mx = -INFINITY, s = 0.f;
    for (int i = tid; i < p.k; i += kQThreads) {
      const float x = v[i];
      mn = fminf(mn, x), mx = fmaxf(mx, x), s += x;
    }
    mn = warp_min(mn), mx = warp_max(mx), s = warp_sum(s);
    if (lane == 0) red[warp][0] = mn, red[warp][1] = mx, red[warp][2] = s;
    __syncthreads();
    if (tid == 0) {
      float a = red[0][0], b = red[0][1], c = red[0][2];
      for (int w2 = 1; w2 < kQWarps; ++w2) a = fminf(a, red[w2][0]), b = fmaxf(b, red[w2][1]), c += red[w2][2];
      stat[0] = a, stat[1] = b, stat[2] = c / (float)p.k;
    }
    __syncthreads();
    const float mean = stat[2];
    float q = 0.f;
    for (int i = tid; i < p.k; i += kQThreads) {
      const float d = v[i] - mean;
      q += d * d;
    }
    q = warp_sum(q);
    __syncthreads();
    if (lane == 0) red[warp][0] = q;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w2 = 0; w2 < kQWarps; ++w2) a += red[w2][0];
      stat[3] = a / (float)p.k * p.tol;  // kmeans._tolerance: mean variance * tol
      // torch.linspace(mn, mx, 16): start + step * i for the first half, end - step * (15 - i) for the second
      const float step = (stat[1] - stat[0]) / 15.0f;
      for (int j = 0; j < K; ++j) cen[j] = j < K / 2 ? stat[0] + step * (float)j : stat[1] - step * (float)(K - 1 - j);
    }
  }
  for (int i = tid; i < p.k; i += kQThreads) lab[i] = 0;  // run_kmeans starts from labels = 0
  __syncthreads();

  // ---- Lloyd ----
  int it = 0;
  for (; it < p.max_iter; ++it) {
    if (tid < K - 1) bnd[tid] = 0.5f * (cen[tid] + cen[tid + 1]);
    if (tid == 0) flag[0] = 0;
    __syncthreads();
    float b[K - 1];
#pragma unroll
    for (int j = 0; j < K - 1; ++j) b[j] = bnd[j];
    float sx[K], sw_[K], cn[K];
#pragma unroll
    for (int j = 0; j < K; ++j) sx[j] = 0.f, sw_[j] = 0.f, cn[j] = 0.f;
    float swx[K];
#pragma unroll
    for (int j = 0; j < K; ++j) swx[j] = 0.f;
    int changed = 0;
    for (int i = tid; i < p.k; i += kQThreads) {
      const float x = v[i];
      int l = 0;
#pragma unroll
      for (int j = 0; j < K - 1; ++j) l += x > b[j] ? 1 : 0;
      changed |= (l != (int)lab[i]);
      lab[i] = (uint8_t)l;
      const float wgt = p.sw ? p.sw[i] : 1.0f;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const bool mine = l == j;
        sx[j] += mine ? x : 0.f;
        swx[j] += mine ? wgt * x : 0.f;
        sw_[j] += mine ? wgt : 0.f;
        cn[j] += mine ? 1.f : 0.f;
      }
    }
    if (changed) flag[0] = 1;  // benign race: every writer stores 1
    // deterministic block reduction of the 4 x 16 partial sums (two rounds through red[][3 * K])
    float tot_sx = 0.f, tot_swx = 0.f, tot_sw = 0.f, tot_cn = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) sx[j] = warp_sum(sx[j]), swx[j] = warp_sum(swx[j]), sw_[j] = warp_sum(sw_[j]), cn[j] = warp_sum(cn[j]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < K; ++j) red[warp][j] = sx[j], red[warp][K + j] = swx[j], red[warp][2 * K + j] = sw_[j];
    }
    __syncthreads();
    if (tid < K) {
      for (int w2 = 0; w2 < kQWarps; ++w2) tot_sx += red[w2][tid], tot_swx += red[w2][K + tid], tot_sw += red[w2][2 * K + tid];
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < K; ++j) red[warp][j] = cn[j];
    }
    __syncthreads();
    if (!flag[0]) break;  // labels repeat: converged (kmeans.py:228-232), centroids stay
    if (tid < K) {
      for (int w2 = 0; w2 < kQWarps; ++w2) tot_cn += red[w2][tid];
      const float old = cen[tid];
      float c = old;
      if (tot_cn > 0.f) c = tot_sw == 0.f ? tot_sx / tot_cn : tot_swx / tot_sw;  // np.average with / without weights
      cen[tid] = c;
      const float d = c - old;
      bnd[tid] = d * d;  // (bnd is rebuilt at the top of the next iteration)
    }
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int j = 0; j < K; ++j) s += bnd[j];
      flag[1] = (it > 0 && sqrtf(s) < stat[3]) ? 1 : 0;  // kmeans.py:254: centroids moved less than tol
    }
    __syncthreads();
    if (flag[1]) {
      ++it;
      break;
    }
  }
  __syncthreads();
  if (p.iters && tid == 0) p.iters[row] = it;

  // ---- outputs: any4 / lut in the weight dtype; codes follow the labels of the last assignment (run_kmeans returns
  // `labels`, which belong to the centroids BEFORE the final update - kept as is) ----
  if (tid < K) {
    const uint16_t a = q_round<DT>(cen[tid]);
    p.any4[(int64_t)row * K + tid] = a;
    p.lut[(int64_t)row * K + tid] = q_round<DT>(q_widen<DT>(a) - 8.0f);
  }
  if (p.codes) {
    for (int i = tid; i < p.k; i += kQThreads) p.codes[(int64_t)row * p.k + i] = (int32_t)lab[i];
  }
  if (p.packed) {
    // B int4 layout (TinyGemmConvertB.cu:252-308): n-tile nt = row / 8, g8 = row % 8; for outer k index ko and lane
    // t = g8 * 4 + q the lane holds ik / 2 words, word j = k-tiles (2j, 2j + 1) of the ik-group:
    //   v0..v3 = codes at k0, k0+1, k0+8, k0+9 of tile 2j (k0 = tile * 16 + 2q), v4..v7 the same of tile 2j+1
    //   word = v7<<28 | v5<<24 | v3<<20 | v1<<16 | v6<<12 | v4<<8 | v2<<4 | v0          (ConvertB.cu:280-303)
    const int nt = row >> 3, g8 = row & 7;
    const int k_tiles = p.k >> 4;
    const int outer = (k_tiles + p.ik - 1) / p.ik;
    const int wpl = p.ik >> 1;  // words per lane
    const int n_words = outer * 4 * wpl;
    for (int wi = tid; wi < n_words; wi += kQThreads) {
      const int j = wi % wpl, q = (wi / wpl) & 3, ko = wi / (wpl * 4);
      uint32_t word = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int tile = ko * p.ik + 2 * j + h;
        uint32_t c[4] = {0u, 0u, 0u, 0u};
        if (tile < k_tiles) {
          const int k0 = tile * 16 + 2 * q;
          c[0] = lab[k0], c[1] = lab[k0 + 1], c[2] = lab[k0 + 8], c[3] = lab[k0 + 9];
        }
        // nibble positions: (v0, v2, v4, v6) at bits 0, 4, 8, 12 and (v1, v3, v5, v7) at 16, 20, 24, 28
        word |= (c[0] << (8 * h)) | (c[2] << (8 * h + 4)) | (c[1] << (16 + 8 * h)) | (c[3] << (20 + 8 * h));
      }
      p.packed[(((int64_t)nt * outer + ko) * 32 + (g8 * 4 + q)) * wpl + j] = (int32_t)word;
    }
  }
}

}  // namespace

int launch_quantize_any4_rows(const void* w, const float* sample_weight, int64_t n, int64_t k, int group, int inner_k_tiles,
                              int max_iter, float tol, int32_t* codes, int32_t* packed, void* sz, void* any4, void* lut,
                              int* iters, tg_dtype dt, cudaStream_t st) {
  QParams p{};
  p.w = w, p.sw = sample_weight, p.codes = codes, p.packed = packed;
  p.sz = static_cast<uint16_t*>(sz), p.any4 = static_cast<uint16_t*>(any4), p.lut = static_cast<uint16_t*>(lut);
  p.iters = iters;
  p.n = (int)n, p.k = (int)k, p.ik = inner_k_tiles, p.max_iter = max_iter, p.tol = tol;
  p.glog2 = group == 32 ? 5 : group == 64 ? 6 : group == 128 ? 7 : 8;
  const size_t smem = (size_t)k * 5;
  auto kern = dt == TG_BF16 ? any4_quantize_rows_kernel<TG_BF16> : any4_quantize_rows_kernel<TG_FP16>;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("tg_quantize_any4_rows: k = %lld needs %zu bytes of shared memory: %s", (long long)k, smem,
                cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_UNSUPPORTED;
    }
  }
  kern<<<(unsigned)n, kQThreads, smem, st>>>(p);
  TG_CHECK_LAUNCH("any4_quantize_rows_kernel");
  return TG_OK;
}

}  // namespace tg
