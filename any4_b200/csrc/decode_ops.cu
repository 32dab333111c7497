// Decode-step plumbing around the quantized GEMVs: the element-wise / attention work of one
// single-token Llama-style decoder layer as three small kernels, so that a layer is
//   add+rmsnorm -> QKV GEMV -> rope+attention -> O GEMV -> add+rmsnorm -> gate/up GEMV -> silu*mul -> down GEMV
// (8 launches instead of ~30 stock element-wise launches).  The reference leaves this to the HF model code
// (benchmark.py:145-146 times the whole model); nothing here changes the GEMV's numerics.
//
// All three kernels are PDL citizens: `griddepcontrol.launch_dependents` at entry lets the NEXT kernel of the stream
// (a weight-streaming GEMV with static weights) become resident and start its TMA stream while this kernel is still
// waiting for / working on its inputs; `griddepcontrol.wait` orders them behind the previous kernel.  They are sized
// to co-reside with a GEMV CTA (<= 256 threads, <= 32 registers would be ideal; <= 4 KiB static shared memory).
#include "common.cuh"

namespace tg {
namespace w4 {
extern bool g_pdl;
}
namespace {

template <typename T>
struct Cvt;
template <>
struct Cvt<__nv_bfloat16> {
  __device__ static float f(uint16_t v) { return __uint_as_float((uint32_t)v << 16); }
  __device__ static uint16_t r(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
};
template <>
struct Cvt<__half> {
  __device__ static float f(uint16_t v) { return __half2float(__ushort_as_half(v)); }
  __device__ static uint16_t r(float v) { return __half_as_ushort(__float2half_rn(v)); }
};

__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------
// h <- h + delta (rounded to T, as the framework's residual add does); out <- rmsnorm(h) * weight
// (fp32 statistics, one rounding).  One CTA; n <= 256 * 8 * kMaxVec.
// ---------------------------------------------------------------------------------------
constexpr int kNormThreads = 256;
constexpr int kNormMaxPerThread = 32;  // n <= 8192

template <typename T>
__global__ void __launch_bounds__(kNormThreads, 1)
add_rmsnorm_kernel(uint16_t* __restrict__ h, const uint16_t* __restrict__ delta, const uint16_t* __restrict__ weight,
                   uint16_t* __restrict__ out, int n, float eps) {
  __shared__ float red[kNormThreads / 32];
  // the norm weights are static: fetched before the dependency resolves (one L2 round trip off the critical path)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint4 wreg[kNormMaxPerThread / 8];
#pragma unroll
  for (int i = 0; i < kNormMaxPerThread / 8; ++i) {
    const int vi = (int)threadIdx.x + i * kNormThreads;
    wreg[i] = vi < (n >> 3) ? reinterpret_cast<const uint4*>(weight)[vi] : make_uint4(0, 0, 0, 0);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t hp[kNormMaxPerThread / 2];  // the (updated) residual stream, packed pairs
  float ss = 0.f;
  const int nvec = n >> 3;  // 16-byte vectors
#pragma unroll
  for (int i = 0; i < kNormMaxPerThread / 8; ++i) {
    const int vi = (int)threadIdx.x + i * kNormThreads;
    if (vi < nvec) {
      uint4 hv = reinterpret_cast<const uint4*>(h)[vi];
      uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
      if (delta != nullptr) {
        const uint4 dv = reinterpret_cast<const uint4*>(delta)[vi];
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint16_t lo = Cvt<T>::r(Cvt<T>::f((uint16_t)hw[j]) + Cvt<T>::f((uint16_t)dw[j]));
          const uint16_t hi = Cvt<T>::r(Cvt<T>::f((uint16_t)(hw[j] >> 16)) + Cvt<T>::f((uint16_t)(dw[j] >> 16)));
          hw[j] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        reinterpret_cast<uint4*>(h)[vi] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = Cvt<T>::f((uint16_t)hw[j]), b = Cvt<T>::f((uint16_t)(hw[j] >> 16));
        hp[i * 4 + j] = hw[j];
        ss += a * a + b * b;
      }
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) tot += red[w];
  const float rs = rsqrtf(tot / (float)n + eps);
#pragma unroll
  for (int i = 0; i < kNormMaxPerThread / 8; ++i) {
    const int vi = (int)threadIdx.x + i * kNormThreads;
    if (vi < nvec) {
      const uint4 wv = wreg[i];
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint16_t lo = Cvt<T>::r(Cvt<T>::f((uint16_t)hp[i * 4 + j]) * rs * Cvt<T>::f((uint16_t)ww[j]));
        const uint16_t hi = Cvt<T>::r(Cvt<T>::f((uint16_t)(hp[i * 4 + j] >> 16)) * rs * Cvt<T>::f((uint16_t)(ww[j] >> 16)));
        o[j] = (uint32_t)lo | ((uint32_t)hi << 16);
      }
      reinterpret_cast<uint4*>(out)[vi] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// out[i] = silu(gate[i]) * up[i], gate = gate_up[0..n), up = gate_up[n..2n): silu rounded to T, then the product
// rounded to T (the two roundings of the framework's separate silu and mul kernels).
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) silu_mul_kernel(const uint16_t* __restrict__ gate_up, uint16_t* __restrict__ out,
                                                       int n) {
  pdl_prologue();
  const int vi = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (vi >= (n >> 3)) return;
  const uint4 gv = reinterpret_cast<const uint4*>(gate_up)[vi];
  const uint4 uv = reinterpret_cast<const uint4*>(gate_up + n)[vi];
  const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, uw[4] = {uv.x, uv.y, uv.z, uv.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint16_t r2[2];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const float g = Cvt<T>::f((uint16_t)(gw[j] >> (16 * hlf)));
      const float u = Cvt<T>::f((uint16_t)(uw[j] >> (16 * hlf)));
      const float s = Cvt<T>::f(Cvt<T>::r(g / (1.f + __expf(-g))));
      r2[hlf] = Cvt<T>::r(s * u);
    }
    o[j] = (uint32_t)r2[0] | ((uint32_t)r2[1] << 16);
  }
  reinterpret_cast<uint4*>(out)[vi] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------
// Rotary embedding of q and the new k (half-rotation convention), KV-cache append at `pos`, and single-query
// attention over positions 0..pos for one token.  head_dim = 128.  One CTA (128 threads) per query head; the CTAs of a
// GQA group all compute the group's new k (cheap) and use it from shared memory, the group's first CTA stores k and
// v to the cache, so no CTA reads what another one writes.
//   qkv [n_heads*128 | n_kv*128 | n_kv*128]; cos/sin [128]; caches [n_kv][cache_len][128]; out [n_heads*128]
// q and the new k are rounded to T after the rotation (as stored / as the framework's rope does), scores and the
// softmax are fp32, probabilities stay fp32 for the value sum, one rounding at the end.
// ---------------------------------------------------------------------------------------
constexpr int kHeadDim = 128;
constexpr int kAttnMaxLen = 512;  // positions held in static shared memory (keeps the CTA small enough to co-reside)

template <typename T>
__global__ void __launch_bounds__(kHeadDim, 1)
rope_attn_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ cosv, const uint16_t* __restrict__ sinv,
                 uint16_t* __restrict__ kc, uint16_t* __restrict__ vc, uint16_t* __restrict__ out, int n_heads,
                 int n_kv, int pos, int cache_len, float scale, int n_pre) {
  __shared__ float q_s[kHeadDim], k_s[kHeadDim], sc[kAttnMaxLen + 1], red[4];
  extern __shared__ __align__(16) uint4 kv_s[];  // [n_pre][16] K rows, then [n_pre][16] V rows of this kv head
  const int h = blockIdx.x, kvh = h / (n_heads / n_kv);
  const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
  // Programmatic dependent launch: the cached positions 0..pos-1 were written by EARLIER tokens, not by the kernel this
  // one depends on (the q|k|v GEMV) - so the first n_pre cache rows are pulled into shared memory (cp.async, no
  // registers) while that GEMV is still running; only q, the new k and v wait for it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  {
    const uint4* kg = reinterpret_cast<const uint4*>(kc + (size_t)kvh * cache_len * kHeadDim);
    const uint4* vg = reinterpret_cast<const uint4*>(vc + (size_t)kvh * cache_len * kHeadDim);
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(kv_s);
    for (int i = d; i < n_pre * 16; i += kHeadDim) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + (uint32_t)i * 16u), "l"(kg + i) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + (uint32_t)(n_pre * 16 + i) * 16u), "l"(vg + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float rope_c = Cvt<T>::f(cosv[d]), rope_s = Cvt<T>::f(sinv[d]);  // static tables: before the wait as well
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint16_t* qh = qkv + (size_t)h * kHeadDim;
  const uint16_t* kh = qkv + (size_t)n_heads * kHeadDim + (size_t)kvh * kHeadDim;
  const uint16_t* vh = qkv + (size_t)(n_heads + n_kv) * kHeadDim + (size_t)kvh * kHeadDim;
  {
    // rotate_half: y[d] = x[d]*cos[d] - x[d+64]*sin[d] (d < 64), y[d] = x[d]*cos[d] + x[d-64]*sin[d] (d >= 64); the
    // framework evaluates x*cos, rot*sin and the sum as three rounded T operations - mirrored here
    const float c = rope_c, s = rope_s;
    const int pd = d < 64 ? d + 64 : d - 64;
    const float sgn = d < 64 ? -1.f : 1.f;
    auto rot = [&](const uint16_t* x) {
      const float a = Cvt<T>::f(Cvt<T>::r(Cvt<T>::f(x[d]) * c));
      const float b = Cvt<T>::f(Cvt<T>::r(sgn * Cvt<T>::f(x[pd]) * s));
      return Cvt<T>::r(a + b);
    };
    const uint16_t qr = rot(qh), kr = rot(kh);
    q_s[d] = Cvt<T>::f(qr);
    k_s[d] = Cvt<T>::f(kr);
    if (h % (n_heads / n_kv) == 0) {
      kc[((size_t)kvh * cache_len + pos) * kHeadDim + d] = kr;
      vc[((size_t)kvh * cache_len + pos) * kHeadDim + d] = vh[d];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // scores: thread = (position slot tid/16, 16-byte segment tid%16): 8 cached positions per pass, the loads of
  // kUnroll passes are in flight together (the cache rows come from L2 / HBM: latency, not bandwidth, is the cost)
  constexpr int kUnroll = 8;
  const int seg = d & 15, slot = d >> 4;
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = q_s[seg * 8 + i];
  const uint16_t* kbase = kc + (size_t)kvh * cache_len * kHeadDim;
  for (int p0 = 0; p0 < pos; p0 += 8 * kUnroll) {
    uint4 kv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int p = p0 + u * 8 + slot;
      kv[u] = p < n_pre ? kv_s[p * 16 + seg]
                        : (p < pos ? reinterpret_cast<const uint4*>(kbase + (size_t)p * kHeadDim)[seg] : make_uint4(0, 0, 0, 0));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t w[4] = {kv[u].x, kv[u].y, kv[u].z, kv[u].w};
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        s += qv[2 * j] * Cvt<T>::f((uint16_t)w[j]) + qv[2 * j + 1] * Cvt<T>::f((uint16_t)(w[j] >> 16));
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int p = p0 + u * 8 + slot;
      if (seg == 0 && p < pos) sc[p] = s * scale;
    }
  }
  if (warp == 0) {
    const float4 q4 = reinterpret_cast<const float4*>(q_s)[lane];
    const float4 k4 = reinterpret_cast<const float4*>(k_s)[lane];
    float s = warp_sum(q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w);
    if (lane == 0) sc[pos] = s * scale;
  }
  __syncthreads();
  // softmax over 0..pos
  float mx = -INFINITY;
  for (int p = d; p <= pos; p += kHeadDim) mx = fmaxf(mx, sc[p]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int p = d; p <= pos; p += kHeadDim) {
    const float e = __expf(sc[p] - mx);
    sc[p] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  // values: the same (slot, segment) decomposition - a thread accumulates its 8 dims over the positions of its
  // slot, then the 8 slots are summed (pairs by shuffle, the rest through shared memory)
  const uint16_t* vbase = vc + (size_t)kvh * cache_len * kHeadDim;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int p0 = 0; p0 < pos; p0 += 8 * kUnroll) {
    uint4 vv[kUnroll];
    float pr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int p = p0 + u * 8 + slot;
      vv[u] = p < n_pre ? kv_s[(n_pre + p) * 16 + seg]
                        : (p < pos ? reinterpret_cast<const uint4*>(vbase + (size_t)p * kHeadDim)[seg] : make_uint4(0, 0, 0, 0));
      pr[u] = p < pos ? sc[p] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint32_t w[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] = fmaf(pr[u], Cvt<T>::f((uint16_t)w[j]), acc[2 * j]);
        acc[2 * j + 1] = fmaf(pr[u], Cvt<T>::f((uint16_t)(w[j] >> 16)), acc[2 * j + 1]);
      }
    }
  }
  const float p_new = sc[pos];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);  // slots 2w and 2w+1
  __syncthreads();                                                                // everyone is done with sc[]
  float* part = sc;                                                               // [4 warps][128 dims]
  if (lane < 16) {
#pragma unroll
    for (int i = 0; i < 8; ++i) part[warp * kHeadDim + seg * 8 + i] = acc[i];
  }
  __syncthreads();
  const float o = part[d] + part[kHeadDim + d] + part[2 * kHeadDim + d] + part[3 * kHeadDim + d] +
                  p_new * Cvt<T>::f(vh[d]);
  out[(size_t)h * kHeadDim + d] = Cvt<T>::r(o * inv);
}

// ---------------------------------------------------------------------------------------
// host activations -> device staging buffer (tg_gemm_w4_rm_hostio): 16-byte pieces straight out of pinned host memory
// (unified addressing).  One CTA: the 8 KB of a decode step are latency, not bandwidth; the dependent GEMV (static
// weights) streams its weights and fills its first tensor-memory slots while these loads cross PCIe.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) stage_host_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
  // x_host is an INPUT written by the host: its first 16 KiB are fetched over PCIe (the slow part, ~2 us) before the
  // previous kernel of the stream has finished; only the staging buffer, which that kernel may still be reading, is
  // ordered behind it.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int kPre = 4;
  uint4 v[kPre];
#pragma unroll
  for (int j = 0; j < kPre; ++j) {
    const int i = (int)threadIdx.x + 256 * j;
    if (i < n16)
      asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(src + i) : "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int j = 0; j < kPre; ++j) {
    const int i = (int)threadIdx.x + 256 * j;
    if (i < n16) dst[i] = v[j];
  }
  for (int i = (int)threadIdx.x + 256 * kPre; i < n16; i += 256) {
    uint4 w;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(src + i) : "memory");
    dst[i] = w;
  }
}

template <typename... KArgs, typename... Args>
int launch_pdl_smem(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const char* what, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = w4::g_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return TG_ERR_CUDA;
  }
  count_launch();
  return TG_OK;
}
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, const char* what, Args... args) {
  return launch_pdl_smem(kern, grid, block, 0, st, what, args...);
}

}  // namespace
}  // namespace tg

using namespace tg;

extern "C" int tg_decode_add_rmsnorm(void* h, const void* delta, const void* weight, void* out, int64_t n, float eps,
                                     tg_dtype dtype, void* stream) {
  const char* fn = "tg_decode_add_rmsnorm";
  TG_REQUIRE(h != nullptr && weight != nullptr && out != nullptr, "%s: null pointer", fn);
  TG_REQUIRE(n > 0 && n % 8 == 0 && n <= kNormThreads * kNormMaxPerThread, "%s: n = %lld must be a multiple of 8, <= %d",
             fn, (long long)n, kNormThreads * kNormMaxPerThread);
  TG_REQUIRE(dtype == TG_BF16 || dtype == TG_FP16, "%s: bad dtype", fn);
  auto st = (cudaStream_t)stream;
  if (dtype == TG_BF16)
    return launch_pdl(add_rmsnorm_kernel<__nv_bfloat16>, dim3(1), dim3(kNormThreads), st, fn, (uint16_t*)h,
                      (const uint16_t*)delta, (const uint16_t*)weight, (uint16_t*)out, (int)n, eps);
  return launch_pdl(add_rmsnorm_kernel<__half>, dim3(1), dim3(kNormThreads), st, fn, (uint16_t*)h, (const uint16_t*)delta,
                    (const uint16_t*)weight, (uint16_t*)out, (int)n, eps);
}

extern "C" int tg_decode_silu_mul(const void* gate_up, void* out, int64_t n, tg_dtype dtype, void* stream) {
  const char* fn = "tg_decode_silu_mul";
  TG_REQUIRE(gate_up != nullptr && out != nullptr, "%s: null pointer", fn);
  TG_REQUIRE(n > 0 && n % 8 == 0 && n < (1ll << 30), "%s: n = %lld must be a positive multiple of 8", fn, (long long)n);
  TG_REQUIRE(dtype == TG_BF16 || dtype == TG_FP16, "%s: bad dtype", fn);
  auto st = (cudaStream_t)stream;
  const dim3 grid((unsigned)div_up(n / 8, 256));
  if (dtype == TG_BF16)
    return launch_pdl(silu_mul_kernel<__nv_bfloat16>, grid, dim3(256), st, fn, (const uint16_t*)gate_up, (uint16_t*)out, (int)n);
  return launch_pdl(silu_mul_kernel<__half>, grid, dim3(256), st, fn, (const uint16_t*)gate_up, (uint16_t*)out, (int)n);
}

extern "C" int tg_decode_rope_attention(const void* qkv, const void* cos, const void* sin, void* k_cache, void* v_cache,
                                        void* out, int n_heads, int n_kv_heads, int head_dim, int pos, int cache_len,
                                        float scale, tg_dtype dtype, void* stream) {
  const char* fn = "tg_decode_rope_attention";
  TG_REQUIRE(qkv && cos && sin && k_cache && v_cache && out, "%s: null pointer", fn);
  TG_REQUIRE(head_dim == kHeadDim, "%s: head_dim must be %d (got %d)", fn, kHeadDim, head_dim);
  TG_REQUIRE(n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0, "%s: n_heads %d / n_kv_heads %d", fn, n_heads,
             n_kv_heads);
  TG_REQUIRE(pos >= 0 && pos < cache_len && pos <= kAttnMaxLen, "%s: pos %d must be < cache_len %d and <= %d", fn, pos,
             cache_len, kAttnMaxLen);
  TG_REQUIRE(dtype == TG_BF16 || dtype == TG_FP16, "%s: bad dtype", fn);
  auto st = (cudaStream_t)stream;
  // cache rows prefetched before the dependency resolves: up to 192 positions = 96 KB of shared memory (a CTA of this
  // size still co-resides with one GEMV CTA)
  constexpr int kPreMax = 192;
  const int n_pre = pos < kPreMax ? pos : kPreMax;
  static thread_local bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[current_device_slot()];
  if (!attr_set) {
    if (cudaFuncSetAttribute(rope_attn_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreMax * 512) != cudaSuccess ||
        cudaFuncSetAttribute(rope_attn_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreMax * 512) != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", fn, cudaGetErrorString(cudaGetLastError()));
      return TG_ERR_CUDA;
    }
    attr_set = true;
  }
  const size_t smem = (size_t)n_pre * 512;
  if (dtype == TG_BF16)
    return launch_pdl_smem(rope_attn_kernel<__nv_bfloat16>, dim3(n_heads), dim3(kHeadDim), smem, st, fn, (const uint16_t*)qkv,
                           (const uint16_t*)cos, (const uint16_t*)sin, (uint16_t*)k_cache, (uint16_t*)v_cache,
                           (uint16_t*)out, n_heads, n_kv_heads, pos, cache_len, scale, n_pre);
  return launch_pdl_smem(rope_attn_kernel<__half>, dim3(n_heads), dim3(kHeadDim), smem, st, fn, (const uint16_t*)qkv,
                         (const uint16_t*)cos, (const uint16_t*)sin, (uint16_t*)k_cache, (uint16_t*)v_cache, (uint16_t*)out,
                         n_heads, n_kv_heads, pos, cache_len, scale, n_pre);
}

extern "C" int tg_gemm_w4_rm_hostio(void* y_host, const void* x_host, void* x_staging, const int32_t* w,
                                    const void* scales_zeros, const void* lut, const uint8_t* exponents, int64_t rows_x,
                                    int64_t w_rows, int64_t k, int group, int inner_k_tiles, tg_w4_format format,
                                    tg_weight_side side, tg_dtype dtype, void* stream) {
  const char* fn = "tg_gemm_w4_rm_hostio";
  TG_REQUIRE(y_host && x_host && x_staging, "%s: null pointer", fn);
  TG_REQUIRE(rows_x >= 1 && k > 0 && (rows_x * k) % 8 == 0 && rows_x * k * 2 <= (1ll << 20),
             "%s: the activations (rows_x * k) must be a multiple of 8 elements and at most 1 MiB", fn);
  TG_REQUIRE(((reinterpret_cast<uintptr_t>(x_host) | reinterpret_cast<uintptr_t>(x_staging)) & 15u) == 0,
             "%s: x_host and x_staging must be 16-byte aligned", fn);
  int rc = launch_pdl(stage_host_kernel, dim3(1), dim3(256), (cudaStream_t)stream, fn, (const uint4*)x_host,
                      (uint4*)x_staging, (int)(rows_x * k / 8));
  if (rc != TG_OK) return rc;
  return tg_gemm_w4_rm(y_host, x_staging, w, scales_zeros, lut, exponents, rows_x, w_rows, k, group, inner_k_tiles, format,
                       side, dtype, stream);
}
