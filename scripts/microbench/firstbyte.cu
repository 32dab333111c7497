// Microbenchmark 3: time to first data at kernel start, LDG.128 vs bulk TMA, cold (never-touched-recently) memory.
// 148 CTAs x 512 threads; each CTA fetches its first 32 KiB either with 4 LDG.128 per thread or with 8 x 4 KiB bulk
// copies issued by one thread; stamps %globaltimer at entry and when the data is complete.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(512, 1) k(const uint8_t* src, size_t cta_stride, int mode, unsigned long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  unsigned long long t0 = gtime();
  const uint8_t* p = src + (size_t)blockIdx.x * cta_stride;
  if (mode == 0) {
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(p + (size_t)(i * 512 + threadIdx.x) * 16));
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) a ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    if (a == 0x1234567) *sink = a;
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x * 4] = t0; out[blockIdx.x * 4 + 1] = gtime(); }
  } else if (mode >= 2) {
    // mode 2: 8 x 4 KiB, one copy per warp (parallel issue); 3: one 32 KiB copy; 4: 1 x 4 KiB; 5: 2 x 4 KiB;
    // mode 6: 8 x 4 KiB by one thread, but only the first 4 KiB is waited for (2 barriers)
    const uint32_t bar = smem_u32(sm), data = bar + 1024;
    const uint32_t want = mode == 4 ? 4096u : mode == 5 ? 8192u : 32768u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar + 8));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (mode == 6) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(4096) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar + 8), "r"(28672) : "memory");
      } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(want) : "memory");
      }
    }
    if (mode == 2) __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (mode == 2) {
      if (lane == 0 && warp < 8)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + warp * 4096), "l"(p + warp * 4096), "r"(4096), "r"(bar) : "memory");
    } else if (threadIdx.x == 0) {
      if (mode == 3) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data), "l"(p), "r"(32768), "r"(bar) : "memory");
      } else if (mode == 6) {
        for (int c = 0; c < 8; ++c)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + c * 4096), "l"(p + c * 4096), "r"(4096), "r"(bar + (c ? 8 : 0)) : "memory");
      } else {
        for (uint32_t c = 0; c < want / 4096u; ++c)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + c * 4096), "l"(p + c * 4096), "r"(4096), "r"(bar) : "memory");
      }
    }
    if (threadIdx.x == 0) {
      out[blockIdx.x * 4 + 2] = gtime();
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar) : "memory");
      out[blockIdx.x * 4] = t0; out[blockIdx.x * 4 + 1] = gtime();
      if (mode == 6) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar + 8) : "memory");
        out[blockIdx.x * 4 + 3] = gtime();
      }
    }
  } else {
    const uint32_t bar = smem_u32(sm), data = bar + 1024;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32768) : "memory");
      for (int c = 0; c < 8; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + c * 4096), "l"(p + c * 4096), "r"(4096), "r"(bar) : "memory");
      out[blockIdx.x * 4 + 2] = gtime();
      asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar) : "memory");
      out[blockIdx.x * 4] = t0; out[blockIdx.x * 4 + 1] = gtime();
    }
  }
}
int main() {
  const int max_ctas = 148;
  const size_t total = (size_t)3 << 30;
  uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  unsigned long long* out; CK(cudaMalloc(&out, max_ctas * 4 * 8)); uint32_t* sink; CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  std::vector<unsigned long long> h(max_ctas * 4);
  size_t base = 0;
  static const char* names[] = {"LDG.128 32K", "bulk 8x4K 1thr", "bulk 8x4K 8warps", "bulk 1x32K", "bulk 1x4K only", "bulk 2x4K only", "bulk 8x4K, first 4K"};
  for (int rep = 0; rep < 3; ++rep)
   for (int ctas : {148, 1})
    for (int mode = 0; mode < 7; ++mode)
      for (size_t stride : {(size_t)65536}) {
        base = (base + ((size_t)700 << 20)) % (total - ((size_t)700 << 20));  // far from anything touched recently
        CK(cudaMemset(out, 0, ctas * 4 * 8));
        k<<<ctas, 512, 40 * 1024>>>(src + base, stride, mode, out, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, ctas * 4 * 8, cudaMemcpyDeviceToHost));
        std::vector<double> d, iss, all;
        for (int c = 0; c < ctas; ++c) { d.push_back((h[c * 4 + 1] - h[c * 4]) / 1e3); iss.push_back(h[c * 4 + 2] ? (h[c * 4 + 2] - h[c * 4]) / 1e3 : 0); all.push_back(h[c * 4 + 3] ? (h[c * 4 + 3] - h[c * 4]) / 1e3 : 0); }
        std::sort(d.begin(), d.end()); std::sort(iss.begin(), iss.end()); std::sort(all.begin(), all.end());
        printf("ctas=%3d %-20s: complete after median %.2f us (p10 %.2f, p90 %.2f); issue done after %.2f us; rest after %.2f\n", ctas, names[mode], d[ctas / 2], d[ctas / 10], d[ctas * 9 / 10], iss[ctas / 2], all[ctas / 2]);
      }
  return 0;
}
