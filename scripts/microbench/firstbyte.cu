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
  const int ctas = 148;
  const size_t total = (size_t)3 << 30;
  uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  unsigned long long* out; CK(cudaMalloc(&out, ctas * 4 * 8)); uint32_t* sink; CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  std::vector<unsigned long long> h(ctas * 4);
  size_t base = 0;
  for (int rep = 0; rep < 3; ++rep)
    for (int mode = 0; mode < 2; ++mode)
      for (size_t stride : {(size_t)65536, (size_t)(4 << 20)}) {
        base = (base + ((size_t)700 << 20)) % (total - ((size_t)700 << 20));  // far from anything touched recently
        CK(cudaMemset(out, 0, ctas * 4 * 8));
        k<<<ctas, 512, 40 * 1024>>>(src + base, stride, mode, out, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, ctas * 4 * 8, cudaMemcpyDeviceToHost));
        std::vector<double> d, iss;
        for (int c = 0; c < ctas; ++c) { d.push_back((h[c * 4 + 1] - h[c * 4]) / 1e3); iss.push_back(h[c * 4 + 2] ? (h[c * 4 + 2] - h[c * 4]) / 1e3 : 0); }
        std::sort(d.begin(), d.end()); std::sort(iss.begin(), iss.end());
        printf("%s cta_stride=%7zu: first 32 KiB complete after median %.2f us (p10 %.2f, p90 %.2f); issue done after %.2f us\n", mode ? "bulk-TMA" : "LDG.128 ", stride, d[ctas / 2], d[ctas / 10], d[ctas * 9 / 10], iss[ctas / 2]);
      }
  return 0;
}
