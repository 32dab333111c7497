// Microbenchmark 2: start-up ramp of a bulk-TMA stream.  148 CTAs each stream `total_kb` KiB in 32 KiB stages
// (8 x 4 KiB copies, either one contiguous run or 4 runs 32 KiB apart like the packed-weight tiles) and record
// %globaltimer when stages 0,1,2,3,... land.  Prints the median over CTAs relative to the first CTA's start.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
constexpr int kStage = 32768;
// mode: 0 contiguous, 1 four tiles strided by `tile_stride`; hint: 0 none, 1 evict_first; ldg: also time a plain LDG
__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, size_t cta_stride, int n_stage, int N, int mode, size_t tile_stride, int hint,
                                           unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t bars = smem_u32(sm), data = bars + 1024;
  unsigned long long* o = out + (size_t)blockIdx.x * 32;
  if (threadIdx.x == 0) {
    o[0] = gtime();
    for (int i = 0; i < N; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + i * 8)); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 256 + i * 8)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* p = src + (size_t)blockIdx.x * cta_stride;
  if (threadIdx.x == 0) {
    uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int j = 0; j < n_stage; ++j) {
      const int s = j % N;
      if (j >= N) mbar_wait(bars + 256 + s * 8, ((j / N) - 1) & 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + s * 8), "r"(kStage) : "memory");
      for (int c = 0; c < 8; ++c) {
        const uint8_t* g = mode == 0 ? p + (size_t)j * kStage + c * 4096 : p + (size_t)(c >> 1) * tile_stride + (size_t)j * 8192 + (c & 1) * 4096;
        if (hint) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(data + s * kStage + c * 4096), "l"(g), "r"(4096), "r"(bars + s * 8), "l"(pol) : "memory");
        else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + s * kStage + c * 4096), "l"(g), "r"(4096), "r"(bars + s * 8) : "memory");
      }
      if (j == 0) o[1] = gtime();
    }
    o[2] = gtime();
  } else if (threadIdx.x == 32) {
    // a plain load for comparison
    unsigned long long t0 = gtime();
    volatile const uint32_t* q = (const uint32_t*)(p + cta_stride / 2);
    uint32_t v = *q;
    o[3] = gtime() - t0 + (v == 0x12345 ? 1 : 0);
    for (int j = 0; j < n_stage; ++j) {
      const int s = j % N;
      mbar_wait(bars + s * 8, (j / N) & 1);
      if (j < 24) o[4 + j] = gtime();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 256 + s * 8) : "memory");
    }
    o[30] = gtime();
  }
}
int main() {
  const int ctas = 148;
  const size_t total = (size_t)1 << 30;
  uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  unsigned long long* out; CK(cudaMalloc(&out, ctas * 32 * 8));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  std::vector<unsigned long long> h(ctas * 32);
  for (int N : {3, 5})
  for (int mode = 0; mode < 2; ++mode)
    for (int hint = 0; hint < 2; ++hint)
      for (int n_stage : {2, 8}) {
        size_t cta_stride = (size_t)n_stage * kStage;          // like one GEMV: CTAs' regions are adjacent
        size_t tile_stride = (size_t)n_stage * 8192;           // = 4k bytes for k = n_stage*2048
        // rotate the base so L2 is cold
        static size_t base = 0; base = (base + ((size_t)64 << 20)) % (total - ((size_t)64 << 20));
        CK(cudaMemset(out, 0, ctas * 32 * 8));
        k<<<ctas, 64, 1024 + N * kStage>>>(src + base, cta_stride, n_stage, N, mode, tile_stride, hint, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out, ctas * 32 * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull; for (int c = 0; c < ctas; ++c) t0 = std::min(t0, h[c * 32]);
        auto med = [&](int slot, bool rel) { std::vector<double> v; for (int c = 0; c < ctas; ++c) v.push_back(rel ? (h[c * 32 + slot] - t0) / 1e3 : h[c * 32 + slot] / 1e3); std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
        printf("stages=%d ring=%d mode=%s hint=%d: start %.2f issued0 %.2f issued_all %.2f ldg_lat %.2f | landed:", n_stage, N, mode ? "4tiles" : "contig", hint, med(0, true), med(1, true), med(2, true), med(3, false));
        for (int j = 0; j < n_stage; ++j) printf(" %.2f", med(4 + j, true));
        printf(" | end %.2f us  (%.0f GB/s)\n", med(30, true), (double)ctas * n_stage * kStage / (med(30, true) * 1e3));
      }
  return 0;
}
