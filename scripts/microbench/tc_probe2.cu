// Micro-timings of the pieces of the tcgen05 GEMV's dequant loop (one CTA per SM, 8 warps like the kernel):
// cycles per "slot" (32 byte-pair lookups + 32 HFMA2 + 4 x tcgen05.st.x8 per thread) for growing subsets of the
// per-slot work, plus the cost of fence.proxy.async, an mbarrier try_wait on a completed phase and a 4 KiB bulk copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe2 tc_probe2.cu && ./tc_probe2
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ uint32_t lds_table(uint32_t idx) {
  uint32_t v;
  asm("ld.shared.b32 %0, [%1+1024];" : "=r"(v) : "r"(idx));
  return v;
}
__device__ __forceinline__ uint32_t fma2(uint32_t v, uint32_t s, uint32_t z) {
  uint32_t r;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(s), "r"(z));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

constexpr int kIters = 64;

template <int MODE>
__device__ __forceinline__ unsigned long long run_slots(uint32_t tmem_lane, uint32_t wbase, uint32_t lane4, uint32_t bar,
                                                        uint32_t& sink) {
  unsigned long long t0 = clock64();
  uint32_t s2 = 0x3f803f80u, z2 = 0x00000000u;
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    const uint4 a = lds128(wbase + (it & 7) * 512), b = lds128(wbase + (it & 7) * 512 + 16);
    const uint32_t W[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int t4 = 0; t4 < 4; ++t4) {
      uint32_t r[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w = W[q * 2 + (t4 >> 1)];
        r[2 * q] = lds_table(prmt(w, lane4, 0x7604u | ((t4 & 1) << 4)));
        r[2 * q + 1] = lds_table(prmt(w, lane4, 0x7604u | (((t4 & 1) + 2) << 4)));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fma2(r[i], s2, z2);
      if (MODE >= 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_lane + 64u + t4 * 8u),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sink ^= r[i];
      }
    }
    if (MODE >= 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (MODE >= 3) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    }
  }
  return clock64() - t0;
}

__global__ void __launch_bounds__(576, 1) probe3(uint32_t nwarps, unsigned long long* __restrict__ out, uint32_t* __restrict__ sink_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t holder = sbase + 65536 + 49536;
  const uint32_t bar = holder + 16;
  for (int i = threadIdx.x; i < (65536 + 49536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 49536);
  uint32_t sink = 0;
  if (warp < (int)nwarps) {
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
    const uint32_t wbase = sbase + 65536 + (warp & 7) * 512 + lane * 32;
    unsigned long long c3 = run_slots<3>(tl, wbase, lane * 4, bar, sink);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = c3 / kIters;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  if (sink == 0x12345u) sink_out[threadIdx.x] = sink;
}

__global__ void __launch_bounds__(352, 2) probe2(const uint8_t* __restrict__ gsrc, unsigned long long* __restrict__ out,
                                                 uint32_t* __restrict__ sink_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t holder = sbase + 65536 + 49536;
  const uint32_t bar = holder + 16, bar2 = holder + 24;
  for (int i = threadIdx.x; i < (65536 + 49536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 49536);
  uint32_t sink = 0;
  if (warp < 8) {
    const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t wbase = sbase + 65536 + warp * 512 + lane * 32;
    const uint32_t lane4 = lane * 4;
    unsigned long long c0 = run_slots<0>(tl, wbase, lane4, bar, sink);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    unsigned long long c1 = run_slots<1>(tl, wbase, lane4, bar, sink);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    unsigned long long c2 = run_slots<2>(tl, wbase, lane4, bar, sink);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    unsigned long long c3 = run_slots<3>(tl, wbase, lane4, bar, sink);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) {
      out[0] = c0 / kIters, out[1] = c1 / kIters, out[2] = c2 / kIters, out[3] = c3 / kIters;
    }
    // fence.proxy.async after a shared store
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + 128 + lane4 + warp * 256), "r"(i) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[4] = (t1 - t0) / 16;
  } else if (warp == 9 && lane == 0) {
    // try_wait on a completed phase
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar2) : "memory");
    unsigned long long t0 = clock64();
    uint32_t ok = 0;
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
      uint32_t o;
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(o) : "r"(bar2) : "memory");
      ok += o;
    }
    unsigned long long t1 = clock64();
    if (blockIdx.x == 0) out[5] = (t1 - t0) / 16, out[7] = ok;
  } else if (warp == 10 && lane == 0) {
    // 4 KiB bulk copies issued back to back by one thread (completion not awaited inside the timed region)
    const uint32_t bar3 = holder + 32;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar3));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar3), "r"(8u * 4096u) : "memory");
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 8; ++i)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       sbase + 65536 + 16384 + i * 4096),
                   "l"(gsrc + ((size_t)blockIdx.x * 8 + i) * 4096), "r"(4096u), "r"(bar3)
                   : "memory");
    unsigned long long t1 = clock64();
    asm volatile("{ .reg .pred p; W3: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D3; bra W3; D3: }" ::"r"(bar3) : "memory");
    unsigned long long t2 = clock64();
    if (blockIdx.x == 0) out[6] = (t1 - t0) / 8, out[8] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  if (sink == 0x12345u) sink_out[threadIdx.x] = sink;
}

int main() {
  const int smem = 65536 + 49536 + 64;
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d_out;
  uint32_t* d_sink;
  uint8_t* d_src;
  cudaMalloc(&d_out, 16 * 8);
  cudaMalloc(&d_sink, 4096);
  cudaMalloc(&d_src, (size_t)296 * 8 * 4096);
  cudaMemset(d_out, 0, 128);
  for (int grid : {1, 148, 296}) {
    for (int rep = 0; rep < 2; ++rep) {
      probe2<<<grid, 352, smem>>>(d_src, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("CUDA error %s\n", cudaGetErrorString(e));
        return 1;
      }
    }
    unsigned long long h[16];
    cudaMemcpy(h, d_out, 128, cudaMemcpyDeviceToHost);
    printf("grid %3d: cycles per slot (8 warps): lookups+fma %llu | +STTM %llu | +wait::st %llu | +fence+arrive %llu || fence.proxy.async %llu | "
           "try_wait(done) %llu | bulk 4K issue %llu (8 copies landed after %llu)\n",
           grid, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[8]);
  }
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int nw : {1, 2, 4, 6, 8, 12, 16}) {
    probe3<<<148, 576, smem>>>(nw, d_out, d_sink);
    cudaDeviceSynchronize();
    probe3<<<148, 576, smem>>>(nw, d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[1];
    cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("one CTA/SM, %2d dequant warps: %llu cycles per 32-lookup slot per warp -> %.1f nibbles/clk/SM\n", nw, h[0], nw * 2048.0 / h[0]);
  }
  return 0;
}
