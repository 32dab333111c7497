// Probe of the tcgen05 pieces gemv_w4_tc.cu relies on (run on a B200 through gpurun):
//   * A operand written to TMEM with tcgen05.st.32x32b.x8 (lane = row, one 32-bit column = K pair (2c, 2c+1))
//   * B operand read from shared memory through a K-major, no-swizzle descriptor with arbitrary LBO / SBO,
//     including "overlapping" placements where unused rows alias other data
//   * M = 128, N = 16/32/64, K = 16, f16 and bf16 inputs, fp32 accumulators, tcgen05.commit -> mbarrier, tcgen05.ld
// The host evaluates the layout hypothesis
//     B(n, kk) at  start + (n % 8) * 16 + (n / 8) * SBO + (kk / 8) * LBO + (kk % 8) * 2       [H1]
// (and the swapped one, H2) and reports which one the hardware follows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu && ./tc_probe
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

struct Cfg {
  uint32_t start_off, lbo, sbo, kstep_stride;  // bytes
  int n;                                        // MMA N
  int ksteps;
  int bf16;
};

constexpr int kImageBytes = 32768;
constexpr int kMaxSteps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const uint32_t* __restrict__ a_words, const uint8_t* __restrict__ b_image,
                                                Cfg cfg, float* __restrict__ d_out, unsigned long long* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_holder;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < kImageBytes / 16; i += 128)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(b_image)[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of B -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_holder;
  // A: columns [0, 8*ksteps), D: columns [64, 64+n)
  const uint32_t my_lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  unsigned long long t0 = clock64();
  for (int s = 0; s < cfg.ksteps; ++s) {
    const uint32_t* src = a_words + ((size_t)threadIdx.x * kMaxSteps + s) * 8;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = src[i];
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(my_lane_addr + 8u * s),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  unsigned long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t fmt = cfg.bf16 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(cfg.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sbase = smem_u32(smem) + cfg.start_off;
    for (int s = 0; s < cfg.ksteps; ++s) {
      const uint32_t addr = sbase + (uint32_t)s * cfg.kstep_stride;
      const uint64_t desc = (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((cfg.lbo >> 4) & 0x3fffu) << 16) |
                            ((uint64_t)((cfg.sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
      const uint32_t acc = s > 0 ? 1u : 0u;
      asm volatile(
          "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(tbase + 64u),
          "r"(tbase + 8u * s), "l"(desc), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  unsigned long long t2 = clock64();
  // wait for the MMAs
  asm volatile(
      "{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar))
      : "memory");
  unsigned long long t3 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < cfg.n; c0 += 16) {
    uint32_t v[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(my_lane_addr + 64u + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) d_out[(size_t)threadIdx.x * 64 + c0 + i] = __uint_as_float(v[i]);
  }
  unsigned long long t4 = clock64();
  if (threadIdx.x == 0) {
    clk[0] = t1 - t0;
    clk[1] = t2 - t1;
    clk[2] = t3 - t2;
    clk[3] = t4 - t3;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(128u));
  (void)lane;
}

static uint16_t enc(int v, int bf16) {
  if (bf16) {
    __nv_bfloat16 b = __float2bfloat16((float)v);
    return *reinterpret_cast<uint16_t*>(&b);
  }
  __half h = __float2half((float)v);
  return *reinterpret_cast<uint16_t*>(&h);
}

int main() {
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kImageBytes);
  std::vector<Cfg> cfgs = {
      // start, lbo, sbo, kstep_stride, n, ksteps, bf16
      {0, 128, 256, 512, 16, 4, 0},     // dense 16 rows: core matrices 128 B, K halves adjacent, row groups 256 B apart
      {0, 256, 128, 512, 16, 4, 0},     // the other way round
      {0, 64, 128, 128, 16, 4, 0},      // m = 1 overlap trick: 4 useful rows, K halves 64 B apart
      {0, 256, 512, 1024, 16, 4, 1},    // half-line placement (pitch 256), bf16
      {128, 512, 256, 1024, 16, 8, 1},  // odd half-lines, two row groups, 8 steps
      {128, 1024, 256, 2048, 32, 4, 1},
      {128, 2048, 256, 4096, 64, 4, 1},
      {64, 256, 0, 512, 16, 4, 1},      // SBO = 0: rows 8..15 alias rows 0..7
  };
  uint32_t* d_a;
  uint8_t* d_b;
  float* d_d;
  unsigned long long* d_clk;
  cudaMalloc(&d_a, 128 * kMaxSteps * 8 * 4);
  cudaMalloc(&d_b, kImageBytes);
  cudaMalloc(&d_d, 128 * 64 * 4);
  cudaMalloc(&d_clk, 64);
  int bad_total = 0;
  for (size_t ci = 0; ci < cfgs.size(); ++ci) {
    const Cfg c = cfgs[ci];
    srand(1234 + (int)ci);
    std::vector<int> a(128 * kMaxSteps * 16), b(kImageBytes / 2);
    for (auto& v : a) v = rand() % 9 - 4;
    for (auto& v : b) v = rand() % 9 - 4;
    std::vector<uint32_t> aw(128 * kMaxSteps * 8);
    for (int l = 0; l < 128; ++l)
      for (int s = 0; s < kMaxSteps; ++s)
        for (int col = 0; col < 8; ++col)
          aw[(l * kMaxSteps + s) * 8 + col] = (uint32_t)enc(a[(l * kMaxSteps + s) * 16 + 2 * col], c.bf16) |
                                              ((uint32_t)enc(a[(l * kMaxSteps + s) * 16 + 2 * col + 1], c.bf16) << 16);
    std::vector<uint16_t> bi(kImageBytes / 2);
    for (size_t i = 0; i < bi.size(); ++i) bi[i] = enc(b[i], c.bf16);
    cudaMemcpy(d_a, aw.data(), aw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, bi.data(), kImageBytes, cudaMemcpyHostToDevice);
    cudaMemset(d_d, 0xff, 128 * 64 * 4);
    probe<<<1, 128, kImageBytes>>>(d_a, d_b, c, d_d, d_clk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("cfg %zu: CUDA error %s\n", ci, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> d(128 * 64);
    unsigned long long clk[4];
    cudaMemcpy(d.data(), d_d, d.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(clk, d_clk, 32, cudaMemcpyDeviceToHost);
    int bad[2] = {0, 0};
    for (int h = 0; h < 2; ++h) {
      const uint32_t lbo = h == 0 ? c.lbo : c.sbo, sbo = h == 0 ? c.sbo : c.lbo;
      for (int l = 0; l < 128; ++l)
        for (int n = 0; n < c.n; ++n) {
          long exp = 0;
          for (int s = 0; s < c.ksteps; ++s)
            for (int kk = 0; kk < 16; ++kk) {
              const uint32_t off = c.start_off + s * c.kstep_stride + (n % 8) * 16 + (n / 8) * sbo + (kk / 8) * lbo + (kk % 8) * 2;
              exp += (long)a[(l * kMaxSteps + s) * 16 + kk] * b[off / 2];
            }
          if ((float)exp != d[l * 64 + n]) ++bad[h];
        }
    }
    printf("cfg %zu (start %u lbo %u sbo %u step %u n %d steps %d %s): H1 mismatches %d, H2(swapped) %d | clk st %llu issue %llu mma-wait %llu ld %llu\n",
           ci, c.start_off, c.lbo, c.sbo, c.kstep_stride, c.n, c.ksteps, c.bf16 ? "bf16" : "f16", bad[0], bad[1], clk[0],
           clk[1], clk[2], clk[3]);
    if (bad[0]) {
      printf("  D[0][0..7] =");
      for (int n = 0; n < 8; ++n) printf(" %g", d[n]);
      printf("\n");
    }
    bad_total += bad[0] != 0;
  }
  printf(bad_total ? "PROBE: %d configs disagree with H1\n" : "PROBE: all configs follow H1\n", bad_total);
  return 0;
}
