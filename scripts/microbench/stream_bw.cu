// Microbenchmark: how fast can ONE CTA per SM stream HBM -> shared memory / registers on B200, by data path?
//   mode 0: cp.async.bulk 1-D (UBLKCP) ring, copy size S, N stages
//   mode 1: LDG.128 to registers (xor-reduced), W warps, unroll U
//   mode 2: cp.async 16 B (LDGSTS) ring with commit groups
//   mode 3: TMA 2-D tensor map (cp.async.bulk.tensor.2d), box = [rows][256 B], N stages
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bw stream_bw.cu ; run: ./stream_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---------------- mode 0: 1-D bulk ----------------
__global__ void __launch_bounds__(64, 1) k_bulk(const uint8_t* src, size_t per_cta, int S, int N, int chunks_per_stage, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t bars = smem_u32(sm);           // full[N], empty[N]
  const uint32_t data = bars + 1024;
  const int stage_bytes = S * chunks_per_stage;
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; ++i) { mbar_init(bars + i * 8, 1); mbar_init(bars + 256 + i * 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* p = src + (size_t)blockIdx.x * per_cta;
  const int iters = (int)(per_cta / stage_bytes);
  if (threadIdx.x == 0) {        // producer
    for (int j = 0; j < iters; ++j) {
      const int s = j % N;
      if (j >= N) mbar_wait(bars + 256 + s * 8, ((j / N) - 1) & 1);
      mbar_expect(bars + s * 8, stage_bytes);
      for (int c = 0; c < chunks_per_stage; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(data + s * stage_bytes + c * S), "l"(p + (size_t)j * stage_bytes + (size_t)c * S), "r"(S), "r"(bars + s * 8) : "memory");
    }
  } else if (threadIdx.x == 32) { // consumer
    unsigned long long acc = 0;
    for (int j = 0; j < iters; ++j) {
      const int s = j % N;
      mbar_wait(bars + s * 8, (j / N) & 1);
      acc += sm[1024 + s * stage_bytes];
      mbar_arrive(bars + 256 + s * 8);
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

// ---------------- mode 1: LDG.128 ----------------
template <int U>
__global__ void __launch_bounds__(1024, 1) k_ldg(const uint4* src, size_t per_cta_vec, unsigned long long* sink) {
  const uint4* p = src + (size_t)blockIdx.x * per_cta_vec;
  uint32_t acc = 0;
  for (size_t i = threadIdx.x; i + (U - 1) * blockDim.x < per_cta_vec; i += (size_t)U * blockDim.x) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + (size_t)u * blockDim.x));
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// ---------------- mode 2: LDGSTS ----------------
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k_ldgsts(const uint4* src, size_t per_cta_vec, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint4* p = src + (size_t)blockIdx.x * per_cta_vec;
  const uint32_t base = smem_u32(sm) + threadIdx.x * 16;
  const size_t iters = per_cta_vec / blockDim.x;
  for (size_t j = 0; j < iters + DEPTH - 1; ++j) {
    if (j < iters)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)(j % DEPTH) * blockDim.x * 16), "l"(p + j * blockDim.x + threadIdx.x) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
  }
  if (sm[threadIdx.x] == 77 && sm[threadIdx.x + 1] == 99 && sm[3] == 1) *sink = 1;
}

// ---------------- mode 3: 2-D tensor TMA ----------------
__global__ void __launch_bounds__(64, 1) k_tensor(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int box_rows, int N, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t bars = smem_u32(sm);
  const uint32_t data = bars + 1024;
  const int stage_bytes = box_rows * 256;
  if (threadIdx.x == 0) {
    for (int i = 0; i < N; ++i) { mbar_init(bars + i * 8, 1); mbar_init(bars + 256 + i * 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int iters = rows_per_cta / box_rows;
  const int row0 = blockIdx.x * rows_per_cta;
  if (threadIdx.x == 0) {
    for (int j = 0; j < iters; ++j) {
      const int s = j % N;
      if (j >= N) mbar_wait(bars + 256 + s * 8, ((j / N) - 1) & 1);
      mbar_expect(bars + s * 8, stage_bytes);
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(data + s * stage_bytes), "l"(&tm), "r"(0), "r"(row0 + j * box_rows), "r"(bars + s * 8) : "memory");
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (int j = 0; j < iters; ++j) {
      const int s = j % N;
      mbar_wait(bars + s * 8, (j / N) & 1);
      acc += sm[1024 + s * stage_bytes];
      mbar_arrive(bars + 256 + s * 8);
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

template <typename F>
float time_it(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}

int main() {
  const int ctas = 148;
  const size_t per_cta = 6u << 20;  // 6 MiB per CTA = 888 MiB total (>> L2)
  const size_t total = per_cta * ctas;
  uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  printf("total %.0f MB, %d CTAs\n", total / 1e6, ctas);

  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int S : {512, 2048, 4096, 8192, 16384})
    for (int inflight_kb : {32, 64, 128, 192}) {
      const int cps = S >= 8192 ? 1 : 4;  // chunks per stage
      const int stage = S * cps;
      int N = inflight_kb * 1024 / stage; if (N < 2 || N > 30) continue;
      float ms = time_it([&] { k_bulk<<<ctas, 64, 1024 + N * stage>>>(src, per_cta, S, N, cps, sink); });
      printf("bulk1d  S=%5d x%d/stage  stages=%2d inflight=%3d KB : %7.1f GB/s\n", S, cps, N, N * stage / 1024, total / ms / 1e6);
    }
  {
    float ms = time_it([&] { k_ldg<4><<<ctas, 1024>>>((const uint4*)src, per_cta / 16, sink); });
    printf("ldg128  1024 thr unroll 4 (64 KB in flight): %7.1f GB/s\n", total / ms / 1e6);
    ms = time_it([&] { k_ldg<8><<<ctas, 1024>>>((const uint4*)src, per_cta / 16, sink); });
    printf("ldg128  1024 thr unroll 8 (128 KB in flight): %7.1f GB/s\n", total / ms / 1e6);
    ms = time_it([&] { k_ldg<2><<<ctas, 1024>>>((const uint4*)src, per_cta / 16, sink); });
    printf("ldg128  1024 thr unroll 2 (32 KB in flight): %7.1f GB/s\n", total / ms / 1e6);
    ms = time_it([&] { k_ldg<8><<<ctas * 2, 512>>>((const uint4*)src, per_cta / 32, sink); });
    printf("ldg128  2x512 thr unroll 8: %7.1f GB/s\n", total / ms / 1e6);
  }
  {
    CK(cudaFuncSetAttribute(k_ldgsts<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_ldgsts<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float ms = time_it([&] { k_ldgsts<8><<<ctas, 512, 8 * 512 * 16>>>((const uint4*)src, per_cta / 16, sink); });
    printf("ldgsts  512 thr depth 8 (64 KB in flight): %7.1f GB/s\n", total / ms / 1e6);
    ms = time_it([&] { k_ldgsts<16><<<ctas, 512, 16 * 512 * 16>>>((const uint4*)src, per_cta / 16, sink); });
    printf("ldgsts  512 thr depth 16 (128 KB in flight): %7.1f GB/s\n", total / ms / 1e6);
  }
  {
    // 2-D view: rows of 256 B
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn enc = (EncodeFn)fn;
    CK(cudaFuncSetAttribute(k_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const uint64_t rows = total / 256;
    for (int box_rows : {16, 32, 64, 128}) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {64, rows};            // 64 x uint32 = 256 B per row
      cuuint64_t gstr[1] = {256};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      for (int inflight_kb : {64, 128, 192}) {
        int N = inflight_kb * 1024 / (box_rows * 256); if (N < 2 || N > 30) continue;
        float ms = time_it([&] { k_tensor<<<ctas, 64, 1024 + N * box_rows * 256>>>(tm, (int)(rows / ctas), box_rows, N, sink); });
        printf("tma2d   box=%3d x 256 B (%2d KB) stages=%2d inflight=%3d KB : %7.1f GB/s\n", box_rows, box_rows / 4, N, N * box_rows / 4, total / ms / 1e6);
      }
    }
  }
  return 0;
}
