// Shared host/device helpers for the tinygemm_b200 C-ABI library.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/tinygemm_b200.h"

namespace tg {

// ---- error reporting -------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TG_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::tg::set_error(__VA_ARGS__);           \
      return TG_ERR_INVALID_ARGUMENT;         \
    }                                         \
  } while (0)

#define TG_CHECK_LAUNCH(what)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::tg::set_error("%s: CUDA error %s", what, cudaGetErrorString(e__));           \
      return TG_ERR_CUDA;                                                            \
    }                                                                                \
    ::tg::count_launch();                                                            \
  } while (0)

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Function attributes (opt-in shared memory size) and occupancy are per DEVICE: one-time setup flags are kept per
// thread and per device ordinal, so a process that drives several GPUs (CUDAGuard in the torch op layer) works too.
constexpr int kMaxDevices = 64;
static inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
  return d % kMaxDevices;
}

// ---- internal launchers shared between translation units ---------------------------
// (all return TG_OK / TG_ERR_*)
int launch_gemm_w4_rm(void* y, const void* x, const int32_t* w, const void* sz, const void* lut,
                      const uint8_t* exps, int64_t rows_x, int64_t w_rows, int64_t k, int group, int ik,
                      tg_w4_format fmt, tg_weight_side side, tg_dtype dt, cudaStream_t st);
int launch_gemm_w8_rm(void* y, const void* x, const int32_t* w, const void* sz, int64_t rows_x,
                      int64_t w_rows, int64_t k, int group, int ik, tg_weight_side side, tg_dtype dt,
                      cudaStream_t st);
int launch_gemm_w16_rm(void* y, const void* x, const void* w, int64_t rows_x, int64_t w_rows, int64_t k,
                       int ik, tg_weight_side side, tg_dtype dt, cudaStream_t st);

}  // namespace tg
