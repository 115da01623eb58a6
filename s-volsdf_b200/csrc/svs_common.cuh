// Shared host/device helpers of libsvolsdf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "svs.h"

namespace svs {

void set_error(const char* fmt, ...);

#define SVS_CHECK_ARG(cond, ...)     \
  do {                               \
    if (!(cond)) {                   \
      svs::set_error(__VA_ARGS__);   \
      return SVS_ERR_INVALID;        \
    }                                \
  } while (0)

#define SVS_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      svs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return SVS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

void count_launch();
bool prof_enabled();
int prof_begin(const char* name, double flops, double bytes, cudaStream_t st);
void prof_end(int id, cudaStream_t st);

// RAII profiling range around one kernel launch (no-op unless svs_prof_enable(1))
struct ProfScope {
  int id;
  cudaStream_t st;
  ProfScope(const char* name, double flops, double bytes, cudaStream_t s)
      : id(prof_enabled() ? prof_begin(name, flops, bytes, s) : -1), st(s) {}
  ~ProfScope() {
    if (id >= 0) prof_end(id, st);
  }
};

#define SVS_LAUNCH_OK()                \
  do {                                 \
    svs::count_launch();               \
    SVS_CUDA_OK(cudaGetLastError());   \
  } while (0)

#define SVS_TRY(expr)          \
  do {                         \
    int rc_ = (expr);          \
    if (rc_ != 0) return rc_;  \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace svs
