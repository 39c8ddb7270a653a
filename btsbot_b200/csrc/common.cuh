// Shared helpers for the btsbot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/btsbot_b200.h"

namespace btsb {

// ---- error plumbing (thread-local message, process-wide launch counter) --------------------------
void set_error(const char* fmt, ...);
int check_device();
extern std::atomic<uint64_t> g_launches;

inline int launch_done(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return BTSB_ECUDA;
  }
  return BTSB_OK;
}

#define BTSB_REQUIRE(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      btsb::set_error(__VA_ARGS__); \
      return BTSB_EINVAL;         \
    }                             \
  } while (0)

#define BTSB_CUDA(call, what)                                              \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) {                                              \
      btsb::set_error("%s: %s", what, cudaGetErrorString(e__));            \
      return BTSB_ECUDA;                                                   \
    }                                                                      \
  } while (0)

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ldf(const double* p) { return (float)*p; }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == BTSB_ACT_GELU) return gelu_erf(x);
  if (act == BTSB_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}

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

constexpr float kLnEps = 1e-6f;  // timm LayerNorm2d eps for ConvNeXt

}  // namespace btsb
