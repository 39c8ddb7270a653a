// Shared helpers for the btsbot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// Kernel attributes (opt-in shared memory above 48 KB) and the SM count are per DEVICE: a process that drives several
// GPUs must set them on each.  `first_use_on_device(mask)` is true once per (call site, device).
inline bool first_use_on_device(std::atomic<uint64_t>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  return (mask.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
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
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
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

// Sum 32 per-lane values across an aligned group of WIDTH lanes; afterwards v[0 .. 32/WIDTH) of lane l hold the group
// totals of value indices (l % WIDTH) * (32/WIDTH) + i.
template <int WIDTH>
__device__ __forceinline__ void seg_reduce32(float (&v)[32], int lane) {
  int n = 32;
#pragma unroll
  for (int off = WIDTH / 2; off >= 1; off >>= 1) {
    n >>= 1;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < n) {
        const float lo = v[i], hi = v[i + n];
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
}

// ---- packed fp32 pairs (sm_100 FFMA2): one instruction = two FMAs.  The depthwise-conv kernels are instruction-issue
// bound with FFMA at 57 % of their instruction mix (profiles/r01f/dwln15), and their natural unit is the channel PAIR
// (bf16x2 input, float2 taps), so (acc_c, acc_c+1) += (w_c, w_c+1) * (x_c, x_c+1) is exactly one fma.rn.f32x2.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack_f32x2(f32x2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
// bf16x2 word -> (float(lo half), float(hi half)) packed
__device__ __forceinline__ f32x2_t bf16x2_to_f32x2(uint32_t v) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(v << 16), "r"(v & 0xffff0000u));
  return r;
}
__device__ __forceinline__ void fma_f32x2(f32x2_t& acc, f32x2_t a, f32x2_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2_t fma3_f32x2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t mul_f32x2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2_t add_f32x2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ---- the residual stream's storage type ------------------------------------------------------------------------------
// In bf16 mode the tensors between blocks (stem output, block outputs, downsample outputs) can be stored as IEEE fp16
// instead of bf16 (dtype code BTSB_BF16_XF16): same bytes, 11 instead of 8 significand bits.  The block update is added
// to this tensor twelve to fourteen times in a row, and its rounding was the largest single contribution to the bf16
// mode's logit error (CPU emulation rounding at the kernels' points: 1.4e-2 of a 1.4-2.2e-2 total on the worst parity
// case, 0.7-1.1e-2 with an fp16 or fp32 stream).  MMA operands (LayerNorm outputs, hidden
// activations, weights) stay bf16.  Stores saturate to the largest finite fp16 instead of overflowing to infinity.
template <bool XF16>
__device__ __forceinline__ uint32_t pack_x2(float lo, float hi) {
  uint32_t r;
  if constexpr (XF16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool XF16>
__device__ __forceinline__ float x2_lo(uint32_t v) {
  if constexpr (XF16) { float f; asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, l;}" : "=f"(f) : "r"(v)); return f; }
  else return __uint_as_float(v << 16);
}
template <bool XF16>
__device__ __forceinline__ float x2_hi(uint32_t v) {
  if constexpr (XF16) { float f; asm("{.reg .b16 l, h; mov.b32 {l, h}, %1; cvt.f32.f16 %0, h;}" : "=f"(f) : "r"(v)); return f; }
  else return __uint_as_float(v & 0xffff0000u);
}
template <bool XF16>
__device__ __forceinline__ f32x2_t x2_to_f32x2(uint32_t v) { return pack_f32x2(x2_lo<XF16>(v), x2_hi<XF16>(v)); }
// (The depthwise-conv kernels convert with the same two cvt per pair.  An ALU-only alternative -- shift the 16 bits into
// fp32 position, which yields value * 2^-112 exactly, and fold 2^112 into the taps -- was built on the assumption that
// HADD2.F32 would load the FMA pipe those kernels are bound by; measured, the cvt form is the faster one: dwln_15x80
// 360 vs 391 us, dwln_7x160 156 vs 184 us, a mixed form 370 / 165 (profiles/r02cv).)

constexpr float kLnEps = 1e-6f;  // timm LayerNorm2d eps for ConvNeXt

}  // namespace btsb
