// Microbenchmark: throughput of GELU formulations on registers (no memory), to find the practical epilogue ceiling.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/micro_gelu scripts/micro/micro_gelu.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float gelu_ex2rcp(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(x2, 1.01426306e-3f, -0.10677572f);
  p = fmaf(x2, p, -2.3011213f);
  float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * x));
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// x * 0.5 * (1 + tanh(q(x))), q = -p/ (2 log2 e) ... same odd quintic expressed for tanh: sigmoid(2q) = 0.5(1+tanh(q))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(x2, -3.5151679e-4f, 0.037005646f);
  p = fmaf(x2, p, 0.7975078843f);
  float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
__device__ __forceinline__ float fma_only(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(x2, 1.01426306e-3f, -0.10677572f);
  p = fmaf(x2, p, -2.3011213f);
  float e = p * x;
  float r = 1.0f + e;
  return x * r;
}
__device__ __forceinline__ float ex2_only(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float rcp_only(float x) { float e; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float tanh_only(float x) { float e; asm("tanh.approx.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }

// two values per MUFU: ex2 on packed halves, rcp still per value
__device__ __forceinline__ void gelu_h2(float xa, float xb, float& oa, float& ob) {
  const float a2 = fminf(xa * xa, 64.0f), b2 = fminf(xb * xb, 64.0f);
  float pa = fmaf(a2, 1.01426306e-3f, -0.10677572f), pb = fmaf(b2, 1.01426306e-3f, -0.10677572f);
  pa = fmaf(a2, pa, -2.3011213f) * xa; pb = fmaf(b2, pb, -2.3011213f) * xb;
  __half2 h = __floats2half2_rn(pa, pb);
  uint32_t hv = *reinterpret_cast<uint32_t*>(&h), ev;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(ev) : "r"(hv));
  __half2 e = *reinterpret_cast<__half2*>(&ev);
  const float2 ef = __half22float2(e);
  float ra, rb;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(1.0f + ef.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(1.0f + ef.y));
  oa = xa * ra; ob = xb * rb;
}
// two values per MUFU: packed tanh in bf16x2
__device__ __forceinline__ void gelu_tanh_h2(float xa, float xb, float& oa, float& ob) {
  const float a2 = fminf(xa * xa, 64.0f), b2 = fminf(xb * xb, 64.0f);
  float pa = fmaf(a2, -3.5151679e-4f, 0.037005646f), pb = fmaf(b2, -3.5151679e-4f, 0.037005646f);
  pa = fmaf(a2, pa, 0.7975078843f) * xa; pb = fmaf(b2, pb, 0.7975078843f) * xb;
  __half2 h = __floats2half2_rn(pa, pb);
  uint32_t hv = *reinterpret_cast<uint32_t*>(&h), tv;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tv) : "r"(hv));
  const float2 tf = __half22float2(*reinterpret_cast<__half2*>(&tv));
  const float ha = 0.5f * xa, hb = 0.5f * xb;
  oa = fmaf(ha, tf.x, ha); ob = fmaf(hb, tf.y, hb);
}

template <int V>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + 0.01f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (V == 0) { v[i] = gelu_ex2rcp(v[i]) + 0.3f; v[i + 1] = gelu_ex2rcp(v[i + 1]) + 0.3f; }
      if (V == 1) { v[i] = gelu_tanh(v[i]) + 0.3f; v[i + 1] = gelu_tanh(v[i + 1]) + 0.3f; }
      if (V == 2) { v[i] = fma_only(v[i]) * 0.5f; v[i + 1] = fma_only(v[i + 1]) * 0.5f; }
      if (V == 3) { v[i] = ex2_only(v[i]) * 0.25f; v[i + 1] = ex2_only(v[i + 1]) * 0.25f; }
      if (V == 4) { v[i] = rcp_only(v[i]) + 1.0f; v[i + 1] = rcp_only(v[i + 1]) + 1.0f; }
      if (V == 5) { v[i] = tanh_only(v[i]) + 0.5f; v[i + 1] = tanh_only(v[i + 1]) + 0.5f; }
      if (V == 6) { float a, b; gelu_h2(v[i], v[i + 1], a, b); v[i] = a + 0.3f; v[i + 1] = b + 0.3f; }
      if (V == 7) { float a, b; gelu_tanh_h2(v[i], v[i + 1], a, b); v[i] = a + 0.3f; v[i + 1] = b + 0.3f; }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;
}

template <int V>
void run(const char* name, int threads, float* d) {
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<148, threads>>>(d, 10, 0.1f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<V><<<148, threads>>>(d, iters, 0.1f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = 148.0 * threads * 16.0 * iters;
  printf("%-22s threads/SM %4d: %7.3f ms  %6.2f elem/clk/SM (at 1.965 GHz)  %6.1f clk per warp-elem/SMSP\n", name, threads, ms,
         n / 148.0 / (ms * 1e-3 * 1.965e9), (ms * 1e-3 * 1.965e9) / (n / 148.0 / 32.0 / 4.0));
}

int main() {
  float* d; cudaMalloc(&d, 1024);
  for (int threads : {512, 1024}) {
    run<0>("gelu ex2+rcp", threads, d);
    run<1>("gelu tanh.f32", threads, d);
    run<2>("fma part only", threads, d);
    run<3>("ex2 only", threads, d);
    run<4>("rcp only", threads, d);
    run<5>("tanh only", threads, d);
    run<6>("gelu ex2.f16x2+rcp", threads, d);
    run<7>("gelu tanh.f16x2", threads, d);
  }
  // accuracy of the variants against erf GELU is checked in tests, not here
  return 0;
}
