// Microbenchmark: issue/throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/micro_fma2 scripts/micro/micro_fma2.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int V>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[16], b[16]; unsigned long long p[8], q[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = 1.0f - 0.0001f * (threadIdx.x % 7 + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(q[i]) : "f"(b[2 * i]), "f"(b[2 * i + 1]));
  const float w = 0.999f, c = 0.001f;
  unsigned long long w2, c2;
  asm("mov.b64 %0, {%1,%2};" : "=l"(w2) : "f"(w), "f"(w));
  asm("mov.b64 %0, {%1,%2};" : "=l"(c2) : "f"(c), "f"(c));
  for (int it = 0; it < iters; ++it) {
    if (V == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w, c);
    } else if (V == 2) {
      // conv-like: acc[i] += wv[i] * x, x changes per step, wv loop-invariant registers (three distinct register pairs)
      unsigned long long x2;
      asm volatile("mov.b64 %0, {%1,%2};" : "=l"(x2) : "f"(seed + it), "f"(seed - it));
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i]), "l"(x2));
    } else if (V == 3) {
      // the same pattern with scalar FFMA (16 accumulators, 16 weights in registers)
      const float x = seed + it;
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(x));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(w2), "l"(c2));
    }
  }
  float s = 0;
  if (V == 0 || V == 3) { for (int i = 0; i < 16; ++i) s += a[i]; }
  else { for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += x + y; } }
  if (s == 123.456f) out[0] = s;
}
template <int V> void run(const char* name, int threads, float* d) {
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<148, threads>>>(d, 10, 0.1f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<V><<<148, threads>>>(d, iters, 0.1f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = 148.0 * threads * 16.0 * iters;
  printf("%-8s threads/SM %4d: %7.3f ms  %6.1f FMA/clk/SM (at 1.965 GHz)  %5.1f TFLOP/s\n", name, threads, ms,
         fma / 148.0 / (ms * 1e-3 * 1.965e9), 2.0 * fma / (ms * 1e-3) / 1e12);
}
int main() {
  float* d; cudaMalloc(&d, 1024);
  for (int t : {256, 512, 640, 1024}) {
    run<0>("FFMA", t, d); run<1>("FFMA2", t, d); run<2>("FFMA2cv", t, d); run<3>("FFMAcv", t, d);
  }
  return 0;
}
