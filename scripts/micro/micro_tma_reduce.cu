// Microbenchmark / probe: does cp.reduce.async.bulk.tensor (.add) work on a bf16 tensor map with a [32 rows x 16 cols] box,
// with and without the 32-byte swizzle, and on fp32?  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// scripts/micro/micro_tma_reduce scripts/micro/micro_tma_reduce.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename T>
__global__ void k(const __grid_constant__ CUtensorMap tm, int use_reduce) {
  __shared__ __align__(1024) unsigned char slab[2048];
  T* s = reinterpret_cast<T*>(slab);
  for (int i = threadIdx.x; i < 32 * 16; i += 32) s[i] = (T)2.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(slab);
    if (use_reduce)
      asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&tm), "r"(src), "r"(16), "r"(0) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&tm), "r"(src), "r"(16), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
template <typename T>
void probe(const char* name, CUtensorMapDataType dt, CUtensorMapSwizzle sw, EncodeTiledFn enc) {
  const int rows = 64, cols = 64;
  T* d; cudaMalloc(&d, rows * cols * sizeof(T));
  T* h = new T[rows * cols];
  for (int i = 0; i < rows * cols; ++i) h[i] = (T)1.0f;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemcpy(d, h, rows * cols * sizeof(T), cudaMemcpyHostToDevice);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {cols, rows}; const cuuint64_t strides[1] = {cols * sizeof(T)};
    const cuuint32_t box[2] = {16, 32}; const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    k<T><<<1, 32>>>(tm, mode);
    cudaError_t e = cudaDeviceSynchronize();
    T* o = new T[rows * cols];
    cudaMemcpy(o, d, rows * cols * sizeof(T), cudaMemcpyDeviceToHost);
    printf("%-22s %-6s: %s  out[0][16]=%g out[31][31]=%g out[0][15]=%g out[32][16]=%g\n", name, mode ? "reduce" : "store",
           cudaGetErrorString(e), (float)o[16], (float)o[31 * cols + 31], (float)o[15], (float)o[32 * cols + 16]);
    delete[] o;
    if (e != cudaSuccess) return;
  }
  cudaFree(d); delete[] h;
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encoder\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  probe<float>("fp32 no swizzle", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE, enc);
  probe<__nv_bfloat16>("bf16 no swizzle", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_NONE, enc);
  probe<__nv_bfloat16>("bf16 swizzle 32B", CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_32B, enc);
  return 0;
}
