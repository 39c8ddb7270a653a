// K7 (bf16 mode): operand preparation for the tensor-core training GEMMs of gemm_tc.cu.
//
// The mixed-precision training step (what torch autocast(bf16) does to btsbot/train.py:496-547) keeps the residual
// stream, LayerNorm and the element-wise backward in fp32 and runs the three GEMMs of every Linear / 1x1 conv
// (forward, dgrad, wgrad -- 92 % of the FLOPs) on tcgen05 with bf16 operands and fp32 accumulation.  The wgrad GEMM
// contracts over the ROW dimension of the activations (B*H*W, up to ~2 M), so it wants both operands "transposed":
// one pass over the fp32 tensor therefore emits
//     rm [M, N]  bf16   (A operand of the forward / dgrad GEMM)
//     t  [N, ld] bf16   (operand of the wgrad GEMM, ld = M rounded up to 8 so that the TMA row pitch is 16-byte aligned)
// with the element-wise op that would otherwise be its own kernel folded into the load:
//     OP 0: v = in                         OP 1: v = gelu(in)                (hidden activation, never stored in fp32)
//     OP 2: v = in2 * gelu'(in)            (in = saved pre-activation, in2 = upstream gradient)
// `in` / `in2` are fp32 or bf16 (the bf16-output GEMM epilogue writes the 4C-wide tensors in bf16, as autocast keeps them)
// then v *= colvec[n] (layer-scale backward) and colsum[n] += sum_m v (bias gradients) when those pointers are given.
#include "common.cuh"

namespace btsb {
namespace {

// GELU and its derivative in the same one-MUFU form the bf16 inference epilogues use (tc_common.cuh gelu_fast):
// gelu(x) = 0.5 x (1 + tanh(a(x))), a(x) = x (k1 + k3 x^2 + k5 x^4) with x^2 clamped at 64 (formula error 2.5e-5, tanh.approx
// 2^-11) -- the erff / expf versions made this kernel ALU-bound (~40 instructions per element at 2.3 TB/s); forward and
// backward use the SAME function, so the gradient is exact for the function that was evaluated.
//   gelu'(x) = 0.5 (1 + t) + 0.5 x (1 - t^2) a'(x),   a'(x) = k1 + 3 k3 x^2 + 5 k5 x^4   (= the frozen slope beyond |x| = 8)
constexpr float kG1 = 0.7975078843f, kG3 = 0.037005646f, kG5 = -3.5151679e-4f;
__device__ __forceinline__ float tanh_fast(float a) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(a));
  return t;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = fminf(x * x, 64.0f);
  const float p = fmaf(x2, fmaf(x2, kG5, kG3), kG1);
  const float h = 0.5f * x;
  return fmaf(h, tanh_fast(p * x), h);
}
__device__ __forceinline__ float gelu_grad_tanh(float x) {
  const float xx = x * x;
  const float x2 = fminf(xx, 64.0f);
  const float p = fmaf(x2, fmaf(x2, kG5, kG3), kG1);
  const float t = tanh_fast(p * x);
  const float da = xx > 64.0f ? p : fmaf(x2, fmaf(x2, 5.0f * kG5, 3.0f * kG3), kG1);
  return fmaf(0.5f * x * da, fmaf(-t, t, 1.0f), 0.5f + 0.5f * t);
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int TM = 64, TN = 64;

// IN16: `in` (and, for OP 2, `in2`) are bf16 -- the wide tensors of the step (fc1 pre-activation, its upstream gradient)
// are written by the bf16-output GEMM epilogue (bulk tensor stores, half the bytes) exactly as autocast would keep them
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}

template <int OP, typename TI>
__global__ void __launch_bounds__(256)
cast_dual_kernel(const TI* __restrict__ in, const TI* __restrict__ in2, const float* __restrict__ colvec,
                 __nv_bfloat16* __restrict__ rm, __nv_bfloat16* __restrict__ t, float* __restrict__ colsum, int64_t M, int N,
                 int64_t ld, const float* __restrict__ aux, float* __restrict__ auxsum) {
  __shared__ float tile[TN][TM + 1];        // [n][m]
  __shared__ float red[8][TN];
  __shared__ float red2[8][TN];
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = n0 + 2 * tx;
  const bool ncol = n < N;                  // N is even: n + 1 < N as well
  float g0 = 1.f, g1 = 1.f;
  if (colvec && ncol) { g0 = __ldg(colvec + n); g1 = __ldg(colvec + n + 1); }
  float s0 = 0.f, s1 = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < TM / 8; ++i) {
    const int m = ty + 8 * i;
    const int64_t gm = m0 + m;
    float2 v = make_float2(0.f, 0.f);
    if (gm < M && ncol) {
      v = ld2(in + gm * N + n);
      if (auxsum) {                           // auxsum[n] += sum_m in[m,n] * aux[m,n]  (layer-scale gradient: dcur . v)
        const float2 x = *reinterpret_cast<const float2*>(aux + gm * N + n);
        a0 = fmaf(v.x, x.x, a0); a1 = fmaf(v.y, x.y, a1);
      }
      if (OP == 1) { v.x = gelu_tanh(v.x); v.y = gelu_tanh(v.y); }
      if (OP == 2) {
        const float2 d = ld2(in2 + gm * N + n);
        v.x = d.x * gelu_grad_tanh(v.x); v.y = d.y * gelu_grad_tanh(v.y);
      }
      v.x *= g0; v.y *= g1;
      if (rm) *reinterpret_cast<uint32_t*>(rm + gm * N + n) = pack2(v.x, v.y);
    }
    s0 += v.x; s1 += v.y;
    tile[2 * tx][m] = v.x;
    tile[2 * tx + 1][m] = v.y;
  }
  if (colsum) { red[ty][2 * tx] = s0; red[ty][2 * tx + 1] = s1; }
  if (auxsum) { red2[ty][2 * tx] = a0; red2[ty][2 * tx + 1] = a1; }
  __syncthreads();
  if (colsum && threadIdx.x < TN && n0 + (int)threadIdx.x < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(colsum + n0 + threadIdx.x, s);
  }
  if (auxsum && threadIdx.x >= 64 && threadIdx.x < 64 + TN && n0 + (int)threadIdx.x - 64 < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red2[w][threadIdx.x - 64];
    atomicAdd(auxsum + n0 + threadIdx.x - 64, s);
  }
  if (t) {
    const int64_t gm = m0 + 2 * tx;         // ld is even and >= M rounded up to 8: the pair never leaves its row
#pragma unroll
    for (int i = 0; i < TN / 8; ++i) {
      const int nn = ty + 8 * i;
      const int gn = n0 + nn;
      if (gn < N && gm < M) {
        const float a = tile[nn][2 * tx];
        const float b = gm + 1 < M ? tile[nn][2 * tx + 1] : 0.f;
        *reinterpret_cast<uint32_t*>(t + (int64_t)gn * ld + gm) = pack2(a, b);
      }
    }
  }
}

}  // namespace
}  // namespace btsb

using namespace btsb;

extern "C" int btsb_cast_dual_bf16(const void* in, const void* in2, const float* colvec, void* out_rm, void* out_t,
                                   float* colsum, int64_t M, int N, int64_t ld, int op, int in_dtype, const float* aux,
                                   float* auxsum, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(M >= 0 && N >= 2 && N % 2 == 0, "cast_dual: N=%d must be even", N);
  if (M == 0) return BTSB_OK;
  BTSB_REQUIRE(in && (out_rm || out_t || colsum), "cast_dual: null pointer");
  BTSB_REQUIRE((aux == nullptr) == (auxsum == nullptr) && (!aux || ((uintptr_t)aux % 8) == 0), "cast_dual: aux / auxsum go together");
  BTSB_REQUIRE(op >= 0 && op <= 2 && (op != 2 || in2), "cast_dual: bad op %d", op);
  BTSB_REQUIRE(in_dtype == BTSB_F32 || in_dtype == BTSB_BF16, "cast_dual: in_dtype must be F32 or BF16");
  BTSB_REQUIRE(!out_t || (ld % 8 == 0 && ld >= M), "cast_dual: ld=%lld must be a multiple of 8 and >= M", (long long)ld);
  BTSB_REQUIRE(((uintptr_t)in % 8) == 0 && (!in2 || ((uintptr_t)in2 % 8) == 0) && ((uintptr_t)out_rm % 4) == 0 &&
                   ((uintptr_t)out_t % 4) == 0, "cast_dual: misaligned pointer");
  const int64_t gx = (M + TM - 1) / TM;
  BTSB_REQUIRE(gx < (1ll << 31) && (N + TN - 1) / TN <= 65535, "cast_dual: shape too large");
  dim3 grid((unsigned)gx, (unsigned)((N + TN - 1) / TN));
  __nv_bfloat16* rm = (__nv_bfloat16*)out_rm;
  __nv_bfloat16* t = (__nv_bfloat16*)out_t;
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == BTSB_F32) {
    const float* a = (const float*)in; const float* b = (const float*)in2;
    if (op == 0) cast_dual_kernel<0, float><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
    else if (op == 1) cast_dual_kernel<1, float><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
    else cast_dual_kernel<2, float><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
  } else {
    const __nv_bfloat16* a = (const __nv_bfloat16*)in; const __nv_bfloat16* b = (const __nv_bfloat16*)in2;
    if (op == 0) cast_dual_kernel<0, __nv_bfloat16><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
    else if (op == 1) cast_dual_kernel<1, __nv_bfloat16><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
    else cast_dual_kernel<2, __nv_bfloat16><<<grid, 256, 0, st>>>(a, b, colvec, rm, t, colsum, M, N, ld, aux, auxsum);
  }
  return launch_done("cast_dual_bf16");
}
