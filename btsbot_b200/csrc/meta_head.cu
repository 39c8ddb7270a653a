// K6 -- metadata branch + fusion head as ONE kernel (architectures.py:146-164,168-170; um_nn :282-290;
// image-only heads :109-119; frozen_fusion :357-365).  fp32 FFMA; a CTA owns TA=32 alerts whose activations
// live in shared memory as [k][alert] so each weight (read once, coalesced, from L2) feeds 8 FMAs.
#include "common.cuh"

namespace btsb {

constexpr int kTA = 32;    // alerts per CTA
constexpr int kTAp = 36;   // padded alert stride (16B-aligned rows)
constexpr int kAT = 8;     // alerts per thread (register block)
constexpr int kHeadThreads = 256;

__device__ __forceinline__ void dense_layer(const float* __restrict__ in, int K, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, int N, int act, float* __restrict__ out) {
  const int groups = kTA / kAT;
  for (int it = threadIdx.x; it < N * groups; it += kHeadThreads) {
    const int n = it % N, ag = it / N;
    float acc[kAT];
    const float bv = bias ? __ldg(bias + n) : 0.f;
#pragma unroll
    for (int a = 0; a < kAT; ++a) acc[a] = bv;
    const float* src = in + ag * kAT;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(Wt + (size_t)k * N + n);
      const float4 v0 = *reinterpret_cast<const float4*>(src + k * kTAp);
      const float4 v1 = *reinterpret_cast<const float4*>(src + k * kTAp + 4);
      acc[0] = fmaf(w, v0.x, acc[0]); acc[1] = fmaf(w, v0.y, acc[1]);
      acc[2] = fmaf(w, v0.z, acc[2]); acc[3] = fmaf(w, v0.w, acc[3]);
      acc[4] = fmaf(w, v1.x, acc[4]); acc[5] = fmaf(w, v1.y, acc[5]);
      acc[6] = fmaf(w, v1.z, acc[6]); acc[7] = fmaf(w, v1.w, acc[7]);
    }
    float* dst = out + n * kTAp + ag * kAT;
#pragma unroll
    for (int a = 0; a < kAT; ++a) dst[a] = apply_act(acc[a], act);
  }
}

__global__ void __launch_bounds__(kHeadThreads)
meta_head_kernel(btsb_head_params p, int64_t B, float* __restrict__ logits) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * kTA;
  const int na = (int)min((int64_t)kTA, B - b0);
  const int F = p.F, Mm = p.Mm, m1 = p.m1, m2 = p.m2, c1 = p.c1, c2 = p.c2;
  const int emb = (Mm > 0) ? m2 : 0;
  float* cat = sm;                                   // [F+emb][kTAp]
  float* bufM = cat + (size_t)(F + emb) * kTAp;      // [Mm][kTAp]
  float* buf1 = bufM + (size_t)Mm * kTAp;            // [m1][kTAp]
  float* bufH0 = buf1 + (size_t)m1 * kTAp;           // [c1][kTAp]
  float* bufH1 = bufH0 + (size_t)c1 * kTAp;          // [c2][kTAp]

  // stage inputs (zero-fill alerts beyond the batch tail)
  if (F > 0) {
    for (int i = tid; i < F * kTA; i += kHeadThreads) {
      const int a = i / F, k = i - a * F;
      float v = 0.f;
      if (a < na) {
        v = p.feat_dtype == BTSB_BF16
                ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.feat)[(b0 + a) * F + k])
                : reinterpret_cast<const float*>(p.feat)[(b0 + a) * F + k];
      }
      cat[k * kTAp + a] = v;
    }
  }
  if (Mm > 0) {
    for (int i = tid; i < Mm * kTA; i += kHeadThreads) {
      const int a = i / Mm, k = i - a * Mm;
      float v = 0.f;
      if (a < na) v = fmaf(p.meta[(b0 + a) * Mm + k], __ldg(p.bn_scale + k), __ldg(p.bn_shift + k));   // folded BatchNorm1d (eval)
      bufM[k * kTAp + a] = v;
    }
  }
  __syncthreads();
  if (Mm > 0) {
    dense_layer(bufM, Mm, p.m1t, p.m1b, m1, p.meta_act, buf1);
    __syncthreads();
    dense_layer(buf1, m1, p.m2t, p.m2b, m2, p.meta_out_act, cat + (size_t)F * kTAp);
    __syncthreads();
  }
  const float* last = cat;
  int lastK = F + emb;
  if (c1 > 0) {
    dense_layer(cat, F + emb, p.h0t, p.h0b, c1, p.head_act, bufH0);
    __syncthreads();
    dense_layer(bufH0, c1, p.h1t, p.h1b, c2, p.head_act, bufH1);
    __syncthreads();
    last = bufH1;
    lastK = c2;
  }
  if (tid < na) {
    float acc = __ldg(p.h2b);
    for (int k = 0; k < lastK; ++k) acc = fmaf(last[k * kTAp + tid], __ldg(p.h2 + k), acc);
    logits[b0 + tid] = acc;
  }
}

}  // namespace btsb

using namespace btsb;

extern "C" int btsb_meta_head_fwd(const btsb_head_params* pp, int64_t B, float* logits, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(pp, "meta_head: null params");
  const btsb_head_params& p = *pp;
  BTSB_REQUIRE(B >= 0, "meta_head: B < 0");
  BTSB_REQUIRE(p.F >= 0 && p.Mm >= 0 && (p.F > 0 || p.Mm > 0), "meta_head: need features and/or metadata");
  BTSB_REQUIRE(p.F == 0 || (p.feat && (p.feat_dtype == BTSB_F32 || p.feat_dtype == BTSB_BF16)), "meta_head: bad feat");
  if (p.Mm > 0)
    BTSB_REQUIRE(p.meta && p.bn_scale && p.bn_shift && p.m1t && p.m1b && p.m2t && p.m2b && p.m1 > 0 && p.m2 > 0,
                 "meta_head: metadata branch pointers/sizes missing");
  if (p.c1 > 0)
    BTSB_REQUIRE(p.h0t && p.h0b && p.h1t && p.h1b && p.c2 > 0, "meta_head: head pointers/sizes missing");
  else
    BTSB_REQUIRE(p.F == 0 && p.Mm > 0, "meta_head: c1==0 is only valid for the metadata-only network");
  BTSB_REQUIRE(p.h2 && p.h2b, "meta_head: final layer missing");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(logits, "meta_head: null logits");
  const int emb = p.Mm > 0 ? p.m2 : 0;
  const size_t smem = (size_t)(p.F + emb + p.Mm + (p.Mm > 0 ? p.m1 : 0) + p.c1 + p.c2) * kTAp * sizeof(float);
  BTSB_REQUIRE(smem <= 227 * 1024, "meta_head: layer widths need %zu B of shared memory (> 227 KB)", smem);
  BTSB_CUDA(cudaFuncSetAttribute(meta_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "meta_head attr");
  const int64_t grid = (B + kTA - 1) / kTA;
  meta_head_kernel<<<(unsigned)grid, kHeadThreads, smem, (cudaStream_t)stream>>>(p, B, logits);
  return launch_done("meta_head");
}
