// K6 -- metadata branch + fusion head as ONE kernel (architectures.py:146-164,168-170; um_nn :282-290;
// image-only heads :109-119; frozen_fusion :357-365).  fp32 FFMA; a CTA owns TA=32 alerts whose activations
// live in shared memory as [k][alert] so each weight (read once, coalesced, from L2) feeds 8 FMAs.
#include "common.cuh"

namespace btsb {

constexpr int kTA = 32;    // alerts per CTA
constexpr int kTAp = 36;   // padded alert stride (16B-aligned rows)
constexpr int kAT = 8;     // alerts per thread (register block)
constexpr int kHeadThreads = 256;

// One dense layer on the CTA's kTA alerts: out[n][a] = act(bias[n] + sum_k in[k][a] * Wt[k][n]).
// Thread = NPT adjacent outputs x kAT alerts; the weight stream (L2-resident, coalesced) is prefetched one block of
// kWU k-steps ahead into registers so its latency is hidden behind the kWU*NPT*kAT FMAs of the current block.
constexpr int kWU = 8;

template <int NPT>
__device__ __forceinline__ void load_wblock(const float* __restrict__ wp, int N, int k0, int K, float (&w)[kWU][NPT]) {
#pragma unroll
  for (int u = 0; u < kWU; ++u) {
    const int k = min(k0 + u, K - 1);                  // clamped: tail values are loaded but never used
    if (NPT == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(wp + (size_t)k * N));
      w[u][0] = v.x; w[u][NPT - 1] = v.y;
    } else {
      w[u][0] = __ldg(wp + (size_t)k * N);
    }
  }
}

// Initial accumulator of output n for the CTA's alert a: bias[n], or -- when the caller pre-computed part of the
// contraction (`init`, row-major [alerts][N], rows >= na do not exist) -- init[a][n].
__device__ __forceinline__ float acc_init(const float* __restrict__ bias, const float* __restrict__ init, int N, int n,
                                          int a, int na) {
  if (init) return a < na ? __ldg(init + (size_t)a * N + n) : 0.f;
  return bias ? __ldg(bias + n) : 0.f;
}

template <int NPT>
__device__ __forceinline__ void dense_layer_t(const float* __restrict__ in, int K, const float* __restrict__ Wt,
                                              const float* __restrict__ bias, int N, int act, float* __restrict__ out,
                                              const float* __restrict__ init, int na) {
  const int groups = kTA / kAT;
  const int NV = N / NPT;
  for (int it = threadIdx.x; it < NV * groups; it += kHeadThreads) {
    const int n = (it % NV) * NPT, ag = it / NV;
    float acc[NPT][kAT];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
#pragma unroll
      for (int a = 0; a < kAT; ++a) acc[j][a] = acc_init(bias, init, N, n + j, ag * kAT + a, na);
    }
    const float* src = in + ag * kAT;
    const float* wp = Wt + n;
    float wa[kWU][NPT], wb[kWU][NPT];
    auto block = [&](int k0, const float (&w)[kWU][NPT]) {
#pragma unroll
      for (int u = 0; u < kWU; ++u) {
        if (k0 + u < K) {
          const float4 v0 = *reinterpret_cast<const float4*>(src + (k0 + u) * kTAp);
          const float4 v1 = *reinterpret_cast<const float4*>(src + (k0 + u) * kTAp + 4);
#pragma unroll
          for (int j = 0; j < NPT; ++j) {
            acc[j][0] = fmaf(w[u][j], v0.x, acc[j][0]); acc[j][1] = fmaf(w[u][j], v0.y, acc[j][1]);
            acc[j][2] = fmaf(w[u][j], v0.z, acc[j][2]); acc[j][3] = fmaf(w[u][j], v0.w, acc[j][3]);
            acc[j][4] = fmaf(w[u][j], v1.x, acc[j][4]); acc[j][5] = fmaf(w[u][j], v1.y, acc[j][5]);
            acc[j][6] = fmaf(w[u][j], v1.z, acc[j][6]); acc[j][7] = fmaf(w[u][j], v1.w, acc[j][7]);
          }
        }
      }
    };
    load_wblock<NPT>(wp, N, 0, K, wa);
    for (int k0 = 0; k0 < K; k0 += 2 * kWU) {
      if (k0 + kWU < K) load_wblock<NPT>(wp, N, k0 + kWU, K, wb);
      block(k0, wa);
      if (k0 + 2 * kWU < K) load_wblock<NPT>(wp, N, k0 + 2 * kWU, K, wa);
      if (k0 + kWU < K) block(k0 + kWU, wb);
    }
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      float* dst = out + (n + j) * kTAp + ag * kAT;
#pragma unroll
      for (int a = 0; a < kAT; ++a) dst[a] = apply_act(acc[j][a], act);
    }
  }
}

// Weight-streaming variant: the [K][N] weight matrix is pulled through a 3-deep ring of kKT-row tiles in shared memory
// with cp.async (two tiles = 64 k-steps in flight), so the L2 latency of the weight stream is fully hidden; threads
// read their two weights with one conflict-free LDS.64 and the 8 alert activations with two broadcast LDS.128.
// Needs N % 4 == 0 (16-byte cp.async granules) and a 16-byte aligned Wt.  ALL threads of the CTA must call it.
constexpr int kKT = 32;           // k rows per tile
constexpr int kWRing = 3;
constexpr int kWTileMaxN = 128;   // ring sized for N <= 128 (3 * 32 * 128 * 4 B = 48 KB)

__device__ __forceinline__ void cp_async16_h(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void dense_layer_ring(const float* __restrict__ in, int K, const float* __restrict__ Wt,
                                                 const float* __restrict__ bias, int N, int act, float* __restrict__ out,
                                                 float* __restrict__ ring, const float* __restrict__ init, int na) {
  const int groups = kTA / kAT;
  const int NV = N >> 1;
  const int items = NV * groups;
  const int ntiles = (K + kKT - 1) / kKT;
  const int tile_floats = kKT * N;
  auto issue = [&](int t) {
    if (t < ntiles) {
      const int rows = min(kKT, K - t * kKT);
      const float4* src = reinterpret_cast<const float4*>(Wt + (size_t)t * kKT * N);
      float4* dst = reinterpret_cast<float4*>(ring + (t % kWRing) * tile_floats);
      for (int i = threadIdx.x; i < rows * N / 4; i += kHeadThreads) cp_async16_h(dst + i, src + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int pass = 0; pass < items; pass += kHeadThreads) {
    const int it = pass + threadIdx.x;
    const bool live = it < items;
    const int n = live ? (it % NV) * 2 : 0, ag = live ? it / NV : 0;
    float acc0[kAT], acc1[kAT];
#pragma unroll
    for (int a = 0; a < kAT; ++a) {
      acc0[a] = acc_init(bias, init, N, n, ag * kAT + a, na);
      acc1[a] = acc_init(bias, init, N, n + 1, ag * kAT + a, na);
    }
    const float* src = in + ag * kAT;
    issue(0);
    issue(1);
    for (int t = 0; t < ntiles; ++t) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");     // tile t has landed (tile t+1 may still be in flight)
      __syncthreads();                                          // ... for every thread; tile t-1's buffer is free again
      issue(t + 2);
      const float* wt = ring + (t % kWRing) * tile_floats + n;
      const int rows = min(kKT, K - t * kKT);
      const float* sk = src + (size_t)t * kKT * kTAp;
      if (live) {
#pragma unroll 8
        for (int k = 0; k < rows; ++k) {
          const float2 w = *reinterpret_cast<const float2*>(wt + k * N);
          const float4 v0 = *reinterpret_cast<const float4*>(sk + k * kTAp);
          const float4 v1 = *reinterpret_cast<const float4*>(sk + k * kTAp + 4);
          acc0[0] = fmaf(w.x, v0.x, acc0[0]); acc1[0] = fmaf(w.y, v0.x, acc1[0]);
          acc0[1] = fmaf(w.x, v0.y, acc0[1]); acc1[1] = fmaf(w.y, v0.y, acc1[1]);
          acc0[2] = fmaf(w.x, v0.z, acc0[2]); acc1[2] = fmaf(w.y, v0.z, acc1[2]);
          acc0[3] = fmaf(w.x, v0.w, acc0[3]); acc1[3] = fmaf(w.y, v0.w, acc1[3]);
          acc0[4] = fmaf(w.x, v1.x, acc0[4]); acc1[4] = fmaf(w.y, v1.x, acc1[4]);
          acc0[5] = fmaf(w.x, v1.y, acc0[5]); acc1[5] = fmaf(w.y, v1.y, acc1[5]);
          acc0[6] = fmaf(w.x, v1.z, acc0[6]); acc1[6] = fmaf(w.y, v1.z, acc1[6]);
          acc0[7] = fmaf(w.x, v1.w, acc0[7]); acc1[7] = fmaf(w.y, v1.w, acc1[7]);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                            // ring reusable by the next pass / layer
    if (live) {
      float* d0 = out + n * kTAp + ag * kAT;
#pragma unroll
      for (int a = 0; a < kAT; ++a) { d0[a] = apply_act(acc0[a], act); d0[kTAp + a] = apply_act(acc1[a], act); }
    }
  }
}

__device__ __forceinline__ void dense_layer(const float* __restrict__ in, int K, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, int N, int act, float* __restrict__ out,
                                            float* __restrict__ ring, const float* __restrict__ init = nullptr,
                                            int na = kTA) {
  if (K == 0) {                                     // the whole contraction was pre-computed: out = act(init)
    for (int i = threadIdx.x; i < N * kTA; i += kHeadThreads) {
      const int n = i / kTA, a = i - n * kTA;
      out[n * kTAp + a] = apply_act(acc_init(bias, init, N, n, a, na), act);
    }
    return;
  }
  if ((N & 3) == 0 && N <= kWTileMaxN && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0)
    dense_layer_ring(in, K, Wt, bias, N, act, out, ring, init, na);
  else if ((N & 1) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 7) == 0)
    dense_layer_t<2>(in, K, Wt, bias, N, act, out, init, na);
  else dense_layer_t<1>(in, K, Wt, bias, N, act, out, init, na);
}

__global__ void __launch_bounds__(kHeadThreads)
meta_head_kernel(btsb_head_params p, int64_t B, float* __restrict__ logits) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * kTA;
  const int na = (int)min((int64_t)kTA, B - b0);
  const int F = p.F, Mm = p.Mm, m1 = p.m1, m2 = p.m2, c1 = p.c1, c2 = p.c2;
  const int emb = (Mm > 0) ? m2 : 0;
  // h0_init: the caller already contracted the F image features of the head's first layer (tensor-core GEMM, bias
  // included); the feature rows are then neither staged nor multiplied here
  const bool pre = p.h0_init != nullptr;
  const int Fs = pre ? 0 : F;                        // feature rows staged in `cat`
  float* cat = sm;                                   // [Fs+emb][kTAp]
  float* bufM = cat + (size_t)(Fs + emb) * kTAp;     // [Mm][kTAp]
  float* buf1 = bufM + (size_t)Mm * kTAp;            // [m1][kTAp]
  float* bufH0 = buf1 + (size_t)m1 * kTAp;           // [c1][kTAp]
  float* bufH1 = bufH0 + (size_t)c1 * kTAp;          // [c2][kTAp]
  float* ring = bufH1 + (size_t)c2 * kTAp;           // [kWRing][kKT][<=128] weight tiles (16-byte aligned: kTAp % 4 == 0)

  // stage inputs (zero-fill alerts beyond the batch tail)
  if (Fs > 0) {
    // a warp reads 32 consecutive features of one alert (coalesced) and scatters them down a column of `cat`
    // (4-way bank conflict on the store: 640 warp-stores per CTA, negligible next to the 768-deep contraction)
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
    dense_layer(bufM, Mm, p.m1t, p.m1b, m1, p.meta_act, buf1, ring);
    __syncthreads();
    dense_layer(buf1, m1, p.m2t, p.m2b, m2, p.meta_out_act, cat + (size_t)Fs * kTAp, ring);
    __syncthreads();
  }
  const float* last = cat;
  int lastK = Fs + emb;
  if (c1 > 0) {
    dense_layer(cat, Fs + emb, p.h0t + (size_t)(F - Fs) * c1, p.h0b, c1, p.head_act, bufH0, ring,
                pre ? p.h0_init + (size_t)b0 * c1 : nullptr, na);
    __syncthreads();
    dense_layer(bufH0, c1, p.h1t, p.h1b, c2, p.head_act, bufH1, ring);
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
  BTSB_REQUIRE(p.F == 0 || p.h0_init || (p.feat && (p.feat_dtype == BTSB_F32 || p.feat_dtype == BTSB_BF16)),
               "meta_head: bad feat");
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
  if (p.h0_init) BTSB_REQUIRE(p.c1 > 0 && p.F > 0, "meta_head: h0_init needs a head (c1 > 0) and F > 0 (the h0t row offset)");
  const int Fs = p.h0_init ? 0 : p.F;
  const size_t smem = (size_t)(Fs + emb + p.Mm + (p.Mm > 0 ? p.m1 : 0) + p.c1 + p.c2) * kTAp * sizeof(float) +
                      (size_t)kWRing * kKT * kWTileMaxN * sizeof(float);
  BTSB_REQUIRE(smem <= 227 * 1024, "meta_head: layer widths need %zu B of shared memory (> 227 KB)", smem);
  BTSB_CUDA(cudaFuncSetAttribute(meta_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "meta_head attr");
  const int64_t grid = (B + kTA - 1) / kTA;
  meta_head_kernel<<<(unsigned)grid, kHeadThreads, smem, (cudaStream_t)stream>>>(p, B, logits);
  return launch_done("meta_head");
}
