// K3 v5 (bf16 activations, 15x15 and 7x7 maps, compile-time C): depthwise 7x7 + bias + LayerNorm2d with the two halves
// of the work on DIFFERENT warps so that they overlap in time.
//
// ncu on v3 (dwln3.cu, profiles/r01m/dwln15.*): during the convolution the FP32 pipe is the limiter (top stall "math
// pipe throttle" on the FFMA2s), but over the whole kernel it is busy only 55 % of the time -- the LayerNorm statistics
// (shuffle tree), the normalisation and two CTA barriers per image run on the SAME threads after the convolution, with the
// FMA pipe idle.  Here
//   conv warps (thread = output row x channel pair, exactly v3's mapping and inner loop) convolve image i+1 while
//   4 LayerNorm warps normalise image i from a shared-memory tile T[HW][C] of fp32 conv results (TPP threads per pixel on
//   contiguous channel slices, thread-local sum / sum of squares, log2(TPP) shuffles, LN weights in registers) and write
//   the bf16 rows to global.
// Hand-off through two mbarriers (tile full: one arrival per conv warp; tile empty: one per LayerNorm warp); the conv
// warps synchronise among themselves with a named barrier only.
#include "common.cuh"

namespace btsb {
namespace {

__device__ __forceinline__ void cp_async16_v5(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if ((++spins & 1023u) == 0) {            // bounded: a protocol bug traps instead of hanging the GPU
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) { printf("btsbot_b200: dwln5 barrier wait timed out\n"); __trap(); }
    }
  }
}

template <int S, int C>
struct Dw5 {
  static constexpr int HW = S * S;
  static constexpr int C2 = C / 2;
  static constexpr int A = C2 / 32;
  static constexpr int REM = C2 % 32;
  static constexpr int PER = REM ? 32 / REM : 1;
  static constexpr int TAIL_WARPS = REM ? (S + PER - 1) / PER : 0;
  static constexpr int CONV_WARPS = S * A + TAIL_WARPS;
  static constexpr int CONV_THREADS = CONV_WARPS * 32;
  static constexpr int LN_WARPS = 4;
  static constexpr int LN_THREADS = LN_WARPS * 32;
  static constexpr int THREADS = CONV_THREADS + LN_THREADS;
  static constexpr int TPP = C >= 128 ? 8 : 4;
  static constexpr int CPT = C / TPP;
  static constexpr size_t SMEM = (size_t)(49 + 1) * C * 4 + (size_t)HW * C * 4 + 2 * (size_t)HW * C * 2 + 64;
  static_assert(REM == 0 || REM == 8 || REM == 16, "channel pairs per row must split into aligned lane segments");
  static_assert(CPT % 4 == 0 && C % TPP == 0 && LN_THREADS % TPP == 0, "LayerNorm slices are whole float4s");
  static_assert(THREADS <= 1024 && (HW * C) % 8 == 0, "shape");
};

template <int S, int C, bool XF16>
__global__ void __launch_bounds__(Dw5<S, C>::THREADS, 1)
dwln5_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, const float* __restrict__ wt, const float* __restrict__ bias,
             const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  using P = Dw5<S, C>;
  constexpr int HW = P::HW, TPP = P::TPP, CPT = P::CPT;
  extern __shared__ __align__(16) unsigned char sm[];
  float* wsm = reinterpret_cast<float*>(sm);                 // [49][C]
  float* bsm = wsm + 49 * C;                                 // conv bias
  float* tile = bsm + C;                                     // [HW][C] fp32 conv results of one image
  __nv_bfloat16* tin = reinterpret_cast<__nv_bfloat16*>(tile + HW * C);   // [2][HW][C] bf16 input
  const uint32_t bar_full = s_u32(tin + 2 * HW * C), bar_empty = bar_full + 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < 49 * C; i += P::THREADS) wsm[i] = __ldg(wt + i);   // fp16 input: see common.cuh
  for (int i = tid; i < C; i += P::THREADS) bsm[i] = __ldg(bias + i);
  if (tid == 0) {
    mb_init(bar_full, P::CONV_WARPS);
    mb_init(bar_empty, P::LN_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp < P::CONV_WARPS) {
    // ============================ convolution warps ============================
    int row, c2;
    bool active = true;
    if (warp < S * P::A) {
      row = warp / (P::A > 0 ? P::A : 1); c2 = (warp - row * P::A) * 32 + lane;
    } else {
      row = (warp - S * P::A) * P::PER + lane / (P::REM > 0 ? P::REM : 32);
      c2 = P::A * 32 + lane % (P::REM > 0 ? P::REM : 32);
      active = row < S;
      if (!active) row = S - 1;
    }
    auto issue = [&](int64_t img, int buf) {
      const uint4* src = reinterpret_cast<const uint4*>(x + img * (int64_t)(HW * C));
      uint4* dst = reinterpret_cast<uint4*>(tin + (size_t)buf * (HW * C));
      for (int i = tid; i < HW * C / 8; i += P::CONV_THREADS) cp_async16_v5(dst + i, src + i);
    };
    if ((int64_t)blockIdx.x < B) issue(blockIdx.x, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int it = 0;
    for (int64_t img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      const int buf = it & 1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      // image `it` has landed for every conv thread, and every conv thread has finished reading buffer buf^1
      asm volatile("bar.sync 1, %0;" ::"n"(P::CONV_THREADS) : "memory");
      const int64_t nxt = img + gridDim.x;
      if (nxt < B) issue(nxt, buf ^ 1);                       // lands during this image's convolution
      asm volatile("cp.async.commit_group;" ::: "memory");

      f32x2_t acc[S];
      {
        const f32x2_t bv = *reinterpret_cast<const f32x2_t*>(bsm + 2 * c2);
#pragma unroll
        for (int t = 0; t < S; ++t) acc[t] = bv;
        const __nv_bfloat16* im = tin + (size_t)buf * (HW * C) + 2 * c2;
#pragma unroll
        for (int dy = -3; dy <= 3; ++dy) {
          const int iy = row + dy;
          if (iy < 0 || iy >= S) continue;
          f32x2_t wv[7];
#pragma unroll
          for (int kx = 0; kx < 7; ++kx) wv[kx] = *reinterpret_cast<const f32x2_t*>(wsm + ((dy + 3) * 7 + kx) * C + 2 * c2);
#pragma unroll
          for (int ix = 0; ix < S; ++ix) {
            const f32x2_t xin = x2_to_f32x2<XF16>(*reinterpret_cast<const uint32_t*>(im + (iy * S + ix) * C));
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
              const int t = ix - (kx - 3);
              if (t >= 0 && t < S) fma_f32x2(acc[t], wv[kx], xin);
            }
          }
        }
      }
      if (it > 0) mb_wait(bar_empty, (uint32_t)((it - 1) & 1));   // the LayerNorm warps are done with the previous image
      if (active) {
        float* dst = tile + (row * S) * C + 2 * c2;
#pragma unroll
        for (int t = 0; t < S; ++t) *reinterpret_cast<f32x2_t*>(dst + t * C) = acc[t];
      }
      __syncwarp();
      if (lane == 0) mb_arrive(bar_full);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // ============================ LayerNorm warps ============================
    const int lt = tid - P::CONV_THREADS;                    // 0 .. LN_THREADS-1
    const int s = lt % TPP;                                  // this thread's channel slice (fixed: LN_THREADS % TPP == 0)
    float gw[CPT], gb[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { gw[j] = __ldg(ln_w + s * CPT + j); gb[j] = __ldg(ln_b + s * CPT + j); }
    constexpr float invC = 1.0f / (float)C;
    constexpr int kTasks = HW * TPP;
    constexpr int kRounds = (kTasks + P::LN_THREADS - 1) / P::LN_THREADS;
    int it = 0;
    for (int64_t img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      mb_wait(bar_full, (uint32_t)(it & 1));
#pragma unroll 1
      for (int r = 0; r < kRounds; ++r) {
        const int k = r * P::LN_THREADS + lt;
        const bool live = k < kTasks;
        const int p = live ? k / TPP : 0;
        const float4* src = reinterpret_cast<const float4*>(tile + p * C + s * CPT);
        float4 v[CPT / 4];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          v[j] = src[j];
          sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
        if (r == kRounds - 1) {                               // all of this warp's reads of the tile have been issued and
          __syncwarp();                                       // consumed (v[] is in registers): the tile may be overwritten
          if (lane == 0) mb_arrive(bar_empty);
        }
#pragma unroll
        for (int o = TPP / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * invC;
        // two-pass variance (the row slice is in registers): no E[x^2] - mean^2 cancellation, as torch computes it
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
          sq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq))));
        }
#pragma unroll
        for (int o = TPP / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * invC + kLnEps);
        if (live) {
          uint2* dst = reinterpret_cast<uint2*>(out + (img * HW + p) * (int64_t)C + s * CPT);
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) {
            const __nv_bfloat162 o0 = __floats2bfloat162_rn((v[j].x - mean) * rstd * gw[4 * j] + gb[4 * j],
                                                            (v[j].y - mean) * rstd * gw[4 * j + 1] + gb[4 * j + 1]);
            const __nv_bfloat162 o1 = __floats2bfloat162_rn((v[j].z - mean) * rstd * gw[4 * j + 2] + gb[4 * j + 2],
                                                            (v[j].w - mean) * rstd * gw[4 * j + 3] + gb[4 * j + 3]);
            dst[j] = make_uint2(*reinterpret_cast<const uint32_t*>(&o0), *reinterpret_cast<const uint32_t*>(&o1));
          }
        }
      }
    }
  }
}

}  // namespace

int num_sms();

template <int S, int C, bool XF16>
static int launch_dwln5(const void* x, int64_t B, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                        void* out, cudaStream_t st) {
  using P = Dw5<S, C>;
  auto kern = dwln5_kernel<S, C, XF16>;
  static std::atomic<uint64_t> attr_done{0};
  if (first_use_on_device(attr_done)) {
    BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM), "dwln5 attr");
  }
  const int64_t cap = num_sms();
  const int grid = (int)(B < cap ? B : cap);
  kern<<<grid, P::THREADS, P::SMEM, st>>>((const __nv_bfloat16*)x, B, w, bias, ln_w, ln_b, (__nv_bfloat16*)out);
  return launch_done("dwln5");
}

// returns 1 if the shape is not handled here (caller falls back to v3 / v2 / the generic kernel)
int dwln_bf16_v5(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, bool xf16, cudaStream_t st) {
  if (H != W) return 1;
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 16) != 0) return 1;
  // nano (80 / 160) and pico (64 / 128) widths.  Round 1 kept the pico widths on v3 because this kernel's different fp32
  // summation order of the LayerNorm statistics moved the synthetic frozen-fusion / pico case across the 2e-2 bar
  // (1.34e-2 -> 2.35e-2): that case sat inside the bf16 noise band, which the fp16 residual stream has since halved
  // (DESIGN.md lesson 20), so kernels are no longer chosen by which rounding pattern the goldens happen to pass with.
  if (xf16) {
    if (H == 15 && C == 80) return launch_dwln5<15, 80, true>(x, B, w, bias, ln_w, ln_b, out, st);
    if (H == 7 && C == 160) return launch_dwln5<7, 160, true>(x, B, w, bias, ln_w, ln_b, out, st);
    if (H == 15 && C == 64) return launch_dwln5<15, 64, true>(x, B, w, bias, ln_w, ln_b, out, st);
    if (H == 7 && C == 128) return launch_dwln5<7, 128, true>(x, B, w, bias, ln_w, ln_b, out, st);
    return 1;
  }
  if (H == 15 && C == 80) return launch_dwln5<15, 80, false>(x, B, w, bias, ln_w, ln_b, out, st);
  if (H == 7 && C == 160) return launch_dwln5<7, 160, false>(x, B, w, bias, ln_w, ln_b, out, st);
  if (H == 15 && C == 64) return launch_dwln5<15, 64, false>(x, B, w, bias, ln_w, ln_b, out, st);
  if (H == 7 && C == 128) return launch_dwln5<7, 128, false>(x, B, w, bias, ln_w, ln_b, out, st);
  return 1;
}

}  // namespace btsb
