// K3 v3 (bf16 or fp32 activations, square maps 15/7/3/1): depthwise 7x7 + bias + LayerNorm2d with the LayerNorm statistics
// reduced straight from the convolution registers -- no fp32 staging buffer, no second pass over the map.
//
//   * persistent CTA, one image per iteration, next image prefetched with cp.async
//   * thread = (output row, channel pair); the S outputs of the row live in registers (compile-time tap pruning)
//   * channel pairs of a row are spread over A full warps (32 pairs each) plus, when C/2 is not a multiple of 32,
//     aligned segments of REM = 8 or 16 lanes inside "tail" warps, so every cross-channel reduction is an aligned
//     power-of-two shuffle tree
//   * per-pixel sum / sum-of-squares: recursive-halving shuffle reduction (31 shuffles for 30 values instead of 150),
//     partial results of the A+1 contributors meet in a 4 KB shared-memory table; each thread then normalises its own
//     registers and writes bf16x2 (coalesced 128 B per warp)
//   * T = float (the 1e-4 mode): the same structure on fp32 rows (8-byte loads / stores of a channel pair), with a TWO-pass
//     variance (a second shuffle reduction over the squared deviations) as the generic fp32 kernel and torch compute it;
//     replaces the generic thread-per-channel kernel there (2.0 ms per 8192 images at 15 x 15 x 80)
#include "common.cuh"

namespace btsb {

constexpr int kDw3MaxThreads = 608;

__device__ __forceinline__ void cp_async16_v3(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// CT > 0: channel count known at compile time (nano 80/160, pico 64/128) -- every shared-memory address becomes an
// immediate offset, which removes ~1/3 of the issued instructions (integer address arithmetic; profiles/r01c);
// CT == 0: generic runtime C.
template <int S, int CT, typename T = __nv_bfloat16, bool XF16 = false>   // XF16: 2-byte input rows are IEEE fp16 (out: bf16)
__global__ void __launch_bounds__(kDw3MaxThreads, 1)
dwln3_kernel(const T* __restrict__ x, int64_t B, int C_rt, int A_rt, int REM_rt, const float* __restrict__ wt,
             const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
             T* __restrict__ out) {
  constexpr bool F32 = sizeof(T) == 4;
  constexpr int EPV = 16 / (int)sizeof(T);                   // elements per 16-byte cp.async piece
  constexpr int R = S > 3 ? 3 : S - 1;
  constexpr int NT = 2 * R + 1;
  constexpr int HW = S * S;
  static_assert(2 * S <= 32, "row statistics must fit the 32-value reduction");
  const int C = CT > 0 ? CT : C_rt;
  const int A = CT > 0 ? (CT / 2) / 32 : A_rt;
  const int REM = CT > 0 ? (CT / 2) % 32 : REM_rt;
  extern __shared__ __align__(16) unsigned char sm[];
  float* wsm = reinterpret_cast<float*>(sm);                 // [NT*NT][C] reachable taps
  float* bsm = wsm + NT * NT * C;                            // conv bias
  float* gsm = bsm + C;                                      // LN weight
  float* hsm = gsm + C;                                      // LN bias
  const int ncontrib = A + (REM > 0 ? 1 : 0);
  float* part = hsm + C;                                     // [S rows][32 values][ncontrib] (+ a second table for the fp32 two-pass)
  T* tin = reinterpret_cast<T*>(part + (F32 ? 2 : 1) * S * 32 * ncontrib);           // [2][HW*C]

  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int img_elems = HW * C;
  const int C2 = C >> 1;

  // ---- thread -> (row, channel pair, contributor slot) -------------------------------------------------------------
  const int main_warps = S * A;
  int row, c2, contrib, width;
  bool active = true;
  if (warp < main_warps) {
    row = warp / A; contrib = warp - row * A; c2 = contrib * 32 + lane; width = 32;
  } else {
    const int per = 32 / REM;                                // rows per tail warp (REM > 0 here)
    row = (warp - main_warps) * per + lane / REM;
    c2 = A * 32 + lane % REM; contrib = A; width = REM;
    active = row < S;
    if (!active) row = S - 1;                                // keep shuffles convergent; results discarded
  }

  auto issue = [&](int64_t img, int buf) {
    const uint4* src = reinterpret_cast<const uint4*>(x + img * img_elems);
    uint4* dst = reinterpret_cast<uint4*>(tin + (size_t)buf * img_elems);
    for (int i = tid; i < img_elems / EPV; i += nthr) cp_async16_v3(dst + i, src + i);
  };
  if ((int64_t)blockIdx.x < B) issue(blockIdx.x, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = tid; i < NT * NT * C; i += nthr) {
    const int t = i / C, c = i - t * C;
    const int ty = t / NT, tx = t - ty * NT;
    wsm[i] = __ldg(wt + ((ty + 3 - R) * 7 + (tx + 3 - R)) * C + c);   // fp16 input: common.cuh
  }
  for (int i = tid; i < C; i += nthr) { bsm[i] = __ldg(bias + i); gsm[i] = __ldg(ln_w + i); hsm[i] = __ldg(ln_b + i); }

  const float invC = 1.0f / (float)C;
  int it = 0;
  for (int64_t img = blockIdx.x; img < B; img += gridDim.x, ++it) {
    const int buf = it & 1;
    const int64_t nxt = img + gridDim.x;
    if (nxt < B) issue(nxt, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                          // input landed; previous iteration's `part` reads are done

    // ---- depthwise conv of one output row, two channels ------------------------------------------------------------
    float acc0[S], acc1[S];
    {
      f32x2_t acc[S];
      const f32x2_t bv = *reinterpret_cast<const f32x2_t*>(bsm + 2 * c2);
#pragma unroll
      for (int t = 0; t < S; ++t) acc[t] = bv;
      const T* im = tin + (size_t)buf * img_elems + 2 * c2;
#pragma unroll
      for (int dy = -R; dy <= R; ++dy) {
        const int iy = row + dy;
        if (iy < 0 || iy >= S) continue;
        f32x2_t wv[NT];
#pragma unroll
        for (int kx = 0; kx < NT; ++kx) wv[kx] = *reinterpret_cast<const f32x2_t*>(wsm + ((dy + R) * NT + kx) * C + 2 * c2);
#pragma unroll
        for (int ix = 0; ix < S; ++ix) {
          f32x2_t xin;
          if constexpr (F32) xin = *reinterpret_cast<const f32x2_t*>(im + (size_t)(iy * S + ix) * C);
          else xin = x2_to_f32x2<XF16>(*reinterpret_cast<const uint32_t*>(im + (size_t)(iy * S + ix) * C));
#pragma unroll
          for (int kx = 0; kx < NT; ++kx) {
            const int t = ix - (kx - R);
            if (t >= 0 && t < S) fma_f32x2(acc[t], wv[kx], xin);   // both channels of the pair in one FFMA2
          }
        }
      }
#pragma unroll
      for (int t = 0; t < S; ++t) { const float2 a = unpack_f32x2(acc[t]); acc0[t] = a.x; acc1[t] = a.y; }
    }

    if constexpr (F32) {
      // ---- fp32 mode: two-pass statistics.  Pass 1: per-pixel channel sums (values [0,S)); pass 2: sums of squared
      // deviations from the mean (second table), each a recursive-halving shuffle reduction + one shared-memory hop.
      float* part2 = part + S * 32 * ncontrib;
      auto reduce_to = [&](float (&red)[32], float* tab) {
        if (width == 32) {
          seg_reduce32<32>(red, lane);
          if (lane < S) tab[(row * 32 + lane) * ncontrib + contrib] = red[0];
        } else if (width == 16) {
          seg_reduce32<16>(red, lane);
          const int base = (lane & 15) * 2;
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (active && base + i < S) tab[(row * 32 + base + i) * ncontrib + contrib] = red[i];
        } else {
          seg_reduce32<8>(red, lane);
          const int base = (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (active && base + i < S) tab[(row * 32 + base + i) * ncontrib + contrib] = red[i];
        }
      };
      float red[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
      for (int t = 0; t < S; ++t) red[t] = acc0[t] + acc1[t];
      reduce_to(red, part);
      __syncthreads();
      float mean[S];
#pragma unroll
      for (int t = 0; t < S; ++t) {
        float s = 0.f;
        for (int k = 0; k < ncontrib; ++k) s += part[(row * 32 + t) * ncontrib + k];
        mean[t] = s * invC;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
      for (int t = 0; t < S; ++t) {
        const float d0 = acc0[t] - mean[t], d1 = acc1[t] - mean[t];
        red[t] = fmaf(d0, d0, d1 * d1);
      }
      reduce_to(red, part2);
      __syncthreads();
      if (active) {
        const float2 gw = *reinterpret_cast<const float2*>(gsm + 2 * c2);
        const float2 gb = *reinterpret_cast<const float2*>(hsm + 2 * c2);
        float2* dst = reinterpret_cast<float2*>(out + (img * HW + row * S) * (int64_t)C) + c2;
#pragma unroll
        for (int t = 0; t < S; ++t) {
          float q = 0.f;
          for (int k = 0; k < ncontrib; ++k) q += part2[(row * 32 + t) * ncontrib + k];
          const float rstd = rsqrtf(q * invC + kLnEps);
          dst[(size_t)t * C2] = make_float2((acc0[t] - mean[t]) * rstd * gw.x + gb.x, (acc1[t] - mean[t]) * rstd * gw.y + gb.y);
        }
      }
    } else {
    // ---- per-pixel channel statistics: values [0,S) = sums, [16,16+S) = sums of squares -----------------------------
    float red[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      red[t] = acc0[t] + acc1[t];
      red[16 + t] = fmaf(acc0[t], acc0[t], acc1[t] * acc1[t]);
    }
    if (width == 32) {
      seg_reduce32<32>(red, lane);
      if (lane < S || (lane >= 16 && lane < 16 + S)) part[(row * 32 + lane) * ncontrib + contrib] = red[0];
    } else if (width == 16) {
      seg_reduce32<16>(red, lane);
      const int base = (lane & 15) * 2;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = base + i;
        if (active && (idx < S || (idx >= 16 && idx < 16 + S))) part[(row * 32 + idx) * ncontrib + contrib] = red[i];
      }
    } else {
      seg_reduce32<8>(red, lane);
      const int base = (lane & 7) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = base + i;
        if (active && (idx < S || (idx >= 16 && idx < 16 + S))) part[(row * 32 + idx) * ncontrib + contrib] = red[i];
      }
    }
    __syncthreads();

    // ---- normalise the registers and store bf16x2 -------------------------------------------------------------------
    if (active) {
      const float2 gw = *reinterpret_cast<const float2*>(gsm + 2 * c2);
      const float2 gb = *reinterpret_cast<const float2*>(hsm + 2 * c2);
      uint32_t* dst = reinterpret_cast<uint32_t*>(out + (img * HW + row * S) * (int64_t)C) + c2;
#pragma unroll
      for (int t = 0; t < S; ++t) {
        float s = 0.f, q = 0.f;
        for (int k = 0; k < ncontrib; ++k) {
          s += part[(row * 32 + t) * ncontrib + k];
          q += part[(row * 32 + 16 + t) * ncontrib + k];
        }
        const float mean = s * invC;
        const float var = fmaxf(q * invC - mean * mean, 0.f);
        const float rstd = rsqrtf(var + kLnEps);
        __nv_bfloat162 o = __floats2bfloat162_rn((acc0[t] - mean) * rstd * gw.x + gb.x, (acc1[t] - mean) * rstd * gw.y + gb.y);
        dst[(size_t)t * C2] = *reinterpret_cast<uint32_t*>(&o);
      }
    }
    }
    // the next iteration's first __syncthreads orders these `part` reads before the next writes
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

int num_sms();

template <int S, int CT, typename T = __nv_bfloat16, bool XF16 = false>
static int launch_dwln3(const void* x, int64_t B, int C, const float* w, const float* bias, const float* ln_w,
                        const float* ln_b, void* out, cudaStream_t st) {
  constexpr int R = S > 3 ? 3 : S - 1;
  constexpr int NT = 2 * R + 1;
  constexpr int HW = S * S;
  const int C2 = C / 2, A = C2 / 32, REM = C2 % 32;
  if (!(REM == 0 || REM == 8 || REM == 16)) return 1;
  const int tail_warps = REM ? (S * REM + 31) / 32 : 0;
  const int warps = S * A + tail_warps;
  if (warps * 32 > kDw3MaxThreads || warps < 1) return 1;
  const int ncontrib = A + (REM ? 1 : 0);
  const size_t smem = (size_t)(NT * NT + 3) * C * 4 + (size_t)(sizeof(T) == 4 ? 2 : 1) * S * 32 * ncontrib * 4 +
                      2 * (size_t)HW * C * sizeof(T);
  if (smem > 227 * 1024) return 1;
  auto kern = dwln3_kernel<S, CT, T, XF16>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "dwln3 attr");
  const int sms = num_sms();
  // small maps: several CTAs per SM are possible (few warps, little shared memory)
  int per_sm = 1;
  if (warps * 32 <= 304 && smem <= 100 * 1024) per_sm = 2;
  const int64_t cap = (int64_t)sms * per_sm;
  const int grid = (int)(B < cap ? B : cap);
  kern<<<grid, warps * 32, smem, st>>>((const T*)x, B, C, A, REM, w, bias, ln_w, ln_b, (T*)out);
  return launch_done("dwln3");
}

// returns 1 if the shape is not handled here (caller falls back to dwln2 / the generic kernel)
int dwln_bf16_v3(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, bool xf16, cudaStream_t st) {
  if (H != W || C % 16 != 0 || C > 640) return 1;
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 4) != 0) return 1;
  using BF = __nv_bfloat16;
  if (xf16) {                       // fp16 residual stream: the widths that reach this kernel in the nano / pico trunks
    if (H == 15 && C == 64) return launch_dwln3<15, 64, BF, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    if (H == 7 && C == 128) return launch_dwln3<7, 128, BF, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    if (H == 15) return launch_dwln3<15, 0, BF, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    if (H == 7) return launch_dwln3<7, 0, BF, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    return 1;
  }
  switch (H) {
    case 15:
      if (C == 80) return launch_dwln3<15, 80>(x, B, C, w, bias, ln_w, ln_b, out, st);
      if (C == 64) return launch_dwln3<15, 64>(x, B, C, w, bias, ln_w, ln_b, out, st);
      return launch_dwln3<15, 0>(x, B, C, w, bias, ln_w, ln_b, out, st);
    case 7:
      if (C == 160) return launch_dwln3<7, 160>(x, B, C, w, bias, ln_w, ln_b, out, st);
      if (C == 128) return launch_dwln3<7, 128>(x, B, C, w, bias, ln_w, ln_b, out, st);
      return launch_dwln3<7, 0>(x, B, C, w, bias, ln_w, ln_b, out, st);
    // 3x3 and 1x1 maps: one image per CTA iteration is too little work per barrier; dwln2 (several images per
    // iteration) is faster there (measured: 0.078 vs 0.106 ms and 0.020 vs 0.050 ms at B = 8192)
    default: return 1;
  }
}

// fp32 rows (the 1e-4 mode): 15 x 15 and 7 x 7 maps at the nano / pico widths; returns 1 otherwise (generic kernel)
int dwln_f32_v3(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                const float* ln_b, void* out, cudaStream_t st) {
  if (H != W || ((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 8) != 0) return 1;
  if (H == 15 && C == 80) return launch_dwln3<15, 80, float>(x, B, C, w, bias, ln_w, ln_b, out, st);
  if (H == 15 && C == 64) return launch_dwln3<15, 64, float>(x, B, C, w, bias, ln_w, ln_b, out, st);
  if (H == 7 && C == 160) return launch_dwln3<7, 160, float>(x, B, C, w, bias, ln_w, ln_b, out, st);
  if (H == 7 && C == 128) return launch_dwln3<7, 128, float>(x, B, C, w, bias, ln_w, ln_b, out, st);
  return 1;
}

}  // namespace btsb
