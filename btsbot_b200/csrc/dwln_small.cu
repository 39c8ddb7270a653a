// K3 for the small maps (bf16 activations, 3x3 and 1x1): depthwise 7x7 + bias + LayerNorm2d, one image per CTA
// iteration, everything in registers.
//
// On a 3x3 map every output pixel sees every input pixel (81 (out,in) pairs, 25 distinct reachable taps), so a thread
// that owns a CHANNEL PAIR loads the image's 9 bf16x2 values straight from global memory (coalesced: a warp reads 128
// contiguous bytes per pixel), runs the 162 FMAs, and the LayerNorm statistics of the 9 pixels are reduced over the
// channel pairs with the 31-shuffle recursive-halving tree (common.cuh) + one shared-memory hop across the warps.
// The previous path (dwln2: fp32 staging in shared memory + a second warp-per-pixel pass) ran at 1.2 TB/s on this
// shape; the work is 94 MB of traffic and 0.2 GFMA per 8192 alerts, i.e. this kernel should sit on the HBM roofline.
#include <stdlib.h>

#include "common.cuh"

namespace btsb {

// CT > 0: channel count known at compile time (nano 320/640, pico 256/512): the 9 loads, 9 stores and 25 tap reads become
// immediate offsets -- the kernel was instruction-issue-bound with ~240 of its ~800 instructions per image being 64-bit
// address arithmetic (3 CTAs of 120 registers per SM leave no latency slack either).  CT == 0: generic runtime C.
// PF: prefetch distance in images.  One image per thread is 9 x 4 B; with 5-6 resident CTAs of C/2 threads an SM has
// ~35 KB of loads in flight at PF = 1, the bare Little's-law minimum for its share of the HBM bandwidth (the kernel ran
// at 2.5 of ~6.5 TB/s); PF = 2 keeps two images per thread in flight for 9 more registers.
// CPA (opt-in, UNMEASURED, CT > 0 only): the reachable taps -- NT contiguous runs of NT*C floats in the [49][C] tap matrix --
// are staged with one wave of 16-byte cp.async instead of ~NT*NT*C/T dependent batches of scalar loads per thread.
template <int S, int CT, int PF, bool CPA = false, bool XF16 = false>   // XF16: input rows are IEEE fp16 (out: bf16)
__global__ void dwln_small_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int C_rt, const float* __restrict__ wt,
                                  const float* __restrict__ bias, const float* __restrict__ ln_w,
                                  const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  constexpr int R = S - 1;
  constexpr int NT = 2 * R + 1;
  constexpr int HW = S * S;
  static_assert(HW <= 16, "statistics layout: sums in [0,16), sums of squares in [16,32)");
  const int C = CT > 0 ? CT : C_rt;
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  float* wsm = reinterpret_cast<float*>(sm);                 // [NT*NT][C] reachable taps
  float* part = wsm + NT * NT * C;                           // [2][nw][32]
  float2* stat = reinterpret_cast<float2*>(part + 2 * nw * 32);   // [2][16] (mean, rstd)
  const int C2 = C >> 1;
  const bool active = tid < C2;
  const int c2 = active ? tid : 0;

  if constexpr (CPA && CT > 0) {
    static_assert(!CPA || CT % 4 == 0, "16-byte granules");
    constexpr int kRun = NT * (CT > 0 ? CT : 4) / 4;         // 16-byte granules per tap row
    for (int i = tid; i < NT * kRun; i += T) {
      const int ty = i / kRun, k = i - ty * kRun;
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(wsm + ty * NT * C + 4 * k);
      const float* g = wt + ((ty + 3 - R) * 7 + (3 - R)) * C + 4 * k;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    for (int i = tid; i < NT * NT * C; i += T) {
      const int t = i / C, c = i - t * C;
      const int ty = t / NT, tx = t - ty * NT;
      wsm[i] = __ldg(wt + ((ty + 3 - R) * 7 + (tx + 3 - R)) * C + c);
    }
  }
  const float2 bv = *reinterpret_cast<const float2*>(bias + 2 * c2);
  const float2 gw = *reinterpret_cast<const float2*>(ln_w + 2 * c2);
  const float2 gb = *reinterpret_cast<const float2*>(ln_b + 2 * c2);
  __syncthreads();

  uint32_t cur[HW], nxt[HW], nx2[PF == 2 ? HW : 1];
  auto load_img = [&](int64_t img, uint32_t (&v)[HW]) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + img * (int64_t)HW * C) + c2;
#pragma unroll
    for (int p = 0; p < HW; ++p) v[p] = (active && img < B) ? __ldg(src + (size_t)p * C2) : 0u;
  };
  load_img(blockIdx.x, cur);
  if (PF == 2) load_img((int64_t)blockIdx.x + gridDim.x, nxt);
  const float invC = 1.0f / (float)C;
  int it = 0;
  for (int64_t img = blockIdx.x; img < B; img += gridDim.x, ++it) {
    if constexpr (PF == 2) load_img(img + 2 * (int64_t)gridDim.x, nx2);   // prefetch: in flight during the math below
    else load_img(img + gridDim.x, nxt);
    float a0[HW], a1[HW];
    {
      f32x2_t acc[HW], xin[HW];
      const f32x2_t bvp = pack_f32x2(bv.x, bv.y);
#pragma unroll
      for (int p = 0; p < HW; ++p) { acc[p] = bvp; xin[p] = x2_to_f32x2<XF16>(cur[p]); }   // HBM / latency-bound: exact cvt
#pragma unroll
      for (int ty = 0; ty < NT; ++ty) {
#pragma unroll
        for (int tx = 0; tx < NT; ++tx) {
          const f32x2_t w = *reinterpret_cast<const f32x2_t*>(wsm + (ty * NT + tx) * C + 2 * c2);
          const int dy = ty - R, dx = tx - R;                  // input = output + (dy, dx)
#pragma unroll
          for (int oy = 0; oy < S; ++oy) {
#pragma unroll
            for (int ox = 0; ox < S; ++ox) {
              const int iy = oy + dy, ix = ox + dx;
              if (iy >= 0 && iy < S && ix >= 0 && ix < S) fma_f32x2(acc[oy * S + ox], w, xin[iy * S + ix]);
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < HW; ++p) { const float2 a = unpack_f32x2(acc[p]); a0[p] = a.x; a1[p] = a.y; }
    }
    // ---- per-pixel channel statistics ------------------------------------------------------------------------------
    const int buf = it & 1;
    float* pb = part + (buf * nw + warp) * 32;
    if (HW == 1) {
      const float s = warp_sum(active ? a0[0] + a1[0] : 0.f);
      const float q = warp_sum(active ? fmaf(a0[0], a0[0], a1[0] * a1[0]) : 0.f);
      if (lane == 0) { pb[0] = s; pb[16] = q; }
    } else {
      float red[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) red[i] = 0.f;
      if (active) {
#pragma unroll
        for (int p = 0; p < HW; ++p) { red[p] = a0[p] + a1[p]; red[16 + p] = fmaf(a0[p], a0[p], a1[p] * a1[p]); }
      }
      seg_reduce32<32>(red, lane);
      pb[lane] = red[0];
    }
    __syncthreads();
    if (tid < HW) {
      float s = 0.f, q = 0.f;
      for (int w = 0; w < nw; ++w) { s += part[(buf * nw + w) * 32 + tid]; q += part[(buf * nw + w) * 32 + 16 + tid]; }
      const float mean = s * invC;
      const float var = fmaxf(q * invC - mean * mean, 0.f);
      stat[buf * 16 + tid] = make_float2(mean, rsqrtf(var + kLnEps));
    }
    __syncthreads();
    if (active) {
      uint32_t* dst = reinterpret_cast<uint32_t*>(out + img * (int64_t)HW * C) + c2;
#pragma unroll
      for (int p = 0; p < HW; ++p) {
        const float2 mr = stat[buf * 16 + p];
        __nv_bfloat162 o = __floats2bfloat162_rn((a0[p] - mr.x) * mr.y * gw.x + gb.x, (a1[p] - mr.x) * mr.y * gw.y + gb.y);
        dst[(size_t)p * C2] = *reinterpret_cast<uint32_t*>(&o);
      }
    }
#pragma unroll
    for (int p = 0; p < HW; ++p) {
      cur[p] = nxt[p];
      if constexpr (PF == 2) nxt[p] = nx2[p];
    }
  }
}

// ---- 3x3 maps, one WARP per image (C = 64 * NP: nano 320, pico 256) --------------------------------------------------
// The kernel above gives one channel pair to a thread and an image to a CTA, so the LayerNorm statistics cross warps:
// two CTA barriers and a shared-memory hop per image, and ~420 of each thread's ~500 instructions per image are not FFMA2
// (loads, conversions, the statistics tree, the LayerNorm application) -- it runs at 55 % of its issue floor.  Here a lane
// owns NP channel pairs of the image (pairs lane, lane + 32, ...), keeps all NP x 9 accumulators in registers, and the
// statistics of the 9 pixels need one 31-shuffle tree per IMAGE instead of per warp: no barrier, no shared-memory hop.
template <int C, bool XF16>
__global__ void __launch_bounds__(128)
dwln_w3_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, const float* __restrict__ wt, const float* __restrict__ bias,
               const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  constexpr int NP = C / 64, C2 = C / 2, HW = 9, NT = 5;
  extern __shared__ __align__(16) unsigned char sm[];
  float* wsm = reinterpret_cast<float*>(sm);                 // [25][C] reachable taps
  float* bsm = wsm + NT * NT * C;                            // conv bias, LN weight, LN bias: [3][C]
  const int tid = threadIdx.x, lane = tid & 31;
  {
    constexpr int kRun = NT * C / 4;                         // 16-byte granules per tap row (5 taps x C floats, contiguous)
    for (int i = tid; i < NT * kRun; i += 128) {
      const int ty = i / kRun, k = i - ty * kRun;
      const uint32_t d = (uint32_t)__cvta_generic_to_shared(wsm + ty * NT * C + 4 * k);
      const float* g = wt + ((ty + 1) * 7 + 1) * C + 4 * k;  // taps (ty + 1, 1 .. 5) of the 7x7 kernel
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < C; i += 128) { bsm[i] = __ldg(bias + i); bsm[C + i] = __ldg(ln_w + i); bsm[2 * C + i] = __ldg(ln_b + i); }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int64_t gw = (int64_t)blockIdx.x * 4 + (tid >> 5), nwarps = (int64_t)gridDim.x * 4;
  constexpr float invC = 1.0f / (float)C;
  for (int64_t img = gw; img < B; img += nwarps) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + img * (int64_t)HW * C);
    f32x2_t acc[NP][HW];
    uint32_t xr[HW];
#pragma unroll
    for (int p = 0; p < HW; ++p) xr[p] = __ldg(src + p * C2 + lane);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int pair = lane + 32 * j;
      f32x2_t xin[HW];
#pragma unroll
      for (int p = 0; p < HW; ++p) xin[p] = x2_to_f32x2<XF16>(xr[p]);
      if (j + 1 < NP) {                                      // the next pair's pixels are in flight during this pair's FMAs
#pragma unroll
        for (int p = 0; p < HW; ++p) xr[p] = __ldg(src + p * C2 + pair + 32);
      }
      const f32x2_t bv = *reinterpret_cast<const f32x2_t*>(bsm + 2 * pair);
#pragma unroll
      for (int p = 0; p < HW; ++p) acc[j][p] = bv;
#pragma unroll
      for (int ty = 0; ty < NT; ++ty) {
#pragma unroll
        for (int tx = 0; tx < NT; ++tx) {
          const f32x2_t w = *reinterpret_cast<const f32x2_t*>(wsm + (ty * NT + tx) * C + 2 * pair);
          const int dy = ty - 2, dx = tx - 2;                // input = output + (dy, dx)
#pragma unroll
          for (int oy = 0; oy < 3; ++oy) {
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
              const int iy = oy + dy, ix = ox + dx;
              if (iy >= 0 && iy < 3 && ix >= 0 && ix < 3) fma_f32x2(acc[j][oy * 3 + ox], w, xin[iy * 3 + ix]);
            }
          }
        }
      }
    }
    // ---- per-pixel statistics over the C channels: lane-local partial sums, one recursive-halving tree per image ----
    float red[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
#pragma unroll
      for (int p = 0; p < HW; ++p) {
        const float2 a = unpack_f32x2(acc[j][p]);
        red[p] += a.x + a.y;
        red[16 + p] = fmaf(a.x, a.x, fmaf(a.y, a.y, red[16 + p]));
      }
    }
    seg_reduce32<32>(red, lane);                             // lane l now holds the total of value l in red[0]
    float mean[HW], rstd[HW];
#pragma unroll
    for (int p = 0; p < HW; ++p) {
      const float s = __shfl_sync(0xffffffffu, red[0], p), q = __shfl_sync(0xffffffffu, red[0], 16 + p);
      mean[p] = s * invC;
      rstd[p] = rsqrtf(fmaxf(q * invC - mean[p] * mean[p], 0.f) + kLnEps);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + img * (int64_t)HW * C);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int pair = lane + 32 * j;
      const float2 gwv = *reinterpret_cast<const float2*>(bsm + C + 2 * pair);
      const float2 gbv = *reinterpret_cast<const float2*>(bsm + 2 * C + 2 * pair);
#pragma unroll
      for (int p = 0; p < HW; ++p) {
        const float2 a = unpack_f32x2(acc[j][p]);
        const __nv_bfloat162 o = __floats2bfloat162_rn((a.x - mean[p]) * rstd[p] * gwv.x + gbv.x,
                                                       (a.y - mean[p]) * rstd[p] * gwv.y + gbv.y);
        dst[p * C2 + pair] = *reinterpret_cast<const uint32_t*>(&o);
      }
    }
  }
}

// 1x1 maps (last stage): the 7x7 depthwise conv sees one pixel, i.e. y = LN(w_centre * x + b); a warp per image,
// C / 64 channel pairs per lane, two warp reductions for the statistics -- no shared memory at all.
template <int C, bool XF16>
__global__ void __launch_bounds__(256)
dwln_w1_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, const float* __restrict__ wt, const float* __restrict__ bias,
               const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  constexpr int NP = C / 64, C2 = C / 2;
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  f32x2_t w[NP], bv[NP];
  float2 gwv[NP], gbv[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const int pair = lane + 32 * j;
    w[j] = *reinterpret_cast<const f32x2_t*>(wt + 24 * C + 2 * pair);        // centre tap (ky = kx = 3) of the [49][C] matrix
    bv[j] = *reinterpret_cast<const f32x2_t*>(bias + 2 * pair);
    gwv[j] = *reinterpret_cast<const float2*>(ln_w + 2 * pair);
    gbv[j] = *reinterpret_cast<const float2*>(ln_b + 2 * pair);
  }
  constexpr float invC = 1.0f / (float)C;
  for (int64_t img = gw; img < B; img += nwarps) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + img * (int64_t)C);
    float2 a[NP];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      f32x2_t acc = bv[j];
      fma_f32x2(acc, w[j], x2_to_f32x2<XF16>(__ldg(src + lane + 32 * j)));
      a[j] = unpack_f32x2(acc);
      s += a[j].x + a[j].y;
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) { const float d0 = a[j].x - mean, d1 = a[j].y - mean; q = fmaf(d0, d0, fmaf(d1, d1, q)); }
    const float rstd = rsqrtf(warp_sum(q) * invC + kLnEps);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + img * (int64_t)C);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const __nv_bfloat162 o = __floats2bfloat162_rn((a[j].x - mean) * rstd * gwv[j].x + gbv[j].x,
                                                     (a[j].y - mean) * rstd * gwv[j].y + gbv[j].y);
      dst[lane + 32 * j] = *reinterpret_cast<const uint32_t*>(&o);
    }
  }
}

int num_sms();

template <int C, bool XF16>
static int launch_w1(const void* x, int64_t B, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                     void* out, cudaStream_t st) {
  const int64_t cap = (int64_t)num_sms() * 8, need = (B + 7) / 8;
  const int grid = (int)(need < cap ? need : cap);
  dwln_w1_kernel<C, XF16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, w, bias, ln_w, ln_b, (__nv_bfloat16*)out);
  return launch_done("dwln_w1");
}

template <int C, bool XF16>
static int launch_w3(const void* x, int64_t B, const float* w, const float* bias, const float* ln_w, const float* ln_b,
                     void* out, cudaStream_t st) {
  const size_t smem = (size_t)(25 + 3) * C * 4;
  auto kern = dwln_w3_kernel<C, XF16>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "dwln_w3 attr");
  int per_sm = 1;
  BTSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem), "dwln_w3 occupancy");
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)num_sms() * per_sm, need = (B + 3) / 4;
  const int grid = (int)(need < cap ? need : cap);
  kern<<<grid, 128, smem, st>>>((const __nv_bfloat16*)x, B, w, bias, ln_w, ln_b, (__nv_bfloat16*)out);
  return launch_done("dwln_w3");
}


template <int S, int CT, int PF, bool CPA = false, bool XF16 = false>
static int launch_small_pf(const void* x, int64_t B, int C, const float* w, const float* bias, const float* ln_w,
                        const float* ln_b, void* out, cudaStream_t st) {
  constexpr int NT = 2 * (S - 1) + 1;
  const int threads = ((C / 2 + 31) / 32) * 32;
  if (threads > 1024) return 1;
  const int nw = threads / 32;
  const size_t smem = (size_t)NT * NT * C * 4 + 2 * nw * 32 * 4 + 2 * 16 * 8;
  if (smem > 200 * 1024) return 1;
  auto kern = dwln_small_kernel<S, CT, PF, CPA, XF16>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "dwln_small attr");
  int per_sm = 1;
  BTSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem), "dwln_small occupancy");
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  const int grid = (int)(B < cap ? B : cap);
  kern<<<grid, threads, smem, st>>>((const __nv_bfloat16*)x, B, C, w, bias, ln_w, ln_b, (__nv_bfloat16*)out);
  return launch_done("dwln_small");
}

// cp.async tap staging (compile-time C, 16-byte aligned taps) is the default where it applies: 39 -> 33 us per 8192
// images at 3x3x320 (profiles/r02a); the two-image prefetch variant (PF = 2) measured slower (37 -> 39 us,
// profiles/r01n) and is not dispatched.
template <int S, int CT, bool XF16 = false>
static int launch_small(const void* x, int64_t B, int C, const float* w, const float* bias, const float* ln_w,
                        const float* ln_b, void* out, cudaStream_t st) {
  if constexpr (CT > 0 && CT % 4 == 0) {
    if (((uintptr_t)w % 16) == 0) return launch_small_pf<S, CT, 1, true, XF16>(x, B, C, w, bias, ln_w, ln_b, out, st);
  }
  return launch_small_pf<S, CT, 1, false, XF16>(x, B, C, w, bias, ln_w, ln_b, out, st);
}

// returns 1 if the shape is not handled here
int dwln_bf16_small(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                    const float* ln_b, void* out, bool xf16, cudaStream_t st) {
  if (H != W || (C & 1) || C > 2048) return 1;
  if (((uintptr_t)x % 4) != 0 || ((uintptr_t)out % 4) != 0 || ((uintptr_t)bias % 8) != 0 || ((uintptr_t)ln_w % 8) != 0 ||
      ((uintptr_t)ln_b % 8) != 0)
    return 1;
  // 3x3 maps at the nano / pico widths: one warp per image (BTSB_DWLN_W3=0 keeps the thread-per-pair kernel for A/B)
  static const bool w3 = !(getenv("BTSB_DWLN_W3") && atoi(getenv("BTSB_DWLN_W3")) == 0);
  if (w3 && H == 3 && (C == 320 || C == 256) && ((uintptr_t)w % 16) == 0) {
    if (C == 320) return xf16 ? launch_w3<320, true>(x, B, w, bias, ln_w, ln_b, out, st)
                              : launch_w3<320, false>(x, B, w, bias, ln_w, ln_b, out, st);
    return xf16 ? launch_w3<256, true>(x, B, w, bias, ln_w, ln_b, out, st)
                : launch_w3<256, false>(x, B, w, bias, ln_w, ln_b, out, st);
  }
  if (w3 && H == 1 && (C == 640 || C == 512) && ((uintptr_t)w % 8) == 0) {
    if (C == 640) return xf16 ? launch_w1<640, true>(x, B, w, bias, ln_w, ln_b, out, st)
                              : launch_w1<640, false>(x, B, w, bias, ln_w, ln_b, out, st);
    return xf16 ? launch_w1<512, true>(x, B, w, bias, ln_w, ln_b, out, st)
                : launch_w1<512, false>(x, B, w, bias, ln_w, ln_b, out, st);
  }
  if (xf16) {
    if (H == 3) {
      if (C == 320) return launch_small<3, 320, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
      if (C == 256) return launch_small<3, 256, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
      return launch_small<3, 0, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    }
    if (H == 1) return launch_small<1, 0, true>(x, B, C, w, bias, ln_w, ln_b, out, st);
    return 1;
  }
  if (H == 3) {
    if (C == 320) return launch_small<3, 320>(x, B, C, w, bias, ln_w, ln_b, out, st);
    if (C == 256) return launch_small<3, 256>(x, B, C, w, bias, ln_w, ln_b, out, st);
    return launch_small<3, 0>(x, B, C, w, bias, ln_w, ln_b, out, st);
  }
  if (H == 1) {
    if (C == 640) return launch_small<1, 640>(x, B, C, w, bias, ln_w, ln_b, out, st);
    if (C == 512) return launch_small<1, 512>(x, B, C, w, bias, ln_w, ln_b, out, st);
    return launch_small<1, 0>(x, B, C, w, bias, ln_w, ln_b, out, st);
  }
  return 1;
}

}  // namespace btsb
