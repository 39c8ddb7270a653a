// K3 v2 (bf16 activations, square maps 15/7/3/1): depthwise 7x7 + bias + LayerNorm2d.
//
//   * persistent CTAs; the next image group is prefetched with cp.async while the current one is convolved
//   * a thread owns a CHANNEL PAIR (one 32-bit shared-memory load feeds two channels) and one output row;
//     the row of S outputs is register-blocked and out-of-range taps are pruned at compile time
//     (15x15: 93 of 105 FMAs per input row, 3x3: 9 of 21, 1x1: a single tap)
//   * only the reachable taps (|d| < S) are staged in shared memory ([49][C] fp32 would not fit at C = 640)
//   * fp32 conv results are staged once in shared memory, then one warp per pixel runs a two-pass LayerNorm
//     and writes bf16x2 rows (coalesced)
// The generic kernel in convnext_simt.cu (dwln_kernel) remains the path for fp32 and for other map shapes.
#include "common.cuh"

namespace btsb {

constexpr int kDw2MaxThreads = 608;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int S>
__global__ void __launch_bounds__(kDw2MaxThreads, 1)
dwln2_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int C, int G, const float* __restrict__ wt,
             const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
             __nv_bfloat16* __restrict__ out) {
  constexpr int R = S > 3 ? 3 : S - 1;      // reachable tap radius
  constexpr int NT = 2 * R + 1;
  constexpr int HW = S * S;
  extern __shared__ __align__(16) unsigned char sm[];
  float* wsm = reinterpret_cast<float*>(sm);                 // [NT*NT][C]
  float* bsm = wsm + NT * NT * C;                            // conv bias [C]
  float* gsm = bsm + C;                                      // LN weight
  float* hsm = gsm + C;                                      // LN bias
  float* conv = hsm + C;                                     // [G*HW][C] fp32
  __nv_bfloat16* tin = reinterpret_cast<__nv_bfloat16*>(conv + (size_t)G * HW * C);   // [2][G*HW*C]

  const int tid = threadIdx.x, T = blockDim.x;
  const int img_elems = HW * C;
  const int ngroups = (int)((B + G - 1) / G);

  auto issue = [&](int grp, int buf) {
    const int64_t b0 = (int64_t)grp * G;
    const int cnt = (int)min((int64_t)G, B - b0);
    const int nvec = cnt * img_elems / 8;                    // 16-byte chunks
    const uint4* src = reinterpret_cast<const uint4*>(x + b0 * img_elems);
    uint4* dst = reinterpret_cast<uint4*>(tin + (size_t)buf * G * img_elems);
    for (int i = tid; i < nvec; i += T) cp_async16(dst + i, src + i);
  };

  if ((int)blockIdx.x < ngroups) issue(blockIdx.x, 0);
  cp_async_commit();
  for (int i = tid; i < NT * NT * C; i += T) {
    const int t = i / C, c = i - t * C;
    const int ty = t / NT, tx = t - ty * NT;
    wsm[i] = __ldg(wt + ((ty + 3 - R) * 7 + (tx + 3 - R)) * C + c);
  }
  for (int i = tid; i < C; i += T) { bsm[i] = __ldg(bias + i); gsm[i] = __ldg(ln_w + i); hsm[i] = __ldg(ln_b + i); }

  const int C2 = C >> 1;
  const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  int it = 0;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, ++it) {
    const int buf = it & 1;
    const int nxt = grp + gridDim.x;
    if (nxt < ngroups) issue(nxt, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();                                      // everything but the newest group has landed
    __syncthreads();

    const int64_t b0 = (int64_t)grp * G;
    const int cnt = (int)min((int64_t)G, B - b0);
    const __nv_bfloat16* tbuf = tin + (size_t)buf * G * img_elems;
    const int items = cnt * S * C2;
    for (int item = tid; item < items; item += T) {
      const int c2 = item % C2;
      const int r = item / C2;
      const int oy = r % S, g = r / S;
      float acc0[S], acc1[S];
      const float2 bv = *reinterpret_cast<const float2*>(bsm + 2 * c2);
#pragma unroll
      for (int t = 0; t < S; ++t) { acc0[t] = bv.x; acc1[t] = bv.y; }
      const __nv_bfloat16* img = tbuf + (size_t)g * img_elems + 2 * c2;
#pragma unroll
      for (int dy = -R; dy <= R; ++dy) {
        const int iy = oy + dy;
        if (iy < 0 || iy >= S) continue;
        float2 wv[NT];
#pragma unroll
        for (int kx = 0; kx < NT; ++kx) wv[kx] = *reinterpret_cast<const float2*>(wsm + ((dy + R) * NT + kx) * C + 2 * c2);
#pragma unroll
        for (int ix = 0; ix < S; ++ix) {
          const uint32_t v = *reinterpret_cast<const uint32_t*>(img + (size_t)(iy * S + ix) * C);
          const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
#pragma unroll
          for (int kx = 0; kx < NT; ++kx) {
            const int t = ix - (kx - R);                      // output column fed by this (input, tap)
            if (t >= 0 && t < S) { acc0[t] = fmaf(wv[kx].x, lo, acc0[t]); acc1[t] = fmaf(wv[kx].y, hi, acc1[t]); }
          }
        }
      }
      float* dst = conv + ((size_t)(g * HW + oy * S)) * C + 2 * c2;
#pragma unroll
      for (int t = 0; t < S; ++t) *reinterpret_cast<float2*>(dst + (size_t)t * C) = make_float2(acc0[t], acc1[t]);
    }
    __syncthreads();

    // LayerNorm over channels: one warp per pixel, two-pass over the staged fp32 row (re-read from shared memory, so
    // no per-lane register array sized for the widest C), bf16x2 stores
    const int npix = cnt * HW;
    const float invC = 1.0f / (float)C;
    for (int p = wid; p < npix; p += nw) {
      const float2* v = reinterpret_cast<const float2*>(conv + (size_t)p * C);
      float s = 0.f;
      for (int k2 = lane; k2 < C2; k2 += 32) { const float2 t = v[k2]; s += t.x + t.y; }
      const float mean = warp_sum(s) * invC;
      float q = 0.f;
      for (int k2 = lane; k2 < C2; k2 += 32) { const float2 t = v[k2]; const float d0 = t.x - mean, d1 = t.y - mean; q += d0 * d0 + d1 * d1; }
      const float rstd = rsqrtf(warp_sum(q) * invC + kLnEps);
      uint32_t* dst = reinterpret_cast<uint32_t*>(out + (b0 * HW + p) * (int64_t)C);
      for (int k2 = lane; k2 < C2; k2 += 32) {
        const float2 t = v[k2];
        const float2 gw = *reinterpret_cast<const float2*>(gsm + 2 * k2);
        const float2 gb = *reinterpret_cast<const float2*>(hsm + 2 * k2);
        __nv_bfloat162 o = __floats2bfloat162_rn((t.x - mean) * rstd * gw.x + gb.x, (t.y - mean) * rstd * gw.y + gb.y);
        dst[k2] = *reinterpret_cast<uint32_t*>(&o);
      }
    }
    // the next iteration's top-of-loop __syncthreads orders these conv reads before the next conv writes
  }
  cp_async_wait<0>();
}

int num_sms();

template <int S>
static int launch_dwln2(const void* x, int64_t B, int C, const float* w, const float* bias, const float* ln_w,
                        const float* ln_b, void* out, cudaStream_t st) {
  constexpr int R = S > 3 ? 3 : S - 1;
  constexpr int NT = 2 * R + 1;
  constexpr int HW = S * S;
  const size_t wbytes = (size_t)(NT * NT + 3) * C * 4;
  const size_t per_img = (size_t)HW * C * (4 + 2 * 2);
  const size_t budget = 200 * 1024;
  BTSB_REQUIRE(wbytes + per_img <= 225 * 1024, "dwln2: map %dx%dx%d does not fit in shared memory", S, S, C);
  int G = (int)((budget - wbytes) / per_img);
  if (G < 1) G = 1;
  if (G > 32) G = 32;
  const int sms = num_sms();
  while (G > 1 && (B + G - 1) / G < 2 * sms) G = (G + 1) / 2;
  const int items = G * S * (C / 2);
  int bestT = 256; double bestU = 0.0;
  for (int T = kDw2MaxThreads; T >= 256; T -= 32) {
    const int rounds = (items + T - 1) / T;
    const double u = (double)items / ((double)rounds * T);
    if (u > bestU + 1e-9) { bestU = u; bestT = T; }
  }
  const size_t smem = wbytes + (size_t)G * per_img;
  auto kern = dwln2_kernel<S>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "dwln2 attr");
  const int ngroups = (int)((B + G - 1) / G);
  const int grid = ngroups < sms ? ngroups : sms;
  kern<<<grid, bestT, smem, st>>>((const __nv_bfloat16*)x, B, C, G, w, bias, ln_w, ln_b, (__nv_bfloat16*)out);
  return launch_done("dwln2");
}

// returns 1 if the shape is not handled here (caller falls back to the generic kernel)
int dwln_bf16_v2(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, cudaStream_t st) {
  if (H != W || C % 16 != 0 || C > 640) return 1;
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 4) != 0) return 1;
  switch (H) {
    case 15: return launch_dwln2<15>(x, B, C, w, bias, ln_w, ln_b, out, st);
    case 7: return launch_dwln2<7>(x, B, C, w, bias, ln_w, ln_b, out, st);
    case 3: return launch_dwln2<3>(x, B, C, w, bias, ln_w, ln_b, out, st);
    case 1: return launch_dwln2<1>(x, B, C, w, bias, ln_w, ln_b, out, st);
    default: return 1;
  }
}

}  // namespace btsb
