// CUDA-core kernels of the ConvNeXt path: stem conv+LN (K2a), depthwise 7x7+LN (K3), LN+2x2 patch gather (K5a),
// pool+LN head prologue, fp32 GEMM with fused epilogues (the 1e-4 "correctness" mode of K4/K5b), scoring epilogue.
// Activations are NHWC pixel rows [B*H*W, C]; see include/btsbot_b200.h for the contracts and reference citations.
#include <stdlib.h>

#include "common.cuh"

namespace btsb {

// =====================================================================================================
// K2a  stem: Conv2d(3,C0,k4,s4)+bias -> LayerNorm over C0, NCHW fp32 in -> NHWC rows out
//   one warp per group of 4 horizontally adjacent output pixels; lane owns channels lane, lane+32, ...
// =====================================================================================================
constexpr int kStemThreads = 256;
constexpr int kStemPix = 4;      // pixels per warp pass
constexpr int kStemMaxCh = 4;    // channels per lane (C0 <= 128)

template <typename TO>
__global__ void __launch_bounds__(kStemThreads)
stem_kernel(const float* __restrict__ x, int64_t B, int H, int W, int ho, int wo, const float* __restrict__ wt,
            const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b, int C0,
            TO* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                                  // [48][C0]
  float* patch = smem + 48 * C0;                     // [warps][kStemPix][48]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 48 * C0; i += kStemThreads) ws[i] = wt[i];
  __syncthreads();
  float* mypatch = patch + wid * (kStemPix * 48);

  const int wgroups = (wo + kStemPix - 1) / kStemPix;          // pixel groups per output row
  const int64_t total = B * (int64_t)ho * wgroups;
  const int nwarps = kStemThreads / 32;
  float bch[kStemMaxCh], gch[kStemMaxCh], hch[kStemMaxCh];
#pragma unroll
  for (int j = 0; j < kStemMaxCh; ++j) {
    const int c = lane + 32 * j;
    bch[j] = c < C0 ? bias[c] : 0.f; gch[j] = c < C0 ? ln_w[c] : 0.f; hch[j] = c < C0 ? ln_b[c] : 0.f;
  }
  for (int64_t g = (int64_t)blockIdx.x * nwarps + wid; g < total; g += (int64_t)gridDim.x * nwarps) {
    const int gx = (int)(g % wgroups);
    const int64_t t = g / wgroups;
    const int oy = (int)(t % ho);
    const int64_t b = t / ho;
    const int ox0 = gx * kStemPix;
    // gather 4 patches x 48 values: k = (ci*4+ky)*4+kx
    for (int i = lane; i < kStemPix * 48; i += 32) {
      const int p = i / 48, k = i - p * 48;
      const int ci = k >> 4, ky = (k >> 2) & 3, kx = k & 3;
      const int ox = ox0 + p;
      float v = 0.f;
      if (ox < wo) v = __ldg(x + ((b * 3 + ci) * H + (oy * 4 + ky)) * (int64_t)W + ox * 4 + kx);
      mypatch[i] = v;
    }
    __syncwarp();
    float acc[kStemPix][kStemMaxCh];
#pragma unroll
    for (int p = 0; p < kStemPix; ++p)
#pragma unroll
      for (int j = 0; j < kStemMaxCh; ++j) acc[p][j] = bch[j];
#pragma unroll 4
    for (int k = 0; k < 48; ++k) {
      float wv[kStemMaxCh];
#pragma unroll
      for (int j = 0; j < kStemMaxCh; ++j) { const int c = lane + 32 * j; wv[j] = c < C0 ? ws[k * C0 + c] : 0.f; }
#pragma unroll
      for (int p = 0; p < kStemPix; ++p) {
        const float pv = mypatch[p * 48 + k];
#pragma unroll
        for (int j = 0; j < kStemMaxCh; ++j) acc[p][j] = fmaf(pv, wv[j], acc[p][j]);
      }
    }
    __syncwarp();
    // LayerNorm over channels (two-pass), one pixel at a time
#pragma unroll
    for (int p = 0; p < kStemPix; ++p) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kStemMaxCh; ++j) if (lane + 32 * j < C0) s += acc[p][j];
      const float mean = warp_sum(s) / (float)C0;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < kStemMaxCh; ++j) if (lane + 32 * j < C0) { const float d = acc[p][j] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / (float)C0 + kLnEps);
      const int ox = ox0 + p;
      if (ox < wo) {
        TO* dst = out + ((b * ho + oy) * (int64_t)wo + ox) * C0;
#pragma unroll
        for (int j = 0; j < kStemMaxCh; ++j) {
          const int c = lane + 32 * j;
          if (c < C0) stf(dst + c, (acc[p][j] - mean) * rstd * gch[j] + hch[j]);
        }
      }
    }
  }
}

// =====================================================================================================
// K3  depthwise 7x7 (pad 3) + bias + LayerNorm2d.  One CTA per group of G images, blockDim = r*C so each
// thread owns one channel (49 taps in registers) and walks (image, row) items; a row of TW outputs is
// register-blocked so every shared-memory load feeds up to 7 FMAs.  Conv results go to a second smem
// buffer, then one warp per pixel does a two-pass LayerNorm and writes coalesced NHWC rows.
// =====================================================================================================
template <int TW> struct DwBounds { static constexpr int kMaxThreads = (TW >= 8) ? 320 : 640; };

template <typename T, int TW, typename TO = T>          // TO != T: fp16 residual-stream rows in, bf16 rows out
__global__ void __launch_bounds__(DwBounds<TW>::kMaxThreads)
dwln_kernel(const T* __restrict__ x, int64_t B, int H, int W, int C, int G, const float* __restrict__ wt,
            const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
            TO* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int HW = H * W;
  float* conv = reinterpret_cast<float*>(smem_raw);                                  // [G*HW][C] fp32
  T* tin = reinterpret_cast<T*>(smem_raw + (size_t)G * HW * C * sizeof(float));      // [G*HW][C]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * G;
  const int gcount = (int)min((int64_t)G, B - b0);
  const int nelem = gcount * HW * C;
  const T* src = x + b0 * (int64_t)HW * C;
  {  // 16-byte vector copy (C*sizeof(T) is a multiple of 16, base pointers are 16B aligned)
    const int nvec = (int)((size_t)nelem * sizeof(T) / 16);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(tin);
    for (int i = tid; i < nvec; i += nthr) d4[i] = __ldg(s4 + i);
  }

  const int c = tid % C;
  const int rsub = tid / C, rcount = nthr / C;
  float wreg[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wreg[k] = __ldg(wt + k * C + c);
  const float bc = __ldg(bias + c);
  __syncthreads();

  const int xchunks = (W + TW - 1) / TW;
  const int items = gcount * H * xchunks;
  for (int it = rsub; it < items; it += rcount) {
    const int xc = it % xchunks;
    const int t2 = it / xchunks;
    const int oy = t2 % H, g = t2 / H;
    const int x0 = xc * TW;
    float acc[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) acc[t] = bc;
    const T* img = tin + (size_t)g * HW * C + c;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy + ky - 3;
      if (iy < 0 || iy >= H) continue;
      float row[TW + 6];
#pragma unroll
      for (int t = 0; t < TW + 6; ++t) {
        const int ix = x0 + t - 3;
        row[t] = (ix >= 0 && ix < W) ? ldf(img + (size_t)(iy * W + ix) * C) : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 7; ++kx)
#pragma unroll
        for (int t = 0; t < TW; ++t) acc[t] = fmaf(wreg[ky * 7 + kx], row[t + kx], acc[t]);
    }
    float* dst = conv + ((size_t)g * HW + oy * W + x0) * C + c;
#pragma unroll
    for (int t = 0; t < TW; ++t)
      if (x0 + t < W) dst[(size_t)t * C] = acc[t];
  }
  __syncthreads();

  // LayerNorm: one warp per pixel
  const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  const int npix = gcount * HW;
  for (int p = wid; p < npix; p += nw) {
    const float* v = conv + (size_t)p * C;
    float s = 0.f;
    for (int k = lane; k < C; k += 32) s += v[k];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
    for (int k = lane; k < C; k += 32) { const float d = v[k] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
    TO* dst = out + (b0 * HW + p) * (int64_t)C;
    for (int k = lane; k < C; k += 32) stf(dst + k, (v[k] - mean) * rstd * __ldg(ln_w + k) + __ldg(ln_b + k));
  }
}

// =====================================================================================================
// K5a  LayerNorm2d + 2x2/s2 patch gather: one warp per *used* input pixel.
// =====================================================================================================
template <typename T, typename TO = T>
__global__ void __launch_bounds__(256)
lnpatch_kernel(const T* __restrict__ x, int64_t B, int H, int W, int C, int Ho, int Wo,
               const float* __restrict__ ln_w, const float* __restrict__ ln_b, TO* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Hu = 2 * Ho, Wu = 2 * Wo;
  const int64_t total = B * (int64_t)Hu * Wu;
  constexpr int MAXJ = 20;   // C <= 640
  for (int64_t p = warp; p < total; p += nwarps) {
    const int ix = (int)(p % Wu);
    const int64_t t = p / Wu;
    const int iy = (int)(t % Hu);
    const int64_t b = t / Hu;
    const T* src = x + ((b * H + iy) * (int64_t)W + ix) * C;
    float v[MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int k = lane + 32 * j;
      v[j] = k < C ? ldf(src + k) : 0.f;
      s += v[j];
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) if (lane + 32 * j < C) { const float d = v[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
    const int oy = iy >> 1, dy = iy & 1, ox = ix >> 1, dx = ix & 1;
    TO* dst = out + ((b * Ho + oy) * (int64_t)Wo + ox) * (4 * (int64_t)C) + (dy * 2 + dx) * C;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int k = lane + 32 * j;
      if (k < C) stf(dst + k, (v[j] - mean) * rstd * __ldg(ln_w + k) + __ldg(ln_b + k));
    }
  }
}

// bf16 specialisation: a lane owns channel PAIRS (32-bit loads/stores); MAXJ = ceil(C/64) pairs per lane.
template <int MAXJ>
__global__ void __launch_bounds__(256)
lnpatch_bf16x2_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int H, int W, int C, int Ho, int Wo,
                      const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int Hu = 2 * Ho, Wu = 2 * Wo, C2 = C >> 1;
  const int64_t total = B * (int64_t)Hu * Wu;
  float2 gw[MAXJ], gb[MAXJ];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    const int k2 = lane + 32 * j;
    gw[j] = k2 < C2 ? __ldg(reinterpret_cast<const float2*>(ln_w) + k2) : make_float2(0.f, 0.f);
    gb[j] = k2 < C2 ? __ldg(reinterpret_cast<const float2*>(ln_b) + k2) : make_float2(0.f, 0.f);
  }
  for (int64_t p = warp; p < total; p += nwarps) {
    const int ix = (int)(p % Wu);
    const int64_t t = p / Wu;
    const int iy = (int)(t % Hu);
    const int64_t b = t / Hu;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + ((b * H + iy) * (int64_t)W + ix) * C);
    float2 v[MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int k2 = lane + 32 * j;
      const uint32_t u = k2 < C2 ? __ldg(src + k2) : 0u;
      v[j] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
      s += v[j].x + v[j].y;
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (lane + 32 * j < C2) { const float d0 = v[j].x - mean, d1 = v[j].y - mean; q += d0 * d0 + d1 * d1; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
    const int oy = iy >> 1, dy = iy & 1, ox = ix >> 1, dx = ix & 1;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((b * Ho + oy) * (int64_t)Wo + ox) * (4 * (int64_t)C) + (dy * 2 + dx) * C);
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int k2 = lane + 32 * j;
      if (k2 < C2) {
        __nv_bfloat162 o = __floats2bfloat162_rn((v[j].x - mean) * rstd * gw[j].x + gb[j].x,
                                                 (v[j].y - mean) * rstd * gw[j].y + gb[j].y);
        dst[k2] = *reinterpret_cast<uint32_t*>(&o);
      }
    }
  }
}

// =====================================================================================================
// head prologue: global average pool over HW + optional LayerNorm -> [B,C] fp32; one warp per image.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
poolln_kernel(const T* __restrict__ x, int64_t B, int HW, int C, const float* __restrict__ ln_w,
              const float* __restrict__ ln_b, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int MAXJ = 20;
  const float inv = 1.0f / (float)HW;
  for (int64_t b = warp; b < B; b += nwarps) {
    float v[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) v[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      const T* src = x + (b * HW + p) * (int64_t)C;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) { const int k = lane + 32 * j; if (k < C) v[j] += ldf(src + k); }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) { v[j] *= inv; s += v[j]; }
    float mean = 0.f, rstd = 1.f;
    if (ln_w) {
      mean = warp_sum(s) / (float)C;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) if (lane + 32 * j < C) { const float d = v[j] - mean; q += d * d; }
      rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
    }
    float* dst = out + b * (int64_t)C;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int k = lane + 32 * j;
      if (k < C) dst[k] = ln_w ? (v[j] - mean) * rstd * __ldg(ln_w + k) + __ldg(ln_b + k) : v[j];
    }
  }
}

// =====================================================================================================
// fp32 GEMM  out[M,N] = epi(A[M,K] . Wt[N,K]^T + bias)   (CUDA cores; the 1e-4 path)
// 128x64 tile, BK=16, 256 threads, 8x4 micro-tile.
// =====================================================================================================
constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ bias,
             const float* __restrict__ gamma, const float* __restrict__ res, float* __restrict__ out, int64_t M,
             int N, int K) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int tx = tid & 15, ty = tid >> 4;     // tx -> 4 columns, ty -> 8 rows
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // A tile: 128 rows x 16 k = 512 float4 -> 2 per thread
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * 256;
      const int row = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t gm = m0 + row;
      if (gm < M && k0 + kq < K) v = *reinterpret_cast<const float4*>(A + gm * K + k0 + kq);
      As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
    }
    {
      const int row = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int gn = n0 + row;
      if (gn < N && k0 + kq < K) v = *reinterpret_cast<const float4*>(Wt + (int64_t)gn * K + k0 + kq);
      Ws[kq + 0][row] = v.x; Ws[kq + 1][row] = v.y; Ws[kq + 2][row] = v.z; Ws[kq + 3][row] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[8], b[4];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + bias[gn];
      if (EPI == BTSB_EPI_BIAS_GELU) v = gelu_erf(v);
      if (EPI == BTSB_EPI_BIAS_SILU) v = v / (1.0f + expf(-v));
      if (EPI == BTSB_EPI_SCALE_RES) v = res[gm * N + gn] + gamma[gn] * v;
      out[gm * N + gn] = v;
    }
  }
}

__global__ void score_kernel(const float* __restrict__ logits, int64_t B, float* __restrict__ scores,
                             uint8_t* __restrict__ labels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float s = 1.0f / (1.0f + expf(-logits[i]));
  if (scores) scores[i] = s;
  if (labels) labels[i] = s > 0.5f ? 1 : 0;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

static int pick_grid(int64_t work_items, int per_block, int cap = 148 * 16) {
  int64_t g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace btsb

using namespace btsb;

extern "C" int btsb_convnext_stem_fwd(const float* x, int64_t B, int H, int W, const float* w, const float* bias,
                                      const float* ln_w, const float* ln_b, int C0, void* out, int out_dtype,
                                      void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && H >= 4 && W >= 4, "stem: bad shape B=%lld H=%d W=%d", (long long)B, H, W);
  BTSB_REQUIRE(C0 >= 1 && C0 <= 32 * kStemMaxCh, "stem: C0=%d not in [1,128]", C0);
  BTSB_REQUIRE(out_dtype == BTSB_F32 || out_dtype == BTSB_BF16, "stem: out_dtype must be F32 or BF16");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && w && bias && ln_w && ln_b && out, "stem: null pointer");
  const int ho = (H - 4) / 4 + 1, wo = (W - 4) / 4 + 1;
  const int wgroups = (wo + kStemPix - 1) / kStemPix;
  const int64_t total = B * (int64_t)ho * wgroups;
  const int nwarps = kStemThreads / 32;
  const int grid = pick_grid(total, nwarps * 4);
  const int smem = (48 * C0 + nwarps * kStemPix * 48) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == BTSB_F32)
    stem_kernel<float><<<grid, kStemThreads, smem, st>>>(x, B, H, W, ho, wo, w, bias, ln_w, ln_b, C0, (float*)out);
  else
    stem_kernel<__nv_bfloat16><<<grid, kStemThreads, smem, st>>>(x, B, H, W, ho, wo, w, bias, ln_w, ln_b, C0,
                                                                 (__nv_bfloat16*)out);
  return launch_done("stem");
}

template <typename T, int TW, typename TO = T>
static int launch_dwln(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias,
                       const float* ln_w, const float* ln_b, void* out, cudaStream_t st) {
  // threads: multiple of C in [256, 640]
  int r = 1;
  while (C * r < 256) ++r;
  const int threads = C * r;
  BTSB_REQUIRE(threads <= DwBounds<TW>::kMaxThreads && threads % 32 == 0,
               "dwln: C=%d unsupported at map width %d (need a multiple of C and of 32 in [256,%d] threads)", C, W,
               DwBounds<TW>::kMaxThreads);
  BTSB_REQUIRE((C * sizeof(T)) % 16 == 0 && ((uintptr_t)x % 16) == 0, "dwln: rows must be 16-byte aligned");
  const int HW = H * W;
  const size_t per_img = (size_t)HW * C * (sizeof(float) + sizeof(T));
  const size_t budget = 100 * 1024;          // <= ~100 KB so two CTAs fit per SM
  int G = (int)(budget / per_img);
  if (G < 1) G = 1;
  // keep enough CTAs to fill the machine
  while (G > 1 && (B + G - 1) / G < 2 * 148) G = (G + 1) / 2;
  if (G > 64) G = 64;
  const size_t smem = (size_t)G * per_img;
  BTSB_REQUIRE(smem <= 227 * 1024, "dwln: map %dx%dx%d needs %zu B of shared memory (> 227 KB)", H, W, C, smem);
  auto kern = dwln_kernel<T, TW, TO>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)), "dwln attr");
  const int64_t grid = (B + G - 1) / G;
  kern<<<(unsigned)grid, threads, smem, st>>>((const T*)x, B, H, W, C, G, w, bias, ln_w, ln_b, (TO*)out);
  return launch_done("dwln");
}

template <typename T, typename TO = T>
static int dispatch_dwln(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias,
                         const float* ln_w, const float* ln_b, void* out, cudaStream_t st) {
  if (W == 15) return launch_dwln<T, 15, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  if (W == 7) return launch_dwln<T, 7, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  if (W == 3) return launch_dwln<T, 3, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  if (W == 1) return launch_dwln<T, 1, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  // any other map width (e.g. the 19 / 9 / 4 / 2 maps of larger "LS" cutouts): 8 outputs per thread while the CTA
  // (a multiple of C threads) fits that variant's register budget, else the 3-output variant (up to 640 threads)
  int r = 1;
  while (C * r < 256) ++r;
  if (C * r <= DwBounds<8>::kMaxThreads) return launch_dwln<T, 8, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  return launch_dwln<T, 3, TO>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
}

namespace btsb {
int dwln_bf16_v2(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, cudaStream_t st);
int dwln_bf16_v3(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, bool xf16, cudaStream_t st);
int dwln_bf16_small(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                    const float* ln_b, void* out, bool xf16, cudaStream_t st);
int dwln_bf16_v5(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                 const float* ln_b, void* out, bool xf16, cudaStream_t st);
int dwln_f32_v3(const void* x, int64_t B, int H, int W, int C, const float* w, const float* bias, const float* ln_w,
                const float* ln_b, void* out, cudaStream_t st);
}
extern "C" int btsb_convnext_dwln_fwd(const void* x, int dtype, int64_t B, int H, int W, int C, const float* w,
                                      const float* bias, const float* ln_w, const float* ln_b, void* out,
                                      void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 32, "dwln: bad shape");
  BTSB_REQUIRE(dtype == BTSB_F32 || dtype == BTSB_BF16 || dtype == BTSB_BF16_XF16, "dwln: dtype must be F32, BF16 or BF16_XF16");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && w && bias && ln_w && ln_b && out, "dwln: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_F32) {
    const int rc = dwln_f32_v3(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);       // 15 x 15 / 7 x 7 at the nano / pico widths
    if (rc != 1) return rc;
    return dispatch_dwln<float>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  }
  const bool xf16 = dtype == BTSB_BF16_XF16;       // x: fp16 residual-stream rows; out: bf16 (the fc1 operand) either way
  // each specialised kernel returns 1 when the shape is not its own: small maps -> v5 -> v3 -> v2 -> generic
  {
    const int rc = dwln_bf16_small(x, B, H, W, C, w, bias, ln_w, ln_b, out, xf16, st);   // 3x3 and 1x1 maps
    if (rc != 1) return rc;
  }
  {
    const int rc = dwln_bf16_v5(x, B, H, W, C, w, bias, ln_w, ln_b, out, xf16, st);      // conv and LayerNorm on different warps
    if (rc != 1) return rc;
  }
  {
    const int rc = dwln_bf16_v3(x, B, H, W, C, w, bias, ln_w, ln_b, out, xf16, st);
    if (rc != 1) return rc;
  }
  if (xf16) return dispatch_dwln<__half, __nv_bfloat16>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
  {
    const int rc = dwln_bf16_v2(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
    if (rc != 1) return rc;
  }
  return dispatch_dwln<__nv_bfloat16>(x, B, H, W, C, w, bias, ln_w, ln_b, out, st);
}

// bf16 fast path for C = 8*NV*TPP (nano 80/160/320, pico 64/128/256): TPP threads per used input pixel, each thread keeps
// its NV uint4 (8*NV channels, contiguous 16*NV bytes) in registers -- 128-bit loads/stores, thread-local two-pass
// statistics, only log2(TPP) shuffles.  ~18 warp-instructions per pixel instead of ~200 for the warp-per-pixel kernels
// above (which were issue-bound at 1.4-1.7 TB/s: profiles/r01c misc.summary).
template <int NV, int TPP, bool XF16 = false>          // XF16: x holds IEEE fp16 residual-stream rows (out: bf16)
__global__ void __launch_bounds__(256)
lnpatch_tpp_kernel(const __nv_bfloat16* __restrict__ x, int64_t B, int H, int W, int Ho, int Wo,
                   const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out) {
  constexpr int CH = 8 * NV, C = CH * TPP;
  __shared__ __align__(16) float gws[C], gbs[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { gws[i] = __ldg(ln_w + i); gbs[i] = __ldg(ln_b + i); }
  __syncthreads();
  const int Hu = 2 * Ho, Wu = 2 * Wo;
  const int64_t total = B * (int64_t)Hu * Wu * TPP;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int sub = (int)(i % TPP);
    const int64_t p = i / TPP;
    const int ix = (int)(p % Wu);
    const int64_t t = p / Wu;
    const int iy = (int)(t % Hu);
    const int64_t b = t / Hu;
    const uint4* src = reinterpret_cast<const uint4*>(x + ((b * H + iy) * (int64_t)W + ix) * C + sub * CH);
    uint4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = __ldg(src + j);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) s += x2_lo<XF16>(u[k]) + x2_hi<XF16>(u[k]);
    }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d0 = x2_lo<XF16>(u[k]) - mean, d1 = x2_hi<XF16>(u[k]) - mean;
        q = fmaf(d0, d0, q); q = fmaf(d1, d1, q);
      }
    }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / (float)C) + kLnEps);
    const int oy = iy >> 1, dy = iy & 1, ox = ix >> 1, dx = ix & 1;
    uint4* dst = reinterpret_cast<uint4*>(out + ((b * Ho + oy) * (int64_t)Wo + ox) * (4 * (int64_t)C) + (dy * 2 + dx) * C + sub * CH);
    const float4* gw4 = reinterpret_cast<const float4*>(gws + sub * CH);
    const float4* gb4 = reinterpret_cast<const float4*>(gbs + sub * CH);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
      const float4 w0 = gw4[2 * j], w1 = gw4[2 * j + 1], b0 = gb4[2 * j], b1 = gb4[2 * j + 1];
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = (x2_lo<XF16>(u[k]) - mean) * rstd * wv[2 * k] + bb[2 * k];
        const float hi = (x2_hi<XF16>(u[k]) - mean) * rstd * wv[2 * k + 1] + bb[2 * k + 1];
        __nv_bfloat162 ob = __floats2bfloat162_rn(lo, hi);
        o[k] = *reinterpret_cast<uint32_t*>(&ob);
      }
      dst[j] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// fp32 rows (the 1e-4 mode): the same scheme with NV float4 per thread (C = 4 * NV * TPP), thread-local two-pass statistics
template <int NV, int TPP>
__global__ void __launch_bounds__(256)
lnpatch_tpp_f32_kernel(const float* __restrict__ x, int64_t B, int H, int W, int Ho, int Wo, const float* __restrict__ ln_w,
                       const float* __restrict__ ln_b, float* __restrict__ out) {
  constexpr int CH = 4 * NV, C = CH * TPP;
  __shared__ __align__(16) float gws[C], gbs[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { gws[i] = __ldg(ln_w + i); gbs[i] = __ldg(ln_b + i); }
  __syncthreads();
  const int Hu = 2 * Ho, Wu = 2 * Wo;
  const int64_t total = B * (int64_t)Hu * Wu * TPP;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int sub = (int)(i % TPP);
    const int64_t p = i / TPP;
    const int ix = (int)(p % Wu);
    const int64_t t = p / Wu;
    const int iy = (int)(t % Hu);
    const int64_t b = t / Hu;
    const float4* src = reinterpret_cast<const float4*>(x + ((b * H + iy) * (int64_t)W + ix) * C + sub * CH);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = __ldg(src + j);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / (float)C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
      q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
    }
#pragma unroll
    for (int o = TPP / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / (float)C) + kLnEps);
    const int oy = iy >> 1, dy = iy & 1, ox = ix >> 1, dx = ix & 1;
    float4* dst = reinterpret_cast<float4*>(out + ((b * Ho + oy) * (int64_t)Wo + ox) * (4 * (int64_t)C) + (dy * 2 + dx) * C + sub * CH);
    const float4* gw4 = reinterpret_cast<const float4*>(gws + sub * CH);
    const float4* gb4 = reinterpret_cast<const float4*>(gbs + sub * CH);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 w = gw4[j], bb = gb4[j];
      dst[j] = make_float4((v[j].x - mean) * rstd * w.x + bb.x, (v[j].y - mean) * rstd * w.y + bb.y,
                           (v[j].z - mean) * rstd * w.z + bb.z, (v[j].w - mean) * rstd * w.w + bb.w);
    }
  }
}

template <int NV, int TPP>
static void launch_lnpatch_tpp_f32(const float* xi, int64_t B, int H, int W, int Ho, int Wo, const float* ln_w,
                                   const float* ln_b, float* xo, cudaStream_t st) {
  const int64_t total = B * 4 * (int64_t)Ho * Wo * TPP;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  lnpatch_tpp_f32_kernel<NV, TPP><<<(unsigned)grid, 256, 0, st>>>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo);
}

template <int NV, int TPP>
static void launch_lnpatch_tpp(const __nv_bfloat16* xi, int64_t B, int H, int W, int Ho, int Wo, const float* ln_w,
                               const float* ln_b, __nv_bfloat16* xo, bool xf16, cudaStream_t st) {
  const int64_t total = B * 4 * (int64_t)Ho * Wo * TPP;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (xf16) lnpatch_tpp_kernel<NV, TPP, true><<<(unsigned)grid, 256, 0, st>>>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo);
  else lnpatch_tpp_kernel<NV, TPP, false><<<(unsigned)grid, 256, 0, st>>>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo);
}

extern "C" int btsb_convnext_lnpatch_fwd(const void* x, int dtype, int64_t B, int H, int W, int C,
                                         const float* ln_w, const float* ln_b, void* out, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && H >= 2 && W >= 2 && C >= 1 && C <= 640, "lnpatch: bad shape (C<=640, H,W>=2)");
  BTSB_REQUIRE(dtype == BTSB_F32 || dtype == BTSB_BF16 || dtype == BTSB_BF16_XF16, "lnpatch: dtype must be F32, BF16 or BF16_XF16");
  const bool xf16 = dtype == BTSB_BF16_XF16;       // x: fp16 residual-stream rows; out (the downsample GEMM operand): bf16
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && ln_w && ln_b && out, "lnpatch: null pointer");
  const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
  const int64_t total = B * 4 * (int64_t)Ho * Wo;
  const int grid = pick_grid(total, 8 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_F32) {
    const float* xi = (const float*)x;
    float* xo = (float*)out;
    const bool al = ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0;
    if (al && C == 64) launch_lnpatch_tpp_f32<8, 2>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else if (al && C == 80) launch_lnpatch_tpp_f32<10, 2>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else if (al && C == 128) launch_lnpatch_tpp_f32<8, 4>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else if (al && C == 160) launch_lnpatch_tpp_f32<10, 4>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else if (al && C == 256) launch_lnpatch_tpp_f32<8, 8>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else if (al && C == 320) launch_lnpatch_tpp_f32<10, 8>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, st);
    else lnpatch_kernel<float><<<grid, 256, 0, st>>>(xi, B, H, W, C, Ho, Wo, ln_w, ln_b, xo);
  } else if ((C == 64 || C == 80 || C == 128 || C == 160 || C == 256 || C == 320) && ((uintptr_t)x % 16) == 0 &&
           ((uintptr_t)out % 16) == 0) {
    const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
    __nv_bfloat16* xo = (__nv_bfloat16*)out;
    switch (C) {
      case 64: launch_lnpatch_tpp<8, 1>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
      case 80: launch_lnpatch_tpp<10, 1>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
      case 128: launch_lnpatch_tpp<8, 2>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
      case 160: launch_lnpatch_tpp<10, 2>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
      case 256: launch_lnpatch_tpp<8, 4>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
      default: launch_lnpatch_tpp<10, 4>(xi, B, H, W, Ho, Wo, ln_w, ln_b, xo, xf16, st); break;
    }
  } else if (xf16) {
    lnpatch_kernel<__half, __nv_bfloat16><<<grid, 256, 0, st>>>((const __half*)x, B, H, W, C, Ho, Wo, ln_w, ln_b,
                                                                 (__nv_bfloat16*)out);
  } else if (C % 2 == 0 && ((uintptr_t)ln_w % 8) == 0 && ((uintptr_t)ln_b % 8) == 0) {
    const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
    __nv_bfloat16* xo = (__nv_bfloat16*)out;
    if (C <= 128) lnpatch_bf16x2_kernel<2><<<grid, 256, 0, st>>>(xi, B, H, W, C, Ho, Wo, ln_w, ln_b, xo);
    else if (C <= 192) lnpatch_bf16x2_kernel<3><<<grid, 256, 0, st>>>(xi, B, H, W, C, Ho, Wo, ln_w, ln_b, xo);
    else if (C <= 320) lnpatch_bf16x2_kernel<5><<<grid, 256, 0, st>>>(xi, B, H, W, C, Ho, Wo, ln_w, ln_b, xo);
    else lnpatch_bf16x2_kernel<10><<<grid, 256, 0, st>>>(xi, B, H, W, C, Ho, Wo, ln_w, ln_b, xo);
  } else
    lnpatch_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, H, W, C, Ho, Wo, ln_w, ln_b,
                                                        (__nv_bfloat16*)out);
  return launch_done("lnpatch");
}

extern "C" int btsb_convnext_poolln_fwd(const void* x, int dtype, int64_t B, int HW, int C, const float* ln_w,
                                        const float* ln_b, float* out, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && HW >= 1 && C >= 1 && C <= 640, "poolln: bad shape (C<=640)");
  BTSB_REQUIRE(dtype == BTSB_F32 || dtype == BTSB_BF16, "poolln: dtype must be F32 or BF16");
  BTSB_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "poolln: ln_w and ln_b must both be set or both NULL");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && out, "poolln: null pointer");
  const int grid = pick_grid(B, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_F32)
    poolln_kernel<float><<<grid, 256, 0, st>>>((const float*)x, B, HW, C, ln_w, ln_b, out);
  else
    poolln_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, HW, C, ln_w, ln_b, out);
  return launch_done("poolln");
}

namespace btsb {
int gemm_f32(const float* A, const float* Wt, const float* bias, const float* gamma, const float* res, float* out,
             int64_t M, int N, int K, int epilogue, cudaStream_t st) {
  BTSB_REQUIRE(K % 4 == 0, "gemm f32: K=%d must be a multiple of 4", K);
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)((N + SG_BN - 1) / SG_BN));
  if (epilogue == BTSB_EPI_BIAS)
    sgemm_kernel<BTSB_EPI_BIAS><<<grid, 256, 0, st>>>(A, Wt, bias, gamma, res, out, M, N, K);
  else if (epilogue == BTSB_EPI_BIAS_GELU)
    sgemm_kernel<BTSB_EPI_BIAS_GELU><<<grid, 256, 0, st>>>(A, Wt, bias, gamma, res, out, M, N, K);
  else if (epilogue == BTSB_EPI_BIAS_SILU)
    sgemm_kernel<BTSB_EPI_BIAS_SILU><<<grid, 256, 0, st>>>(A, Wt, bias, gamma, res, out, M, N, K);
  else
    sgemm_kernel<BTSB_EPI_SCALE_RES><<<grid, 256, 0, st>>>(A, Wt, bias, gamma, res, out, M, N, K);
  return launch_done("gemm_f32");
}
}  // namespace btsb

extern "C" int btsb_score_epilogue(const float* logits, int64_t B, float* scores, uint8_t* labels, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0, "score: B < 0");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(logits, "score: null logits");
  score_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, B, scores, labels);
  return launch_done("score");
}

extern "C" int btsb_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(in && out, "cast: null pointer");
  cast_f32_bf16_kernel<<<pick_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n);
  return launch_done("cast_f32_bf16");
}
extern "C" int btsb_cast_bf16_to_f32(const void* in, float* out, int64_t n, void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(in && out, "cast: null pointer");
  cast_bf16_f32_kernel<<<pick_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, out, n);
  return launch_done("cast_bf16_f32");
}
