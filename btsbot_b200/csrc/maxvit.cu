// CUDA-core kernels of the MaxViT path (timm maxvit_tiny_rw_224 behind btsbot/architectures.py:25-101): everything
// around the tcgen05 GEMMs -- fused bilinear-resize + stem conv, 3x3 im2col, MBConv depthwise 3x3 + BN + SiLU with the
// squeeze-excitation pooling, SE gate, gate scaling, row LayerNorm, 7x7 window / grid attention with relative-position
// bias, final LayerNorm + average pool.  Activations are NHWC pixel rows [B*H*W, C] (float32 or bf16, fp32 math).
// Contracts and reference citations: include/btsbot_b200.h.
#include <stdlib.h>

#include "common.cuh"

namespace btsb {
namespace {

template <typename T> struct Vec2;
template <> struct Vec2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
};
template <> struct Vec2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
};

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

int grid_for(int64_t items, int per_block, int cap = 148 * 32) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// =====================================================================================================
// M0  bilinear resize (align_corners=False) fused into stem.conv1 (3x3, stride 2, pad 1, no bias) + BN + SiLU
//   x [B,3,Hin,Win] fp32 NCHW -> out rows [B*Ho*Wo, C1], Ho = Wo = S/2.  The S x S resized image is never
//   materialised: a CTA computes the (2*16+1)^2 x 3 patch it needs into shared memory, then thread = output pixel.
//   w: [27][C1] fp32 with BatchNorm scale folded in (k = (ci*3+ky)*3+kx); shift: [C1].
// =====================================================================================================
constexpr int kS1Tile = 16;
constexpr int kS1Patch = 2 * kS1Tile + 1;   // 33
constexpr int kS1MaxC = 32;

template <typename TO>
__global__ void __launch_bounds__(kS1Tile* kS1Tile)
mv_stem1_kernel(const float* __restrict__ x, int Hin, int Win, int S, int Ho, int Wo, const float* __restrict__ w,
                const float* __restrict__ shift, int C1, TO* __restrict__ out, float sy, float sx) {
  __shared__ float patch[3][kS1Patch][kS1Patch + 1];
  __shared__ __align__(16) float ws[27 * kS1MaxC];
  __shared__ float sh[kS1MaxC];
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.z;
  const int oy0 = blockIdx.y * kS1Tile, ox0 = blockIdx.x * kS1Tile;
  for (int i = tid; i < 27 * C1; i += blockDim.x) ws[i] = w[i];
  if (tid < C1) sh[tid] = shift[tid];
  // resized-image coordinates covered: rows 2*oy0-1 .. 2*oy0+31, cols likewise; outside [0,S) is the conv's zero pad
  const float* xb = x + b * 3 * (int64_t)Hin * Win;
  for (int i = tid; i < 3 * kS1Patch * kS1Patch; i += blockDim.x) {
    const int ci = i / (kS1Patch * kS1Patch);
    const int r = i - ci * kS1Patch * kS1Patch;
    const int py = r / kS1Patch, px = r - py * kS1Patch;
    const int uy = 2 * oy0 - 1 + py, ux = 2 * ox0 - 1 + px;
    float v = 0.f;
    if (uy >= 0 && uy < S && ux >= 0 && ux < S) {
      // torch upsample_bilinear2d, align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0
      float fy = fmaxf(sy * ((float)uy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ux + 0.5f) - 0.5f, 0.f);
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float* pc = xb + ci * (int64_t)Hin * Win;
      const float p00 = __ldg(pc + y0 * Win + x0), p01 = __ldg(pc + y0 * Win + x1);
      const float p10 = __ldg(pc + y1 * Win + x0), p11 = __ldg(pc + y1 * Win + x1);
      v = (1.f - ly) * ((1.f - lx) * p00 + lx * p01) + ly * ((1.f - lx) * p10 + lx * p11);
    }
    patch[ci][py][px] = v;
  }
  __syncthreads();
  const int ty = tid / kS1Tile, tx = tid - ty * kS1Tile;
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy >= Ho || ox >= Wo) return;
  float acc[kS1MaxC];
#pragma unroll
  for (int c = 0; c < kS1MaxC; ++c) acc[c] = 0.f;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float v = patch[ci][2 * ty + ky][2 * tx + kx];
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ci * 3 + ky) * 3 + kx) * C1);
#pragma unroll
        for (int c4 = 0; c4 < kS1MaxC / 4; ++c4) {
          if (c4 * 4 < C1) {
            const float4 w4 = wr[c4];
            acc[c4 * 4] = fmaf(v, w4.x, acc[c4 * 4]); acc[c4 * 4 + 1] = fmaf(v, w4.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(v, w4.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(v, w4.w, acc[c4 * 4 + 3]);
          }
        }
      }
  TO* dst = out + ((b * Ho + oy) * (int64_t)Wo + ox) * C1;
#pragma unroll
  for (int c = 0; c < kS1MaxC; c += 2)
    if (c < C1) Vec2<TO>::st(dst + c, silu_exact(acc[c] + sh[c]), silu_exact(acc[c + 1] + sh[c + 1]));
}

// =====================================================================================================
// M1  3x3 / stride 1 / pad 1 im2col: x [B,H,W,C] -> out [B*H*W, 9C], column (ky*3+kx)*C + c (zero padded)
//   one thread = one 16-byte piece of one (pixel, tap)
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
mv_im2col3_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t B, int H, int W, int C) {
  constexpr int V = 16 / (int)sizeof(T);
  const int cv = C / V;
  const int64_t total = B * H * W * 9 * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    int64_t t = i / cv;
    const int tap = (int)(t % 9); t /= 9;
    const int xw = (int)(t % W); t /= W;
    const int yh = (int)(t % H);
    const int64_t b = t / H;
    const int iy = yh + tap / 3 - 1, ix = xw + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + ((b * H + iy) * (int64_t)W + ix) * C) + c);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// =====================================================================================================
// M2  2x2 average pool (MBConv shortcut): [B,H,W,C] -> [B,H/2,W/2,C]
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
mv_avgpool2_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, c2 = C / 2;
  const int64_t total = B * Ho * Wo * c2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c2) * 2;
    int64_t t = i / c2;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int64_t b = t / Ho;
    const T* p = x + ((b * H + 2 * oy) * (int64_t)W + 2 * ox) * C + c;
    const float2 a = Vec2<T>::ld(p), bq = Vec2<T>::ld(p + C), cq = Vec2<T>::ld(p + (int64_t)W * C),
                 d = Vec2<T>::ld(p + (int64_t)W * C + C);
    Vec2<T>::st(out + ((b * Ho + oy) * (int64_t)Wo + ox) * C + c, 0.25f * (a.x + bq.x + cq.x + d.x),
                0.25f * (a.y + bq.y + cq.y + d.y));
  }
}

// =====================================================================================================
// M3  MBConv depthwise 3x3 (stride 1|2, pad 1, no bias) + BatchNorm (folded) + SiLU, plus the SE squeeze:
//   out [B,Ho,Wo,C] and pooled[b,c] = mean over (Ho,Wo) of out -- summed in a fixed order (deterministic).
//   CTA = (64-channel slice, image); lane = channel pair, warp walks output pixels.  w: [9][C] (BN scale folded).
// =====================================================================================================
constexpr int kDwWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kDwWarps * 32)
mv_dw3_kernel(const T* __restrict__ x, int H, int W, int C, int stride, int Ho, int Wo, const float* __restrict__ w,
              const float* __restrict__ shift, T* __restrict__ out, float* __restrict__ pooled) {
  __shared__ float red[kDwWarps][64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + lane * 2;
  const int64_t b = blockIdx.y;
  const bool live = c < C;
  float wa[9], wb[9], sa = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { wa[k] = live ? w[k * C + c] : 0.f; wb[k] = live ? w[k * C + c + 1] : 0.f; }
  if (live) { sa = shift[c]; sb = shift[c + 1]; }
  const T* xb = x + b * (int64_t)H * W * C;
  T* ob = out + b * (int64_t)Ho * Wo * C;
  float pa = 0.f, pb = 0.f;
  if (live) {
    for (int p = wid; p < Ho * Wo; p += kDwWarps) {
      const int oy = p / Wo, ox = p - oy * Wo;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * stride + kx - 1;
          if (ix < 0 || ix >= W) continue;
          const float2 v = Vec2<T>::ld(xb + ((int64_t)iy * W + ix) * C + c);
          a0 = fmaf(v.x, wa[ky * 3 + kx], a0);
          a1 = fmaf(v.y, wb[ky * 3 + kx], a1);
        }
      }
      a0 = silu_exact(a0 + sa); a1 = silu_exact(a1 + sb);
      Vec2<T>::st(ob + (int64_t)p * C + c, a0, a1);
      pa += a0; pb += a1;
    }
  }
  red[wid][lane * 2] = pa; red[wid][lane * 2 + 1] = pb;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kDwWarps; ++k) s += red[k][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) pooled[b * C + cc] = s / (float)(Ho * Wo);
  }
}

// =====================================================================================================
// M4  SE gate: g[b,c] = sigmoid(W2 . silu(W1 . pooled[b] + b1) + b2), one CTA per image.
//   w1 [R][C], w2 TRANSPOSED to [R][C] fp32 (so both phases read coalesced); R <= 128.
// =====================================================================================================
__global__ void __launch_bounds__(256)
mv_se_kernel(const float* __restrict__ pooled, int C, int R, const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float sm[];
  float* pin = sm;          // [C]
  float* hid = sm + C;      // [R]
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C; i += blockDim.x) pin[i] = pooled[b * C + i];
  __syncthreads();
  for (int r = wid; r < R; r += 8) {
    float s = 0.f;
    for (int i = lane; i < C; i += 32) s = fmaf(w1[(int64_t)r * C + i], pin[i], s);
    s = warp_sum(s);
    if (lane == 0) hid[r] = silu_exact(s + b1[r]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int r = 0; r < R; ++r) s = fmaf(w2[(int64_t)r * C + c], hid[r], s);      // w2 is [R][C]: coalesced over c
    gate[b * C + c] = 1.0f / (1.0f + expf(-s));
  }
}

// =====================================================================================================
// M5  SE excite: x[b, p, c] *= gate[b, c]   (in place)
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
mv_scale_kernel(T* __restrict__ x, const float* __restrict__ gate, int64_t B, int HW, int C) {
  const int c2 = C / 2;
  const int64_t total = B * HW * c2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c2) * 2;
    const int64_t b = i / ((int64_t)HW * c2);
    const float2 v = Vec2<T>::ld(x + i * 2);
    const float2 g = *reinterpret_cast<const float2*>(gate + b * C + c);
    Vec2<T>::st(x + i * 2, v.x * g.x, v.y * g.y);
  }
}

// =====================================================================================================
// M6  row LayerNorm (eps 1e-6, biased variance, two-pass fp32): x [M,C] -> out [M,C]; warp per row, C <= 512
// =====================================================================================================
constexpr int kLnMaxPairs = 8;   // C <= 512: 32 lanes x 8 pairs x 2

template <typename T>
__global__ void __launch_bounds__(256)
mv_ln_kernel(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta, T* __restrict__ out,
             int64_t M, int C) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pairs = C / 64;       // pairs per lane (C % 64 == 0)
  for (int64_t row = (int64_t)blockIdx.x * 8 + wid; row < M; row += (int64_t)gridDim.x * 8) {
    const T* xr = x + row * C;
    float2 v[kLnMaxPairs];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) { v[j] = Vec2<T>::ld(xr + (j * 32 + lane) * 2); s += v[j].x + v[j].y; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) { const float d0 = v[j].x - mean, d1 = v[j].y - mean; q += d0 * d0 + d1 * d1; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
    T* orow = out + row * C;
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) {
        const int c = (j * 32 + lane) * 2;
        const float2 gg = *reinterpret_cast<const float2*>(g + c), bb = *reinterpret_cast<const float2*>(bta + c);
        Vec2<T>::st(orow + c, (v[j].x - mean) * rstd * gg.x + bb.x, (v[j].y - mean) * rstd * gg.y + bb.y);
      }
  }
}

// =====================================================================================================
// M7  7x7 window ("block") / grid attention, one CTA per (window, head): 49 tokens x 32 dims.
//   qkv rows [B*H*W, 3C] in IMAGE order, head h owns columns [96h, 96h+96) = q | k | v (timm head_first);
//   out rows [B*H*W, C], head h -> columns [32h, 32h+32).  The window/grid partition is index arithmetic:
//   token (ty,tx) of window (wy,wx):  block: (wy*7+ty, wx*7+tx)   grid: (ty*(H/7)+wy, tx*(W/7)+wx).
//   S = (q*scale) k^T + table[rel(i,j), h]; softmax; O = P v.   thread = query token, K/V in shared memory.
// =====================================================================================================
constexpr int kWin = 7, kTok = 49, kDh = 32;

template <typename T>
__global__ void __launch_bounds__(64)
mv_attn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int H, int W, int C, int heads, int grid_mode,
               const float* __restrict__ table /* [169, heads] */, float scale) {
  __shared__ __align__(16) float ks[kTok][kDh];
  __shared__ __align__(16) float vs[kTok][kDh];
  __shared__ float tb[169];
  const int h = blockIdx.y;
  const int nwx = W / kWin, nwy = H / kWin;
  int64_t win = blockIdx.x;
  const int wx = (int)(win % nwx); win /= nwx;
  const int wy = (int)(win % nwy);
  const int64_t b = win / nwy;
  const int t = threadIdx.x;
  for (int i = t; i < 169; i += 64) tb[i] = table[i * heads + h];
  const int ty = t / kWin, tx = t - ty * kWin;
  int64_t row = 0;
  float q[kDh];
  if (t < kTok) {
    const int y = grid_mode ? ty * nwy + wy : wy * kWin + ty;
    const int xq = grid_mode ? tx * nwx + wx : wx * kWin + tx;
    row = (b * H + y) * (int64_t)W + xq;
    const T* p = qkv + row * 3 * C + h * 3 * kDh;
#pragma unroll
    for (int d = 0; d < kDh; d += 2) {
      const float2 a = Vec2<T>::ld(p + d), k2 = Vec2<T>::ld(p + kDh + d), v2 = Vec2<T>::ld(p + 2 * kDh + d);
      q[d] = a.x * scale; q[d + 1] = a.y * scale;
      ks[t][d] = k2.x; ks[t][d + 1] = k2.y;
      vs[t][d] = v2.x; vs[t][d + 1] = v2.y;
    }
  }
  __syncthreads();
  if (t >= kTok) return;
  float s[kTok];
  float mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(ks[j]);
    float a = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < kDh / 4; ++d4) {
      const float4 k4 = kr[d4];
      a = fmaf(q[d4 * 4], k4.x, a); a = fmaf(q[d4 * 4 + 1], k4.y, a);
      a = fmaf(q[d4 * 4 + 2], k4.z, a); a = fmaf(q[d4 * 4 + 3], k4.w, a);
    }
    const int jy = j / kWin, jx = j - jy * kWin;
    a += tb[(ty - jy + kWin - 1) * (2 * kWin - 1) + (tx - jx + kWin - 1)];
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) { s[j] = __expf(s[j] - mx); den += s[j]; }
  const float inv = 1.0f / den;
  float o[kDh];
#pragma unroll
  for (int d = 0; d < kDh; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    const float4* vr = reinterpret_cast<const float4*>(vs[j]);
    const float pj = s[j];
#pragma unroll
    for (int d4 = 0; d4 < kDh / 4; ++d4) {
      const float4 v4 = vr[d4];
      o[d4 * 4] = fmaf(pj, v4.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(pj, v4.y, o[d4 * 4 + 1]);
      o[d4 * 4 + 2] = fmaf(pj, v4.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(pj, v4.w, o[d4 * 4 + 3]);
    }
  }
  T* op = out + row * C + h * kDh;
#pragma unroll
  for (int d = 0; d < kDh; d += 2) Vec2<T>::st(op + d, o[d] * inv, o[d + 1] * inv);
}

// =====================================================================================================
// M8  final LayerNorm2d + global average pool: x [B*HW, C] -> out [B, C] fp32 (norm first, then mean -- timm
//   MaxxVit.norm followed by head.global_pool).  One CTA per image, warp per row, fixed-order reduction.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
mv_lnpool_kernel(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                 float* __restrict__ out, int HW, int C) {
  extern __shared__ float red[];     // [8][C]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t b = blockIdx.x;
  const int pairs = C / 64;
  float2 acc[kLnMaxPairs];
#pragma unroll
  for (int j = 0; j < kLnMaxPairs; ++j) acc[j] = make_float2(0.f, 0.f);
  for (int r = wid; r < HW; r += 8) {
    const T* xr = x + (b * HW + r) * C;
    float2 v[kLnMaxPairs];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) { v[j] = Vec2<T>::ld(xr + (j * 32 + lane) * 2); s += v[j].x + v[j].y; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) { const float d0 = v[j].x - mean, d1 = v[j].y - mean; q += d0 * d0 + d1 * d1; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + kLnEps);
#pragma unroll
    for (int j = 0; j < kLnMaxPairs; ++j)
      if (j < pairs) { acc[j].x += (v[j].x - mean) * rstd; acc[j].y += (v[j].y - mean) * rstd; }
  }
#pragma unroll
  for (int j = 0; j < kLnMaxPairs; ++j)
    if (j < pairs) { red[wid * C + (j * 32 + lane) * 2] = acc[j].x; red[wid * C + (j * 32 + lane) * 2 + 1] = acc[j].y; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k * C + c];
    out[b * C + c] = s / (float)HW * g[c] + bta[c];     // the affine commutes with the mean
  }
}

}  // namespace
}  // namespace btsb

namespace btsb {
int maxvit_dw3_bf16(const void* x, int64_t B, int H, int W, int C, int stride, int Ho, int Wo, const float* w,
                    const float* shift, void* out, float* pooled, cudaStream_t st);
int maxvit_ln_bf16(const void* x, const float* g, const float* b, void* out, int64_t M, int C, cudaStream_t st);
int maxvit_scale_bf16(void* x, const float* gate, int64_t B, int HW, int C, cudaStream_t st);
int maxvit_avgpool2_bf16(const void* x, void* out, int64_t B, int H, int W, int C, cudaStream_t st);
int maxvit_attn_bf16_tc(const void* qkv, void* out, int64_t B, int H, int W, int C, int grid_mode, const float* table,
                        cudaStream_t st);
}
using namespace btsb;

#define MV_DISPATCH(dtype, CALL_F32, CALL_BF16) \
  do {                                          \
    if ((dtype) == BTSB_F32) { CALL_F32; }      \
    else { CALL_BF16; }                         \
  } while (0)

static int mv_check_dtype(int dtype, const char* what) {
  BTSB_REQUIRE(dtype == BTSB_F32 || dtype == BTSB_BF16, "%s: dtype must be F32 or BF16", what);
  return BTSB_OK;
}

extern "C" int btsb_maxvit_stem1_fwd(const float* x, int64_t B, int Hin, int Win, int S, const float* w,
                                     const float* shift, int C1, void* out, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit stem1")) return e;
  BTSB_REQUIRE(B >= 0 && Hin >= 1 && Win >= 1 && S >= 2 && S % 2 == 0, "maxvit stem1: bad shape");
  BTSB_REQUIRE(C1 >= 4 && C1 <= kS1MaxC && C1 % 4 == 0, "maxvit stem1: C1=%d must be a multiple of 4, <= %d", C1, kS1MaxC);
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && w && shift && out, "maxvit stem1: null pointer");
  BTSB_REQUIRE(B <= 65535, "maxvit stem1: at most 65535 images per call");
  const int Ho = S / 2, Wo = S / 2;
  dim3 grid((Wo + kS1Tile - 1) / kS1Tile, (Ho + kS1Tile - 1) / kS1Tile, (unsigned)B);
  // torch area_pixel_compute_scale (align_corners=False): input_size / output_size in float
  const float sy = (float)Hin / (float)S, sx = (float)Win / (float)S;
  cudaStream_t st = (cudaStream_t)stream;
  MV_DISPATCH(dtype,
              (mv_stem1_kernel<float><<<grid, kS1Tile * kS1Tile, 0, st>>>(x, Hin, Win, S, Ho, Wo, w, shift, C1, (float*)out, sy, sx)),
              (mv_stem1_kernel<__nv_bfloat16><<<grid, kS1Tile * kS1Tile, 0, st>>>(x, Hin, Win, S, Ho, Wo, w, shift, C1,
                                                                                 (__nv_bfloat16*)out, sy, sx)));
  return launch_done("maxvit_stem1");
}

extern "C" int btsb_maxvit_im2col3_fwd(const void* x, void* out, int64_t B, int H, int W, int C, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit im2col3")) return e;
  BTSB_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 1, "maxvit im2col3: bad shape");
  BTSB_REQUIRE((C * (dtype == BTSB_F32 ? 4 : 2)) % 16 == 0, "maxvit im2col3: a pixel's channels must be a multiple of 16 bytes");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && out && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0, "maxvit im2col3: null/unaligned pointer");
  const int64_t total = B * H * W * 9 * (C * (dtype == BTSB_F32 ? 4 : 2) / 16);
  const int grid = grid_for(total, 256 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  MV_DISPATCH(dtype, (mv_im2col3_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, B, H, W, C)),
              (mv_im2col3_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, B, H, W, C)));
  return launch_done("maxvit_im2col3");
}

extern "C" int btsb_maxvit_avgpool2_fwd(const void* x, void* out, int64_t B, int H, int W, int C, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit avgpool2")) return e;
  BTSB_REQUIRE(B >= 0 && H >= 2 && W >= 2 && C >= 2 && C % 2 == 0, "maxvit avgpool2: bad shape");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && out, "maxvit avgpool2: null pointer");
  const int64_t total = B * (H / 2) * (W / 2) * (C / 2);
  const int grid = grid_for(total, 256 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_BF16) { const int r = maxvit_avgpool2_bf16(x, out, B, H, W, C, st); if (r != 1) return r; }
  MV_DISPATCH(dtype, (mv_avgpool2_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, B, H, W, C)),
              (mv_avgpool2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, B, H, W, C)));
  return launch_done("maxvit_avgpool2");
}

extern "C" int btsb_maxvit_dw3_fwd(const void* x, int64_t B, int H, int W, int C, int stride, const float* w,
                                   const float* shift, void* out, float* pooled, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit dw3")) return e;
  BTSB_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 2 && C % 2 == 0 && (stride == 1 || stride == 2), "maxvit dw3: bad shape");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && w && shift && out && pooled, "maxvit dw3: null pointer");
  BTSB_REQUIRE(B <= 65535, "maxvit dw3: at most 65535 images per call");
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  dim3 grid((C + 63) / 64, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_BF16) { const int r = maxvit_dw3_bf16(x, B, H, W, C, stride, Ho, Wo, w, shift, out, pooled, st); if (r != 1) return r; }
  MV_DISPATCH(dtype,
              (mv_dw3_kernel<float><<<grid, kDwWarps * 32, 0, st>>>((const float*)x, H, W, C, stride, Ho, Wo, w, shift, (float*)out, pooled)),
              (mv_dw3_kernel<__nv_bfloat16><<<grid, kDwWarps * 32, 0, st>>>((const __nv_bfloat16*)x, H, W, C, stride, Ho, Wo, w, shift,
                                                                           (__nv_bfloat16*)out, pooled)));
  return launch_done("maxvit_dw3");
}

extern "C" int btsb_maxvit_se_fwd(const float* pooled, int64_t B, int C, int R, const float* w1, const float* b1,
                                  const float* w2, const float* b2, float* gate, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && C >= 1 && R >= 1 && (C + R) * 4 <= 48 * 1024, "maxvit se: bad shape");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(pooled && w1 && b1 && w2 && b2 && gate, "maxvit se: null pointer");
  mv_se_kernel<<<(unsigned)B, 256, (C + R) * 4, (cudaStream_t)stream>>>(pooled, C, R, w1, b1, w2, b2, gate);
  return launch_done("maxvit_se");
}

extern "C" int btsb_maxvit_scale_fwd(void* x, const float* gate, int64_t B, int HW, int C, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit scale")) return e;
  BTSB_REQUIRE(B >= 0 && HW >= 1 && C >= 2 && C % 2 == 0, "maxvit scale: bad shape");
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && gate, "maxvit scale: null pointer");
  const int grid = grid_for(B * HW * (C / 2), 256 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_BF16) { const int r = maxvit_scale_bf16(x, gate, B, HW, C, st); if (r != 1) return r; }
  MV_DISPATCH(dtype, (mv_scale_kernel<float><<<grid, 256, 0, st>>>((float*)x, gate, B, HW, C)),
              (mv_scale_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)x, gate, B, HW, C)));
  return launch_done("maxvit_scale");
}

extern "C" int btsb_layernorm_rows_fwd(const void* x, const float* ln_w, const float* ln_b, void* out, int64_t M, int C,
                                       int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "layernorm rows")) return e;
  BTSB_REQUIRE(M >= 0 && C >= 64 && C % 64 == 0 && C <= 64 * kLnMaxPairs, "layernorm rows: C=%d must be a multiple of 64, <= 512", C);
  if (M == 0) return BTSB_OK;
  BTSB_REQUIRE(x && ln_w && ln_b && out, "layernorm rows: null pointer");
  const int grid = grid_for(M, 8 * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == BTSB_BF16) { const int r = maxvit_ln_bf16(x, ln_w, ln_b, out, M, C, st); if (r != 1) return r; }
  MV_DISPATCH(dtype, (mv_ln_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ln_w, ln_b, (float*)out, M, C)),
              (mv_ln_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ln_w, ln_b, (__nv_bfloat16*)out, M, C)));
  return launch_done("layernorm_rows");
}

extern "C" int btsb_maxvit_attn_fwd(const void* qkv, void* out, int64_t B, int H, int W, int C, int grid_mode,
                                    const float* table, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit attn")) return e;
  BTSB_REQUIRE(B >= 0 && H >= kWin && W >= kWin && H % kWin == 0 && W % kWin == 0, "maxvit attn: H=%d W=%d must be multiples of 7", H, W);
  BTSB_REQUIRE(C >= kDh && C % kDh == 0, "maxvit attn: C=%d must be a multiple of dim_head 32", C);
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(qkv && out && table, "maxvit attn: null pointer");
  const int64_t nwin = B * (H / kWin) * (W / kWin);
  BTSB_REQUIRE(nwin < (1ll << 31), "maxvit attn: too many windows");
  cudaStream_t st = (cudaStream_t)stream;
  // bf16: tcgen05 / TMEM kernel (maxvit_attn_tc.cu).  It replaced an mma.sync kernel (one warp per (window, head)) at equal
  // speed -- 0.72 vs 0.68 ms at C = 64, 0.36 vs 0.36 at 128, 0.19 vs 0.20 at 256, 0.105 vs 0.112 at 512 per 1024 images
  // (profiles/r02j): both sit on the row gather at ~2.2 TB/s, not on the tensor pipe.
  if (dtype == BTSB_BF16) return maxvit_attn_bf16_tc(qkv, out, B, H, W, C, grid_mode, table, st);
  dim3 grid((unsigned)nwin, C / kDh);
  const float scale = 0.17677669529663687f;   // dim_head ** -0.5
  MV_DISPATCH(dtype,
              (mv_attn_kernel<float><<<grid, 64, 0, st>>>((const float*)qkv, (float*)out, H, W, C, C / kDh, grid_mode, table, scale)),
              (mv_attn_kernel<__nv_bfloat16><<<grid, 64, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, H, W, C, C / kDh,
                                                                 grid_mode, table, scale)));
  return launch_done("maxvit_attn");
}

extern "C" int btsb_maxvit_lnpool_fwd(const void* x, const float* ln_w, const float* ln_b, float* out, int64_t B, int HW,
                                      int C, int dtype, void* stream) {
  if (int e = check_device()) return e;
  if (int e = mv_check_dtype(dtype, "maxvit lnpool")) return e;
  BTSB_REQUIRE(B >= 0 && HW >= 1 && C >= 64 && C % 64 == 0 && C <= 64 * kLnMaxPairs, "maxvit lnpool: C=%d must be a multiple of 64, <= 512", C);
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && ln_w && ln_b && out, "maxvit lnpool: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  MV_DISPATCH(dtype, (mv_lnpool_kernel<float><<<(unsigned)B, 256, 8 * C * 4, st>>>((const float*)x, ln_w, ln_b, out, HW, C)),
              (mv_lnpool_kernel<__nv_bfloat16><<<(unsigned)B, 256, 8 * C * 4, st>>>((const __nv_bfloat16*)x, ln_w, ln_b, out, HW, C)));
  return launch_done("maxvit_lnpool");
}
