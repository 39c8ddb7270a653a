// K7 -- fp32 training kernels (forward pieces that must save intermediates + every backward op + loss + AdamW).
// Replaces what autograd/cuDNN/ATen do inside train.py:496-547 (zero_grad -> forward -> BCEWithLogits(pos_weight)
// -> backward -> AdamW.step) for the ConvNeXt models.  CUDA-core fp32 throughout: this is the reference-numerics
// training path (parity with torch autograd on the CPU oracle); a tcgen05 dgrad/wgrad path is future work.
#include "common.cuh"

namespace btsb {

// ---------------------------------------------------------------------------------------------------------
// generic strided GEMM  C[M,N] (+)= sum_k A(m,k) B(k,n)   (covers NT/NN/TN; split-K with atomics for wgrad shapes)
// ---------------------------------------------------------------------------------------------------------
constexpr int GG_BM = 64, GG_BN = 64, GG_BK = 16;

template <bool ATOMIC>
__global__ void __launch_bounds__(256)
gemm_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ Bm, int64_t sbk,
                    int64_t sbn, float* __restrict__ C, int64_t M, int64_t N, int64_t K, int64_t kchunk, int accumulate) {
  __shared__ float As[GG_BK][GG_BM + 4];
  __shared__ float Bs[GG_BK][GG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GG_BM, n0 = (int64_t)blockIdx.y * GG_BN;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = kbeg; k0 < kend; k0 += GG_BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256;
      int m, k;
      if (sak == 1) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
      const int64_t gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? __ldg(A + gm * sam + gk * sak) : 0.f;
      int n, kb;
      if (sbn == 1) { n = e & 63; kb = e >> 6; } else { kb = e & 15; n = e >> 4; }
      const int64_t gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < kend) ? __ldg(Bm + gkb * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GG_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* c = C + gm * N + gn;
      if (ATOMIC) atomicAdd(c, acc[i][j]);
      else *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

// out[n] (+)= sum_m X[m,n] * (Y ? Y[m,n] : 1)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ out, int64_t M, int N,
              int64_t rows_per_block) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;                      // 8 row lanes
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // four loads in flight per thread
    int64_t m = r0 + sub;
    if (Y) {
      for (; m + 24 < r1; m += 32) {
        s0 = fmaf(X[m * N + n], Y[m * N + n], s0);
        s1 = fmaf(X[(m + 8) * N + n], Y[(m + 8) * N + n], s1);
        s2 = fmaf(X[(m + 16) * N + n], Y[(m + 16) * N + n], s2);
        s3 = fmaf(X[(m + 24) * N + n], Y[(m + 24) * N + n], s3);
      }
      for (; m < r1; m += 8) s0 = fmaf(X[m * N + n], Y[m * N + n], s0);
    } else {
      for (; m + 24 < r1; m += 32) {
        s0 += X[m * N + n]; s1 += X[(m + 8) * N + n]; s2 += X[(m + 16) * N + n]; s3 += X[(m + 24) * N + n];
      }
      for (; m < r1; m += 8) s0 += X[m * N + n];
    }
    s = (s0 + s1) + (s2 + s3);
  }
  __shared__ float red[8][33];
  red[sub][threadIdx.x & 31] = s;
  __syncthreads();
  if (sub == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(out + n, t);
  }
}

__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// dout == NULL: out = act(pre); else out = dout * act'(pre)
__global__ void act_kernel(const float* __restrict__ pre, const float* __restrict__ dout, float* __restrict__ out,
                           int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = pre[i];
    float v;
    if (!dout) v = apply_act(x, act);
    else if (act == BTSB_ACT_GELU) v = dout[i] * gelu_grad(x);
    else if (act == BTSB_ACT_RELU) v = x > 0.f ? dout[i] : 0.f;
    else v = dout[i];
    out[i] = v;
  }
}

// out[m,n] = (res ? res[m,n] : 0) + g[n] * X[m,n]
__global__ void colscale_kernel(const float* __restrict__ X, const float* __restrict__ g, const float* __restrict__ res,
                                float* __restrict__ out, int64_t total, int N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    out[i] = (res ? res[i] : 0.f) + g[n] * X[i];
  }
}
// N % 4 == 0 and 16-byte aligned pointers: one float4 per thread, the column index advances without a modulo per element
__global__ void __launch_bounds__(256)
colscale4_kernel(const float4* __restrict__ X, const float4* __restrict__ g, const float4* __restrict__ res,
                 float4* __restrict__ out, int64_t total4, int N4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    const float4 gv = __ldg(g + (int)(i % N4));
    const float4 x = X[i];
    float4 r = res ? res[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    r.x = fmaf(gv.x, x.x, r.x); r.y = fmaf(gv.y, x.y, r.y); r.z = fmaf(gv.z, x.z, r.z); r.w = fmaf(gv.w, x.w, r.w);
    out[i] = r;
  }
}

// X[m,n] += b[n]
__global__ void bias_add_kernel(float* __restrict__ X, const float* __restrict__ b, int64_t total, int N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    X[i] += b[(int)(i % N)];
}

// ---- LayerNorm over the last dim of rows [M,C]; one warp per row, two-pass -----------------------------------
constexpr int kLnMaxJ = 20;   // C <= 640; the kernels are instantiated for 3 / 5 / 10 / 20 elements per lane so that
                              // C = 80 rows do not execute 17 predicated-off iterations of every loop

template <int LN_MAXJ>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ b,
              float* __restrict__ y, int64_t M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nw) {
    const float* src = u + m * C;
    float v[LN_MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) { const int k = lane + 32 * j; v[j] = k < C ? src[k] : 0.f; s += v[j]; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) if (lane + 32 * j < C) { const float d = v[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) { const int k = lane + 32 * j; if (k < C) y[m * C + k] = (v[j] - mean) * rstd * w[k] + b[k]; }
  }
}

template <int LN_MAXJ>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ dy,
              float* __restrict__ du, float* __restrict__ dw, float* __restrict__ db, int64_t M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float aw[LN_MAXJ], ab[LN_MAXJ], wk[LN_MAXJ];
#pragma unroll
  for (int j = 0; j < LN_MAXJ; ++j) { aw[j] = 0.f; ab[j] = 0.f; const int k = lane + 32 * j; wk[j] = k < C ? w[k] : 0.f; }
  for (int64_t m = warp; m < M; m += nw) {
    const float* src = u + m * C;
    const float* g = dy + m * C;
    float v[LN_MAXJ], gy[LN_MAXJ];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) { const int k = lane + 32 * j; v[j] = k < C ? src[k] : 0.f; gy[j] = k < C ? g[k] : 0.f; s += v[j]; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) if (lane + 32 * j < C) { const float d = v[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; ++j) {
      if (lane + 32 * j < C) {
        const float xh = (v[j] - mean) * rstd;
        const float gw = gy[j] * wk[j];
        c1 += gw; c2 += gw * xh;
        aw[j] += gy[j] * xh; ab[j] += gy[j];
        v[j] = xh; gy[j] = gw;
      }
    }
    c1 = warp_sum(c1) / (float)C; c2 = warp_sum(c2) / (float)C;
    if (du) {
#pragma unroll
      for (int j = 0; j < LN_MAXJ; ++j) { const int k = lane + 32 * j; if (k < C) du[m * C + k] = rstd * (gy[j] - c1 - v[j] * c2); }
    }
  }
#pragma unroll
  for (int j = 0; j < LN_MAXJ; ++j) {
    const int k = lane + 32 * j;
    if (k < C) { atomicAdd(dw + k, aw[j]); atomicAdd(db + k, ab[j]); }
  }
}

// ---- depthwise 7x7, pad 3, NHWC rows; flip=1 correlates with the flipped kernel (= dgrad) --------------------------
__global__ void __launch_bounds__(256)
dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w49, const float* __restrict__ bias,
               float* __restrict__ out, int64_t B, int H, int W, int C, int flip) {
  const int64_t total = B * H * W * (int64_t)C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int ox = (int)(t % W); t /= W;
    const int oy = (int)(t % H);
    const int64_t b = t / H;
    float acc = bias ? bias[c] : 0.f;
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy + ky - 3;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = ox + kx - 3;
        if (ix < 0 || ix >= W) continue;
        const int kk = flip ? (6 - ky) * 7 + (6 - kx) : ky * 7 + kx;
        acc = fmaf(w49[kk * C + c], x[((b * H + iy) * W + ix) * (int64_t)C + c], acc);
      }
    }
    out[i] = acc;
  }
}

// dw[k,c] += sum_{b,y,x} du[b,y,x,c] * x[b,y+ky-3,x+kx-3,c];  db[c] += sum du
__global__ void __launch_bounds__(256)
dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ du, float* __restrict__ dw,
                     float* __restrict__ db, int64_t B, int H, int W, int C, int imgs_per_block) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  const int64_t b0 = (int64_t)blockIdx.y * imgs_per_block, b1 = min(B, b0 + imgs_per_block);
  float acc[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) acc[k] = 0.f;
  float accb = 0.f;
  const int HW = H * W;
  if (c < C) {
    for (int64_t p = b0 * HW + sub; p < b1 * HW; p += 8) {
      const int64_t b = p / HW;
      const int r = (int)(p - b * HW);
      const int oy = r / W, ox = r - oy * W;
      const float g = du[p * C + c];
      accb += g;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int iy = oy + ky - 3;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ix = ox + kx - 3;
          if (ix < 0 || ix >= W) continue;
          acc[ky * 7 + kx] = fmaf(g, x[((b * H + iy) * W + ix) * (int64_t)C + c], acc[ky * 7 + kx]);
        }
      }
    }
  }
  __shared__ float red[8][33];
#pragma unroll
  for (int k = 0; k < 50; ++k) {
    red[sub][threadIdx.x & 31] = k < 49 ? acc[k] : accb;
    __syncthreads();
    if (sub == 0 && c < C) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
      if (k < 49) atomicAdd(dw + k * C + c, t); else atomicAdd(db + c, t);
    }
    __syncthreads();
  }
}

// ---- square-map fast paths (S x S maps, S in {15, 7, 3, 1}: every ConvNeXt stage of a 63 x 63 cutout) ----------------
// thread = (channel, output row): the S outputs of the row live in registers, each of the <= 7 input rows is read once
// (coalesced over channels, re-reads by the neighbouring row threads hit L1) and its 7 taps are applied to the whole
// row with compile-time bounds -- 7*S FMAs per S + 7 loads instead of one global load per FMA.
template <int S>
__global__ void __launch_bounds__(32 * S)
dwconv7_rows_kernel(const float* __restrict__ x, const float* __restrict__ w49, const float* __restrict__ bias,
                    float* __restrict__ out, int64_t B, int C, int flip) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int oy = threadIdx.y;
  if (c >= C) return;
  const float bv = bias ? bias[c] : 0.f;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    const float* xb = x + b * S * S * (int64_t)C + c;
    float acc[S];
#pragma unroll
    for (int i = 0; i < S; ++i) acc[i] = bv;
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy + ky - 3;
      if (iy < 0 || iy >= S) continue;
      float xr[S], wv[7];
#pragma unroll
      for (int ix = 0; ix < S; ++ix) xr[ix] = xb[(iy * S + ix) * (int64_t)C];
      const float* wr = w49 + (flip ? (6 - ky) * 7 : ky * 7) * C + c;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) wv[kx] = __ldg(wr + (flip ? 6 - kx : kx) * C);
#pragma unroll
      for (int ox = 0; ox < S; ++ox)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int ix = ox + kx - 3;
          if (ix >= 0 && ix < S) acc[ox] = fmaf(wv[kx], xr[ix], acc[ox]);
        }
    }
    float* ob = out + (b * S * S + oy * S) * (int64_t)C + c;
#pragma unroll
    for (int ox = 0; ox < S; ++ox) ob[ox * (int64_t)C] = acc[ox];
  }
}

// thread = (channel, tap row ky): its 7 tap accumulators stay in registers over `ipb` images and every output row
// (du row oy against x row oy + ky - 3: 7*S FMAs per 2*S loads); every (tap, channel) is owned by exactly one thread
// of the block, so the only reduction is one atomicAdd per block at the end
template <int S>
__global__ void __launch_bounds__(32 * 7)
dwconv7_wgrad_rows_kernel(const float* __restrict__ x, const float* __restrict__ du, float* __restrict__ dw,
                          float* __restrict__ db, int64_t B, int C, int ipb) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int ky = threadIdx.y;
  if (c >= C) return;
  float acc[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = 0.f;
  float accb = 0.f;
  const int64_t b0 = (int64_t)blockIdx.y * ipb, b1 = min(B, b0 + (int64_t)ipb);
  for (int64_t b = b0; b < b1; ++b) {
    const float* xb = x + b * S * S * (int64_t)C + c;
    const float* gb = du + b * S * S * (int64_t)C + c;
#pragma unroll 1
    for (int oy = 0; oy < S; ++oy) {
      const int iy = oy + ky - 3;
      if (iy < 0 || iy >= S) continue;
      float xr[S], dr[S];
#pragma unroll
      for (int i = 0; i < S; ++i) { xr[i] = xb[(iy * S + i) * (int64_t)C]; dr[i] = gb[(oy * S + i) * (int64_t)C]; }
      if (ky == 3) {
#pragma unroll
        for (int i = 0; i < S; ++i) accb += dr[i];               // iy == oy: every du row is summed exactly once
      }
#pragma unroll
      for (int kx = 0; kx < 7; ++kx)
#pragma unroll
        for (int ox = 0; ox < S; ++ox) {
          const int ix = ox + kx - 3;
          if (ix >= 0 && ix < S) acc[kx] = fmaf(dr[ox], xr[ix], acc[kx]);
        }
    }
  }
#pragma unroll
  for (int kx = 0; kx < 7; ++kx) atomicAdd(dw + (ky * 7 + kx) * C + c, acc[kx]);
  if (ky == 3) atomicAdd(db + c, accb);
}

template <int S>
static void launch_dwconv_rows(const float* x, const float* w49, const float* bias, float* out, int64_t B, int C, int flip,
                               cudaStream_t st) {
  const int cblocks = (C + 31) / 32;
  int64_t gy = B;
  const int64_t cap = (int64_t)148 * 16 / cblocks + 1;            // a few waves; blocks loop over images beyond that
  if (gy > cap) gy = cap;
  dwconv7_rows_kernel<S><<<dim3((unsigned)cblocks, (unsigned)gy), dim3(32, S), 0, st>>>(x, w49, bias, out, B, C, flip);
}

template <int S>
static void launch_dwconv_wgrad_rows(const float* x, const float* du, float* dw, float* db, int64_t B, int C, cudaStream_t st) {
  const int cblocks = (C + 31) / 32;
  int64_t ipb = (B * cblocks + 148 * 4 - 1) / (148 * 4);          // ~4 blocks per SM in total
  if (ipb < 1) ipb = 1;
  dwconv7_wgrad_rows_kernel<S><<<dim3((unsigned)cblocks, (unsigned)((B + ipb - 1) / ipb)), dim3(32, 7), 0, st>>>(
      x, du, dw, db, B, C, (int)ipb);
}

// stem im2col: x [B,3,H,W] -> patches [B*ho*wo, 48], k = (ci*4+ky)*4+kx
__global__ void stem_im2col_kernel(const float* __restrict__ x, float* __restrict__ p, int64_t B, int H, int W, int ho, int wo) {
  const int64_t total = B * ho * wo * 48;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % 48);
    int64_t t = i / 48;
    const int ox = (int)(t % wo); t /= wo;
    const int oy = (int)(t % ho);
    const int64_t b = t / ho;
    const int ci = k >> 4, ky = (k >> 2) & 3, kx = k & 3;
    p[i] = x[((b * 3 + ci) * H + oy * 4 + ky) * (int64_t)W + ox * 4 + kx];
  }
}

// 2x2/s2 patch gather (reverse=0): rows [B,H,W,C] -> [B*Ho*Wo, 4C]; scatter (reverse=1) writes rows, zero for dropped pixels
__global__ void patch2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t B, int H, int W, int C,
                                int Ho, int Wo, int reverse) {
  const int64_t total = B * H * W * (int64_t)C;     // iterate over the full-resolution side
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H);
    const int64_t b = t / H;
    const bool used = iy < 2 * Ho && ix < 2 * Wo;
    const int64_t pi = ((b * Ho + (iy >> 1)) * Wo + (ix >> 1)) * (4 * (int64_t)C) + ((iy & 1) * 2 + (ix & 1)) * C + c;
    if (!reverse) { if (used) dst[pi] = src[i]; }
    else dst[i] = used ? src[pi] : 0.f;
  }
}

// mean over HW (reverse=0): [B*HW,C] -> [B,C]; reverse=1: broadcast d[B,C]/HW -> [B*HW,C]
__global__ void pool_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t B, int HW, int C, int reverse) {
  const int64_t total = reverse ? B * HW * (int64_t)C : B * (int64_t)C;
  const float inv = 1.0f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    if (reverse) { const int64_t b = i / ((int64_t)HW * C); dst[i] = src[b * C + c] * inv; }
    else {
      const int64_t b = i / C;
      float s = 0.f;
      for (int p = 0; p < HW; ++p) s += src[(b * HW + p) * (int64_t)C + c];
      dst[i] = s * inv;
    }
  }
}

// ---- BatchNorm1d (training): one block per feature ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn1d_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                      float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, float eps,
                      float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd, int64_t B, int F) {
  const int f = blockIdx.x;
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += 256) s += x[i * F + f];
  red[threadIdx.x] = s; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const double mean = red[0] / (double)B;
  __syncthreads();
  double q = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += 256) { const double d = x[i * F + f] - mean; q += d * d; }
  red[threadIdx.x] = q; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  const double var = red[0] / (double)B;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (threadIdx.x == 0) {
    save_mean[f] = (float)mean; save_rstd[f] = rstd;
    const double unbiased = B > 1 ? red[0] / (double)(B - 1) : var;
    run_mean[f] = (1.f - momentum) * run_mean[f] + momentum * (float)mean;
    run_var[f] = (1.f - momentum) * run_var[f] + momentum * (float)unbiased;
  }
  for (int64_t i = threadIdx.x; i < B; i += 256) y[i * F + f] = (x[i * F + f] - (float)mean) * rstd * w[f] + b[f];
}

__global__ void __launch_bounds__(256)
bn1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ w,
                const float* __restrict__ save_mean, const float* __restrict__ save_rstd, float* __restrict__ dx,
                float* __restrict__ dw, float* __restrict__ db, int64_t B, int F) {
  const int f = blockIdx.x;
  __shared__ double r1[256], r2[256];
  const float mean = save_mean[f], rstd = save_rstd[f];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < B; i += 256) { const float g = dy[i * F + f]; s1 += g; s2 += g * (x[i * F + f] - mean) * rstd; }
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; } __syncthreads(); }
  const float sum_dy = (float)r1[0], sum_dyx = (float)r2[0];
  if (threadIdx.x == 0) { atomicAdd(dw + f, sum_dyx); atomicAdd(db + f, sum_dy); }
  if (dx) {
    const float k = w[f] * rstd / (float)B;
    for (int64_t i = threadIdx.x; i < B; i += 256) {
      const float xh = (x[i * F + f] - mean) * rstd;
      dx[i * F + f] = k * ((float)B * dy[i * F + f] - sum_dy - xh * sum_dyx);
    }
  }
}

__device__ __forceinline__ uint32_t hash_u64(uint64_t z) {   // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}
// reuse_mask=0: draw mask (keep prob 1-p), y = x*mask/(1-p); reuse_mask=1: y = x*mask/(1-p) with the stored mask (backward)
// `counter` (may be NULL) is a device-side step counter mixed into the seed: a CUDA graph replays the launch with the same
// host-side `seed`, the counter (bumped once per forward by btsb_counter_add_i64) still gives every step a fresh mask
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, uint8_t* __restrict__ mask, int64_t n,
                               float p, uint64_t seed, const int64_t* __restrict__ counter, int reuse_mask) {
  if (counter) seed += (uint64_t)(*counter) * 0x9E3779B97F4A7C15ull;
  const float scale = 1.0f / (1.0f - p);
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t k;
    if (reuse_mask) k = mask[i];
    else { k = hash_u64(seed * 0x100000001B3ull + (uint64_t)i) >= thr ? 1 : 0; mask[i] = k; }
    y[i] = k ? x[i] * scale : 0.f;
  }
}

// BCEWithLogitsLoss(pos_weight), reduction='mean': l = (1-y) x + (1 + (pw-1) y) softplus(-x)
__global__ void __launch_bounds__(256)
bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ labels, float pos_weight,
                  float* __restrict__ loss, float* __restrict__ dlogits, int64_t B, float dscale) {
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < B; i += (int64_t)gridDim.x * 256) {
    const float x = logits[i], y = labels[i];
    const float lw = 1.0f + (pos_weight - 1.0f) * y;
    const float sp = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));       // softplus(-x), stable
    s += (double)((1.0f - y) * x + lw * sp);
    if (dlogits) {
      const float sig = 1.0f / (1.0f + expf(-x));
      dlogits[i] = dscale * ((1.0f - y) - lw * (1.0f - sig));
    }
  }
  red[threadIdx.x] = s; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(loss, (float)(red[0] / (double)B));
}

// AdamW (decoupled weight decay, torch.optim.AdamW semantics incl. bias correction); g is multiplied by grad_scale first
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                             float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

// multi-tensor AdamW: up to BTSB_ADAMW_BATCH parameter tensors per launch, described BY VALUE in the kernel parameters
// (no device-side table to keep alive); block b works on 2048 consecutive elements of the tensor whose block range holds b
constexpr int kAdamwChunk = 2048;
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const __grid_constant__ btsb_adamw_batch t, float lr, float beta1, float beta2, float eps, float wd,
                   float bc1, float bc2_sqrt, float grad_scale, const int64_t* __restrict__ step_dev) {
  if (step_dev) {                       // capturable mode: the step count lives on the device (CUDA-graph replays)
    const float st = (float)(*step_dev);
    bc1 = 1.0f - powf(beta1, st);
    bc2_sqrt = sqrtf(1.0f - powf(beta2, st));
  }
  int ti = 0;
  while (ti + 1 < t.count && (int)blockIdx.x >= t.first_block[ti + 1]) ++ti;
  const int64_t base = (int64_t)((int)blockIdx.x - t.first_block[ti]) * kAdamwChunk;
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const int64_t end = min(t.n[ti], base + kAdamwChunk);
  for (int64_t i = base + threadIdx.x; i < end; i += 256) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

static int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 32) g = 148 * 32;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace btsb

using namespace btsb;

extern "C" int btsb_gemm_f32_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                                     float* C, int64_t M, int64_t N, int64_t K, int accumulate, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm_strided: negative dimension");
  if (M == 0 || N == 0) return BTSB_OK;
  BTSB_REQUIRE(A && B && C, "gemm_strided: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t gm = (M + GG_BM - 1) / GG_BM, gn = (N + GG_BN - 1) / GG_BN;
  // split K when the output is too small to fill the machine (weight-gradient shapes)
  int64_t splits = 1;
  if (gm * gn < 2 * 148 && K > 4096) {
    splits = (2 * 148 + gm * gn - 1) / (gm * gn);
    const int64_t maxs = (K + 2047) / 2048;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  if (K == 0) splits = 1;
  int64_t kchunk = (K + splits - 1) / splits;
  kchunk = ((kchunk + GG_BK - 1) / GG_BK) * GG_BK;
  if (kchunk == 0) kchunk = GG_BK;
  splits = K > 0 ? (K + kchunk - 1) / kchunk : 1;
  dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
  if (splits > 1) {
    if (!accumulate) BTSB_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st), "gemm_strided memset");
    gemm_strided_kernel<true><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, M, N, K, kchunk, 1);
  } else {
    gemm_strided_kernel<false><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, M, N, K, kchunk, accumulate);
  }
  return launch_done("gemm_strided");
}

extern "C" int btsb_colsum_f32(const float* X, const float* Y, float* out, int64_t M, int N, int accumulate, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(M >= 0 && N >= 1 && X && out, "colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) BTSB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st), "colsum memset");
  if (M == 0) return BTSB_OK;
  int64_t rpb = 512;
  dim3 grid((unsigned)((N + 31) / 32), (unsigned)((M + rpb - 1) / rpb));
  colsum_kernel<<<grid, 256, 0, st>>>(X, Y, out, M, N, rpb);
  return launch_done("colsum");
}

extern "C" int btsb_act_f32(const float* pre, const float* dout, float* out, int64_t n, int act, void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(pre && out, "act: null pointer");
  act_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(pre, dout, out, n, act);
  return launch_done("act");
}

extern "C" int btsb_colscale_f32(const float* X, const float* g, const float* res, float* out, int64_t M, int N, void* stream) {
  if (int e = check_device()) return e;
  if (M <= 0) return BTSB_OK;
  BTSB_REQUIRE(X && g && out && N >= 1, "colscale: bad arguments");
  if (N % 4 == 0 && (((uintptr_t)X | (uintptr_t)g | (uintptr_t)res | (uintptr_t)out) % 16) == 0)
    colscale4_kernel<<<ew_grid(M * N / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)X, (const float4*)g, (const float4*)res,
                                                                          (float4*)out, M * N / 4, N / 4);
  else
    colscale_kernel<<<ew_grid(M * N), 256, 0, (cudaStream_t)stream>>>(X, g, res, out, M * N, N);
  return launch_done("colscale");
}

extern "C" int btsb_bias_add_f32(float* X, const float* b, int64_t M, int N, void* stream) {
  if (int e = check_device()) return e;
  if (M <= 0) return BTSB_OK;
  BTSB_REQUIRE(X && b && N >= 1, "bias_add: bad arguments");
  bias_add_kernel<<<ew_grid(M * N), 256, 0, (cudaStream_t)stream>>>(X, b, M * N, N);
  return launch_done("bias_add");
}

extern "C" int btsb_layernorm_fwd_f32(const float* u, const float* w, const float* b, float* y, int64_t M, int C, float eps,
                                      void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(C >= 1 && C <= 32 * kLnMaxJ, "layernorm: C=%d not in [1,640]", C);
  if (M <= 0) return BTSB_OK;
  BTSB_REQUIRE(u && w && b && y, "layernorm: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(M * 32);
  if (C <= 96) ln_fwd_kernel<3><<<grid, 256, 0, st>>>(u, w, b, y, M, C, eps);
  else if (C <= 160) ln_fwd_kernel<5><<<grid, 256, 0, st>>>(u, w, b, y, M, C, eps);
  else if (C <= 320) ln_fwd_kernel<10><<<grid, 256, 0, st>>>(u, w, b, y, M, C, eps);
  else ln_fwd_kernel<20><<<grid, 256, 0, st>>>(u, w, b, y, M, C, eps);
  return launch_done("ln_fwd");
}

extern "C" int btsb_layernorm_bwd_f32(const float* u, const float* w, const float* dy, float* du, float* dw, float* db,
                                      int64_t M, int C, float eps, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(C >= 1 && C <= 32 * kLnMaxJ, "layernorm bwd: C=%d not in [1,640]", C);
  if (M <= 0) return BTSB_OK;
  BTSB_REQUIRE(u && w && dy && dw && db, "layernorm bwd: null pointer");
  int grid = ew_grid(M * 32);
  if (grid > 148 * 4) grid = 148 * 4;          // bounds the number of atomic flushes
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 96) ln_bwd_kernel<3><<<grid, 256, 0, st>>>(u, w, dy, du, dw, db, M, C, eps);
  else if (C <= 160) ln_bwd_kernel<5><<<grid, 256, 0, st>>>(u, w, dy, du, dw, db, M, C, eps);
  else if (C <= 320) ln_bwd_kernel<10><<<grid, 256, 0, st>>>(u, w, dy, du, dw, db, M, C, eps);
  else ln_bwd_kernel<20><<<grid, 256, 0, st>>>(u, w, dy, du, dw, db, M, C, eps);
  return launch_done("ln_bwd");
}

extern "C" int btsb_dwconv7_f32(const float* x, const float* w49, const float* bias, float* out, int64_t B, int H, int W,
                                int C, int flip, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && w49 && out && H >= 1 && W >= 1 && C >= 1, "dwconv7: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (H == W && H == 15) launch_dwconv_rows<15>(x, w49, bias, out, B, C, flip, st);
  else if (H == W && H == 7) launch_dwconv_rows<7>(x, w49, bias, out, B, C, flip, st);
  else if (H == W && H == 3) launch_dwconv_rows<3>(x, w49, bias, out, B, C, flip, st);
  else if (H == W && H == 1) launch_dwconv_rows<1>(x, w49, bias, out, B, C, flip, st);
  else dwconv7_kernel<<<ew_grid(B * H * W * C), 256, 0, st>>>(x, w49, bias, out, B, H, W, C, flip);
  return launch_done("dwconv7");
}

extern "C" int btsb_dwconv7_wgrad_f32(const float* x, const float* du, float* dw49, float* dbias, int64_t B, int H, int W,
                                      int C, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && du && dw49 && dbias, "dwconv7 wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (H == W && (H == 15 || H == 7 || H == 3 || H == 1)) {
    if (H == 15) launch_dwconv_wgrad_rows<15>(x, du, dw49, dbias, B, C, st);
    else if (H == 7) launch_dwconv_wgrad_rows<7>(x, du, dw49, dbias, B, C, st);
    else if (H == 3) launch_dwconv_wgrad_rows<3>(x, du, dw49, dbias, B, C, st);
    else launch_dwconv_wgrad_rows<1>(x, du, dw49, dbias, B, C, st);
    return launch_done("dwconv7_wgrad");
  }
  const int cblocks = (C + 31) / 32;
  int64_t ipb = (B * cblocks + 148 * 4 - 1) / (148 * 4);        // ~4 blocks per SM in total
  if (ipb < 1) ipb = 1;
  dim3 grid((unsigned)cblocks, (unsigned)((B + ipb - 1) / ipb));
  dwconv7_wgrad_kernel<<<grid, 256, 0, st>>>(x, du, dw49, dbias, B, H, W, C, (int)ipb);
  return launch_done("dwconv7_wgrad");
}

extern "C" int btsb_stem_im2col_f32(const float* x, float* patches, int64_t B, int H, int W, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && patches && H >= 4 && W >= 4, "im2col: bad arguments");
  const int ho = (H - 4) / 4 + 1, wo = (W - 4) / 4 + 1;
  stem_im2col_kernel<<<ew_grid(B * ho * wo * 48), 256, 0, (cudaStream_t)stream>>>(x, patches, B, H, W, ho, wo);
  return launch_done("im2col");
}

extern "C" int btsb_patch2x2_f32(const float* src, float* dst, int64_t B, int H, int W, int C, int reverse, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(src && dst && H >= 2 && W >= 2, "patch2x2: bad arguments");
  const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
  patch2x2_kernel<<<ew_grid(B * H * W * C), 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W, C, Ho, Wo, reverse);
  return launch_done("patch2x2");
}

extern "C" int btsb_pool_f32(const float* src, float* dst, int64_t B, int HW, int C, int reverse, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(src && dst && HW >= 1 && C >= 1, "pool: bad arguments");
  pool_kernel<<<ew_grid(reverse ? B * HW * C : B * C), 256, 0, (cudaStream_t)stream>>>(src, dst, B, HW, C, reverse);
  return launch_done("pool");
}

extern "C" int btsb_bn1d_train_fwd_f32(const float* x, const float* w, const float* b, float* run_mean, float* run_var,
                                       float momentum, float eps, float* y, float* save_mean, float* save_rstd, int64_t B,
                                       int F, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 2, "BatchNorm1d training needs more than one value per channel (got batch %lld)", (long long)B);
  BTSB_REQUIRE(x && w && b && run_mean && run_var && y && save_mean && save_rstd && F >= 1, "bn1d fwd: bad arguments");
  bn1d_train_fwd_kernel<<<F, 256, 0, (cudaStream_t)stream>>>(x, w, b, run_mean, run_var, momentum, eps, y, save_mean, save_rstd, B, F);
  return launch_done("bn1d_fwd");
}

extern "C" int btsb_bn1d_bwd_f32(const float* x, const float* dy, const float* w, const float* save_mean,
                                 const float* save_rstd, float* dx, float* dw, float* db, int64_t B, int F, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(x && dy && w && save_mean && save_rstd && dw && db && B >= 1 && F >= 1, "bn1d bwd: bad arguments");
  bn1d_bwd_kernel<<<F, 256, 0, (cudaStream_t)stream>>>(x, dy, w, save_mean, save_rstd, dx, dw, db, B, F);
  return launch_done("bn1d_bwd");
}

extern "C" int btsb_dropout_f32(const float* x, float* y, uint8_t* mask, int64_t n, float p, uint64_t seed, int reuse_mask,
                                void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && y && mask && p >= 0.f && p < 1.f, "dropout: bad arguments (0 <= p < 1)");
  dropout_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, mask, n, p, seed, nullptr, reuse_mask);
  return launch_done("dropout");
}

extern "C" int btsb_dropout_ctr_f32(const float* x, float* y, uint8_t* mask, int64_t n, float p, uint64_t seed,
                                    const int64_t* counter, int reuse_mask, void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && y && mask && p >= 0.f && p < 1.f, "dropout: bad arguments (0 <= p < 1)");
  dropout_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, mask, n, p, seed, counter, reuse_mask);
  return launch_done("dropout");
}

__global__ void counter_add_kernel(int64_t* c, int64_t v) { *c += v; }

extern "C" int btsb_counter_add_i64(int64_t* counter, int64_t v, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(counter, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, v);
  return launch_done("counter_add");
}

extern "C" int btsb_bce_logits_f32(const float* logits, const float* labels, float pos_weight, float* loss, float* dlogits,
                                   int64_t B, float dscale, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(logits && labels && loss && B >= 1, "bce: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BTSB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st), "bce memset");
  int grid = (int)((B + 255) / 256);
  if (grid > 148) grid = 148;
  bce_logits_kernel<<<grid, 256, 0, st>>>(logits, labels, pos_weight, loss, dlogits, B, dscale / (float)B);
  return launch_done("bce");
}

extern "C" int btsb_adamw_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                              float eps, float wd, int64_t step, float grad_scale, void* stream) {
  if (int e = check_device()) return e;
  if (n <= 0) return BTSB_OK;
  BTSB_REQUIRE(p && g && m && v && step >= 1, "adamw: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale);
  return launch_done("adamw");
}

extern "C" int btsb_adamw_multi_f32(btsb_adamw_batch* batch, float lr, float beta1, float beta2, float eps, float wd,
                                    int64_t step, float grad_scale, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(batch && batch->count >= 0 && batch->count <= BTSB_ADAMW_BATCH && step >= 1, "adamw_multi: bad arguments");
  if (batch->count == 0) return BTSB_OK;
  int blocks = 0;
  for (int i = 0; i < batch->count; ++i) {
    BTSB_REQUIRE(batch->p[i] && batch->g[i] && batch->m[i] && batch->v[i] && batch->n[i] >= 1, "adamw_multi: bad tensor %d", i);
    batch->first_block[i] = blocks;
    blocks += (int)((batch->n[i] + kAdamwChunk - 1) / kAdamwChunk);
  }
  batch->first_block[batch->count] = blocks;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  adamw_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*batch, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_scale, nullptr);
  return launch_done("adamw_multi");
}

extern "C" int btsb_adamw_multi_ctr_f32(btsb_adamw_batch* batch, float lr, float beta1, float beta2, float eps, float wd,
                                        const int64_t* step_dev, float grad_scale, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(batch && batch->count >= 0 && batch->count <= BTSB_ADAMW_BATCH && step_dev, "adamw_multi_ctr: bad arguments");
  if (batch->count == 0) return BTSB_OK;
  int blocks = 0;
  for (int i = 0; i < batch->count; ++i) {
    BTSB_REQUIRE(batch->p[i] && batch->g[i] && batch->m[i] && batch->v[i] && batch->n[i] >= 1, "adamw_multi: bad tensor %d", i);
    batch->first_block[i] = blocks;
    blocks += (int)((batch->n[i] + kAdamwChunk - 1) / kAdamwChunk);
  }
  batch->first_block[batch->count] = blocks;
  adamw_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*batch, lr, beta1, beta2, eps, wd, 1.f, 1.f, grad_scale, step_dev);
  return launch_done("adamw_multi");
}
