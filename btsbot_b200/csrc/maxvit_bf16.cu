// bf16 fast paths of the MaxViT CUDA-core kernels (the generic float / bf16 versions live in maxvit.cu): every thread
// moves 16 bytes (8 channels) per access so a warp reads / writes whole 512-byte pixel rows.
//   dw3   : depthwise 3x3 + folded BatchNorm + SiLU + SE squeeze.  CTA = (image, 128-channel slab), 16 warps; a warp
//           walks one output row with the 3x3x4-channel window AND the 36 taps in registers (one new column = 3 loads
//           per output, fetched two outputs ahead); the per-channel mean is reduced in a fixed order (deterministic).
//   ln    : row LayerNorm with C/8 lanes per row (4 / 2 / 1 rows per warp for C = 64 / 128 / 256, 2 chunks for 512).
//   scale : SE gate multiply, avgpool2: 2x2 mean -- plain streaming kernels.
#include "common.cuh"

namespace btsb {
namespace {

struct F8 { float v[8]; };

__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { r.v[2 * i] = __uint_as_float(w[i] << 16); r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  return r;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  return make_uint4(pack2(f.v[0], f.v[1]), pack2(f.v[2], f.v[3]), pack2(f.v[4], f.v[5]), pack2(f.v[6], f.v[7]));
}
// x * sigmoid(x) == 0.5 x (1 + tanh(x / 2)): one MUFU op (tanh.approx, rel. error 2^-11, below the bf16 rounding of the
// stored result) instead of ex2 + rcp -- the bf16 dw3 kernel is instruction-issue-bound
__device__ __forceinline__ float silu_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ---- depthwise 3x3 ----------------------------------------------------------------------------------------
// CTA = (image, 128-channel slab), 16 warps, lane = 4 channels.  The 36 taps of a lane's channels live in REGISTERS:
// with 8 channels per lane they had to stay in shared memory and every output re-read 9 KB of taps per warp (18
// LDS.128), which made the kernel shared-memory-bandwidth-bound at ~72 clk per warp-output (1.2-1.4 TB/s, r01h/r01j).
constexpr int kDwThreads = 512, kDwWarps = 16, kDwSlab = 128;

struct F4 { float v[4]; };
__device__ __forceinline__ F4 unpack4(const uint2& u) {
  F4 r;
  r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  return r;
}

template <int STRIDE>
__global__ void __launch_bounds__(kDwThreads, 1)
mv_dw3_bf16_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int Ho, int Wo, const float* __restrict__ w,
                   const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, float* __restrict__ pooled) {
  __shared__ float red[kDwWarps][kDwSlab];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = blockIdx.x * kDwSlab;
  const int64_t b = blockIdx.y;
  const int cl = lane * 4;                                  // this lane's 4 channels inside the slab
  float wt[9][4], sh[4], pool[4];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C + c0 + cl));
    wt[k][0] = t.x; wt[k][1] = t.y; wt[k][2] = t.z; wt[k][3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { sh[i] = shift[c0 + cl + i]; pool[i] = 0.f; }
  const __nv_bfloat16* xb = x + b * (int64_t)H * W * C + c0 + cl;
  __nv_bfloat16* ob = out + b * (int64_t)Ho * Wo * C + c0 + cl;
  const uint2 zero = make_uint2(0, 0);
  for (int oy = wid; oy < Ho; oy += kDwWarps) {
    const int iy0 = oy * STRIDE - 1;
    const bool rv[3] = {iy0 >= 0, true, iy0 + 2 < H};
    const __nv_bfloat16* rp[3] = {xb + (int64_t)(iy0 < 0 ? 0 : iy0) * W * C, xb + (int64_t)(iy0 + 1) * W * C,
                                  xb + (int64_t)(iy0 + 2 < H ? iy0 + 2 : H - 1) * W * C};
    // raw (still packed) input columns are fetched two outputs ahead of their use: a warp walks its row serially
    auto ldraw = [&](int ix, uint2 (&raw)[3]) {
      const bool cv = ix >= 0 && ix < W;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        raw[r] = (cv && rv[r]) ? __ldg(reinterpret_cast<const uint2*>(rp[r] + (int64_t)ix * C)) : zero;
    };
    auto unpack_col = [&](const uint2 (&raw)[3], F4 (&col)[3]) {
#pragma unroll
      for (int r = 0; r < 3; ++r) col[r] = unpack4(raw[r]);
    };
    F4 win[3][3];                                           // [column][row]
    auto emit = [&](int ox) {
      float acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = sh[i];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const F4& v = win[kx][ky];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(v.v[i], wt[ky * 3 + kx][i], acc[i]);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i] = silu_f(acc[i]); pool[i] += acc[i]; }
      *reinterpret_cast<uint2*>(ob + ((int64_t)oy * Wo + ox) * C) = make_uint2(pack2(acc[0], acc[1]), pack2(acc[2], acc[3]));
    };
    {
      uint2 z3[3] = {zero, zero, zero};
      unpack_col(z3, win[0]);                               // zero pad column (ix = -1)
    }
    if (STRIDE == 1) {
      uint2 ra[3], rb[3];
      ldraw(0, ra);
      unpack_col(ra, win[1]);
      ldraw(1, ra);                                         // column of output 0
      ldraw(2, rb);                                         // column of output 1
      for (int ox = 0; ox < Wo; ox += 2) {
        unpack_col(ra, win[2]);
        ldraw(ox + 3, ra);                                  // for output ox + 2
        emit(ox);
#pragma unroll
        for (int r = 0; r < 3; ++r) { win[0][r] = win[1][r]; win[1][r] = win[2][r]; }
        if (ox + 1 < Wo) {
          unpack_col(rb, win[2]);
          ldraw(ox + 4, rb);                                // for output ox + 3
          emit(ox + 1);
#pragma unroll
          for (int r = 0; r < 3; ++r) { win[0][r] = win[1][r]; win[1][r] = win[2][r]; }
        }
      }
    } else {
      uint2 a0[3], a1[3], b0[3], b1[3];                     // columns (2ox, 2ox+1) of the next two outputs
      ldraw(0, a0); ldraw(1, a1);
      ldraw(2, b0); ldraw(3, b1);
      for (int ox = 0; ox < Wo; ox += 2) {
        unpack_col(a0, win[1]); unpack_col(a1, win[2]);
        ldraw(2 * ox + 4, a0); ldraw(2 * ox + 5, a1);       // for output ox + 2
        emit(ox);
#pragma unroll
        for (int r = 0; r < 3; ++r) win[0][r] = win[2][r];
        if (ox + 1 < Wo) {
          unpack_col(b0, win[1]); unpack_col(b1, win[2]);
          ldraw(2 * ox + 6, b0); ldraw(2 * ox + 7, b1);     // for output ox + 3
          emit(ox + 1);
#pragma unroll
          for (int r = 0; r < 3; ++r) win[0][r] = win[2][r];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[wid][cl + i] = pool[i];
  __syncthreads();
  if (threadIdx.x < kDwSlab) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kDwWarps; ++k) s += red[k][threadIdx.x];
    pooled[b * C + c0 + threadIdx.x] = s / (float)(Ho * Wo);
  }
}

// ---- row LayerNorm ----------------------------------------------------------------------------------------------
// LPR lanes per row, CH 16-byte chunks per lane: C = 8 * LPR * CH
template <int LPR, int CH>
__global__ void __launch_bounds__(256)
mv_ln_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                  __nv_bfloat16* __restrict__ out, int64_t M) {
  constexpr int C = 8 * LPR * CH;
  constexpr int RPW = 32 / LPR;                             // rows per warp
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane / LPR, li = lane % LPR;
  float gw[CH][8], gb[CH][8];
#pragma unroll
  for (int j = 0; j < CH; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) { gw[j][i] = g[(j * LPR + li) * 8 + i]; gb[j][i] = bta[(j * LPR + li) * 8 + i]; }
  const int64_t rows_per_block = 8 * RPW;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block; r0 < M; r0 += (int64_t)gridDim.x * rows_per_block) {
    const int64_t row = r0 + wid * RPW + sub;
    const bool live = row < M;
    F8 v[CH];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      v[j] = unpack8(live ? __ldg(reinterpret_cast<const uint4*>(x + row * C) + j * LPR + li) : make_uint4(0, 0, 0, 0));
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[j].v[i];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[j].v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + kLnEps);
    if (live) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        F8 o8;
#pragma unroll
        for (int i = 0; i < 8; ++i) o8.v[i] = (v[j].v[i] - mean) * rstd * gw[j][i] + gb[j][i];
        reinterpret_cast<uint4*>(out + row * C)[j * LPR + li] = pack8(o8);
      }
    }
  }
}

// ---- SE excite and 2x2 average pool ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mv_scale_bf16_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ gate, int64_t B, int HW, int C) {
  const int c8 = C / 8;
  const int64_t total = B * HW * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    const int64_t b = i / ((int64_t)HW * c8);
    F8 v = unpack8(reinterpret_cast<const uint4*>(x)[i]);
    const float4 g0 = *reinterpret_cast<const float4*>(gate + b * C + c), g1 = *reinterpret_cast<const float4*>(gate + b * C + c + 4);
    v.v[0] *= g0.x; v.v[1] *= g0.y; v.v[2] *= g0.z; v.v[3] *= g0.w;
    v.v[4] *= g1.x; v.v[5] *= g1.y; v.v[6] *= g1.z; v.v[7] *= g1.w;
    reinterpret_cast<uint4*>(x)[i] = pack8(v);
  }
}

__global__ void __launch_bounds__(256)
mv_avgpool2_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, c8 = C / 8;
  const int64_t total = B * Ho * Wo * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c8) * 8;
    int64_t t = i / c8;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int64_t b = t / Ho;
    const __nv_bfloat16* p = x + ((b * H + 2 * oy) * (int64_t)W + 2 * ox) * C + c;
    const F8 a = unpack8(__ldg(reinterpret_cast<const uint4*>(p))), bq = unpack8(__ldg(reinterpret_cast<const uint4*>(p + C)));
    const F8 cq = unpack8(__ldg(reinterpret_cast<const uint4*>(p + (int64_t)W * C)));
    const F8 d = unpack8(__ldg(reinterpret_cast<const uint4*>(p + (int64_t)W * C + C)));
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = 0.25f * (a.v[k] + bq.v[k] + cq.v[k] + d.v[k]);
    reinterpret_cast<uint4*>(out)[i] = pack8(o);
  }
}

int grid_for(int64_t items, int per_block, int cap = 148 * 32) {
  int64_t g = (items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace

// returns BTSB_OK after launching, or 1 when the shape is not covered (caller falls back to the generic kernel)
int maxvit_dw3_bf16(const void* x, int64_t B, int H, int W, int C, int stride, int Ho, int Wo, const float* w,
                    const float* shift, void* out, float* pooled, cudaStream_t st) {
  if (C % kDwSlab != 0 || ((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 16) != 0) return 1;
  dim3 grid(C / kDwSlab, (unsigned)B);
  if (stride == 1)
    mv_dw3_bf16_kernel<1><<<grid, kDwThreads, 0, st>>>((const __nv_bfloat16*)x, H, W, C, Ho, Wo, w, shift, (__nv_bfloat16*)out, pooled);
  else
    mv_dw3_bf16_kernel<2><<<grid, kDwThreads, 0, st>>>((const __nv_bfloat16*)x, H, W, C, Ho, Wo, w, shift, (__nv_bfloat16*)out, pooled);
  return launch_done("maxvit_dw3_bf16");
}

int maxvit_ln_bf16(const void* x, const float* g, const float* b, void* out, int64_t M, int C, cudaStream_t st) {
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 16) != 0) return 1;
  const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
  __nv_bfloat16* xo = (__nv_bfloat16*)out;
  if (C == 64) mv_ln_bf16_kernel<8, 1><<<grid_for(M, 32 * 4), 256, 0, st>>>(xi, g, b, xo, M);
  else if (C == 128) mv_ln_bf16_kernel<16, 1><<<grid_for(M, 16 * 4), 256, 0, st>>>(xi, g, b, xo, M);
  else if (C == 256) mv_ln_bf16_kernel<32, 1><<<grid_for(M, 8 * 4), 256, 0, st>>>(xi, g, b, xo, M);
  else if (C == 512) mv_ln_bf16_kernel<32, 2><<<grid_for(M, 8 * 4), 256, 0, st>>>(xi, g, b, xo, M);
  else return 1;
  return launch_done("maxvit_ln_bf16");
}

int maxvit_scale_bf16(void* x, const float* gate, int64_t B, int HW, int C, cudaStream_t st) {
  if (C % 8 != 0 || ((uintptr_t)x % 16) != 0 || ((uintptr_t)gate % 16) != 0) return 1;
  mv_scale_bf16_kernel<<<grid_for(B * HW * (C / 8), 256 * 2), 256, 0, st>>>((__nv_bfloat16*)x, gate, B, HW, C);
  return launch_done("maxvit_scale_bf16");
}

int maxvit_avgpool2_bf16(const void* x, void* out, int64_t B, int H, int W, int C, cudaStream_t st) {
  if (C % 8 != 0 || ((uintptr_t)x % 16) != 0 || ((uintptr_t)out % 16) != 0) return 1;
  mv_avgpool2_bf16_kernel<<<grid_for(B * (H / 2) * (W / 2) * (C / 8), 256 * 2), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, B, H, W, C);
  return launch_done("maxvit_avgpool2_bf16");
}

}  // namespace btsb
