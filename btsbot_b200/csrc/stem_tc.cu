// K2a fused (bf16): ConvNeXt patch stem -- Conv2d(3, C0, k4, s4) + bias + LayerNorm2d -- as ONE tensor-core kernel.
//
// stem_im2col_bf16 + gemm_ln (gemm_tc.cu) moved the im2col matrix through HBM: 390 MB of fp32 pixels in, 236 MB of bf16
// patches out and in again, 295 MB of rows out per 8192 alerts (127 + 137 us).  Here the A operand of the GEMM is built
// in shared memory by eight producer warps straight from the NCHW fp32 image:
//   warps 10-17  producers : two tile slots of four warps, thread = output pixel of a 128-row tile; 48 scalar loads
//                            (3 ch x 4 x 4 patch, neighbouring lanes read neighbouring patches so every 32-byte sector is
//                            used by the 4 kx loads; two tiles' loads are in flight per SM), packed
//                            to bf16 and stored as the six 16-byte chunks of its row in the 128B-swizzled K-major layout
//                            UMMA reads (chunks 6, 7 = K 48..63 are never touched: only K = 48 is multiplied);
//                            fence.proxy.async + one mbarrier arrival per warp
//   warp 0       loads the [C0 x 64] weight tile once (TMA)
//   warp 1       MMA issuer: three tcgen05.mma (K = 16 each) per tile into one of two TMEM accumulators
//   warps 2-9    LayerNorm epilogue: thread = output row, the whole row (C0 fp32 values) in registers, ONE TMEM pass
#include <string.h>

#include "tc_common.cuh"

namespace btsb {

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
int num_sms();

namespace {
constexpr int SM_ = 128;                 // rows (output pixels) per tile
constexpr int kStagesA = 4;              // A tiles in flight
constexpr int kATile = SM_ * 128;        // [128 x 64] bf16
constexpr int kGroups = 2;               // epilogue groups (4 warps each) = TMEM accumulators in flight
constexpr int kAccCols = 128;
constexpr int kEpiW = 4 * kGroups, kProdW = 8;
constexpr int kThreadsS = (2 + kEpiW + kProdW) * 32;          // 576 -> 113 registers per thread: the row fits
constexpr int kOffW = kStagesA * kATile;                       // weight tile [<=128 x 64] bf16 = 16 KB
constexpr int kOffBarS = kOffW + 128 * 128;
constexpr int kOffVecS = kOffBarS + 256;
constexpr int kSmemS = kOffVecS + 3 * 128 * 4 + 1024;

// N = C0 at compile time: the epilogue keeps the whole output row (N fp32 accumulators) in registers -> ONE pass over
// TMEM for mean, variance and normalisation (gemm_tc's runtime-N EPI_LN makes three and was the pacing part: the unfused
// GEMM took 137 us with TMA-fed operands)
template <int N, bool XF16>        // XF16: the output rows open the fp16 residual stream (common.cuh) instead of bf16
__global__ void __launch_bounds__(kThreadsS, 1)
stem_fused_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b, __nv_bfloat16* __restrict__ out,
                  int M, int H, int W, int ho, int wo) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t bar0 = sbase + kOffBarS;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                    // kStagesA, count 4 (producer warps of a tile)
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStagesA + s); };      // kStagesA, count 1 (commit)
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStagesA + s); };  // kGroups
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStagesA + kGroups + s); };
  const uint32_t w_bar = bar0 + 8u * (2 * kStagesA + 2 * kGroups);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + kOffBarS + 8 * (2 * kStagesA + 2 * kGroups + 1));
  float* bias_s = reinterpret_cast<float*>(sal + kOffVecS);
  float* lw_s = bias_s + 128;
  float* lb_s = lw_s + 128;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_tiles = (M + SM_ - 1) / SM_;
  for (int i = threadIdx.x; i < N; i += kThreadsS) { bias_s[i] = __ldg(bias + i); lw_s[i] = __ldg(ln_w + i); lb_s[i] = __ldg(ln_b + i); }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kStagesA; ++s) { mbar_init(full_bar(s), 4); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < kGroups; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_bar, (uint32_t)(N * 128));
      tma_load_2d(sbase + kOffW, &tmW, w_bar, 0, 0);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait_spin(w_bar, 0);
    constexpr uint32_t idesc = idesc_bf16_f32(SM_, N);
    const uint64_t bdesc = smem_desc_sw128(sbase + kOffW);
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_spin(tempty_bar(as), aphase ^ 1u);
      mbar_wait_spin(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = smem_desc_sw128(sbase + stage * kATile);
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccCols);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)                       // K = 48: the fourth 16-wide step would multiply zeros
          umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk != 0 ? 1u : 0u);
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++stage == kStagesA) { stage = 0; phase ^= 1u; }
      if (++as == kGroups) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 2 + kEpiW) {
    // ===================== A producers: two tile slots of four warps, thread = one output pixel =====================
    const int pw = warp - 2 - kEpiW;
    const int slot = pw >> 2;                                  // handles local tiles lt with (lt & 1) == slot
    const int p = (pw & 3) * 32 + lane;                        // row of the tile
    const int hw = ho * wo;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if ((lt & 1) != slot) continue;
      const int stage = lt % kStagesA;
      const uint32_t phase = (uint32_t)(lt / kStagesA) & 1u;
      const int m = tile * SM_ + p;
      float v[48];
      if (m < M) {
        const int b = m / hw, r = m - b * hw;
        const int oy = r / wo, ox = r - oy * wo;
        const float* src = x + ((size_t)b * 3 * H + (size_t)oy * 4) * W + ox * 4;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 4; ++ky) {
            const float* rp = src + ((size_t)ci * H + ky) * W;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) v[(ci * 4 + ky) * 4 + kx] = __ldg(rp + kx);
          }
      } else {
#pragma unroll
        for (int i = 0; i < 48; ++i) v[i] = 0.f;
      }
      mbar_wait_spin(empty_bar(stage), phase ^ 1u);            // the MMAs that read this stage have retired
      unsigned char* rowp = sal + stage * kATile + (p >> 3) * 1024 + (p & 7) * 128;
#pragma unroll
      for (int j = 0; j < 6; ++j) {                            // 16-byte chunk j = k in [8j, 8j+8), swizzled by the row
        const uint4 q = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                   pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        *reinterpret_cast<uint4*>(rowp + ((j ^ (p & 7)) << 4)) = q;
      }
      fence_proxy_async();                                     // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
    }
  } else {
    // ===================== LayerNorm epilogue: thread = one output row, the whole row in registers =====================
    const int group = (warp - 2) >> 2, quarter = warp & 3;
    constexpr int chunks = N / 16;
    constexpr float invN = 1.0f / (float)N;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if ((lt % kGroups) != group) continue;
      const int as = group; const uint32_t aphase = (uint32_t)(lt / kGroups) & 1u;
      mbar_wait_spin(tfull_bar(as), aphase);
      tc_fence_after();
      const int row = tile * SM_ + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccCols);
      uint32_t r[chunks][16];
#pragma unroll
      for (int ch = 0; ch < chunks; ++ch) tmem_ld16(taddr + (uint32_t)(ch * 16), r[ch]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));               // accumulator is in registers: the next tile may overwrite it
      float s = 0.f;
#pragma unroll
      for (int ch = 0; ch < chunks; ++ch)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = __uint_as_float(r[ch][i]) + bias_s[ch * 16 + i];
          r[ch][i] = __float_as_uint(a);
          s += a;
        }
      const float mean = s * invN;
      float q = 0.f;
#pragma unroll
      for (int ch = 0; ch < chunks; ++ch)
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float d = __uint_as_float(r[ch][i]) - mean; q = fmaf(d, d, q); }
      const float rstd = rsqrtf(q * invN + kLnEps);
      if (row < M) {
#pragma unroll
        for (int ch = 0; ch < chunks; ++ch) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            v[i] = (__uint_as_float(r[ch][i]) - mean) * rstd * lw_s[ch * 16 + i] + lb_s[ch * 16 + i];
          uint4 o0, o1;
          o0.x = pack_x2<XF16>(v[0], v[1]); o0.y = pack_x2<XF16>(v[2], v[3]);
          o0.z = pack_x2<XF16>(v[4], v[5]); o0.w = pack_x2<XF16>(v[6], v[7]);
          o1.x = pack_x2<XF16>(v[8], v[9]); o1.y = pack_x2<XF16>(v[10], v[11]);
          o1.z = pack_x2<XF16>(v[12], v[13]); o1.w = pack_x2<XF16>(v[14], v[15]);
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)row * N + ch * 16);
          op[0] = o0; op[1] = o1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

template <int N, bool XF16>
static int launch_stem(const float* x, const CUtensorMap& tmW, const float* bias, const float* ln_w, const float* ln_b,
                       void* out, int64_t M, int H, int W, int ho, int wo, cudaStream_t st) {
  auto kern = stem_fused_kernel<N, XF16>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemS), "stem_fused attr");
  const int tiles = (int)((M + SM_ - 1) / SM_);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kThreadsS, kSmemS, st>>>(x, tmW, bias, ln_w, ln_b, (__nv_bfloat16*)out, (int)M, H, W, ho, wo);
  return launch_done("stem_fused");
}

}  // namespace
}  // namespace btsb

using namespace btsb;

extern "C" int btsb_stem_fused_fwd(const float* x, int64_t B, int H, int W, const void* w_pad, const float* bias,
                                   const float* ln_w, const float* ln_b, void* out, int C0, int out_dtype, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(out_dtype == BTSB_BF16 || out_dtype == BTSB_BF16_XF16, "stem_fused: out_dtype must be BF16 or BF16_XF16");
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && w_pad && bias && ln_w && ln_b && out && H >= 4 && W >= 4, "stem_fused: bad arguments");
  BTSB_REQUIRE(C0 == 64 || C0 == 80 || C0 == 96, "stem_fused: C0=%d is not instantiated (64, 80, 96); use im2col + gemm_ln", C0);
  BTSB_REQUIRE(((uintptr_t)out % 16) == 0 && ((uintptr_t)w_pad % 16) == 0, "stem_fused: out / weights must be 16-byte aligned");
  const int ho = (H - 4) / 4 + 1, wo = (W - 4) / 4 + 1;
  const int64_t M = B * ho * wo;
  BTSB_REQUIRE(M < (1ll << 31), "stem_fused: too many output pixels");
  CUtensorMap tmW;
  if (int e = make_tmap_bf16_2d(&tmW, w_pad, (uint64_t)C0, 64, (uint32_t)C0)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == BTSB_BF16_XF16) {
    switch (C0) {
      case 64: return launch_stem<64, true>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
      case 80: return launch_stem<80, true>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
      case 96: return launch_stem<96, true>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
    }
  }
  switch (C0) {
    case 64: return launch_stem<64, false>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
    case 80: return launch_stem<80, false>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
    case 96: return launch_stem<96, false>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
  }
  set_error("stem_fused: C0=%d is not instantiated (64, 80, 96); use im2col + gemm_ln", C0);
  return BTSB_EINVAL;
}
