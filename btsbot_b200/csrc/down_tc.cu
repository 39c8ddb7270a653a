// K5 fused (bf16): ConvNeXt downsample -- LayerNorm2d + Conv2d(Cin, N, k2, s2) + bias -- as ONE tensor-core kernel for the
// first downsample of the nano / pico trunks (15x15x80 -> 7x7x160, 15x15x64 -> 7x7x128).
//
// lnpatch + gemm_down moved the LayerNorm'ed 2x2 patch matrix through HBM: per 8192 alerts 295 MB of rows in, 257 MB of
// patches out and in again, 128 MB of rows out (125 + 65 us).  Here the A operand of the GEMM is built in shared memory by
// eight producer warps straight from the residual-stream rows (the stem kernel's scheme, stem_tc.cu):
//   warps 10-25  producers : thread = one input pixel (output pixel p of a 128-row tile, dy, dx): it loads the Cin channels
//                            of pixel (2 oy + dy, 2 ox + dx) (Cin / 8 16-byte loads), LayerNorms them in registers
//                            (two-pass variance, packed-pair arithmetic) and stores the bf16 result as Cin / 8 16-byte
//                            chunks of row p at K offset (dy 2 + dx) Cin in the 128B-swizzled K-major layout UMMA reads;
//                            fence.proxy.async + one mbarrier arrival per warp
//   warp 0       streams the weight matrix [N x 4 Cin] as [N x 64] K blocks through a 3-slot ring (TMA, L2-resident)
//   warp 1       MMA issuer: 4 Cin / 16 tcgen05.mma (M128 x N x K16) per tile into one of two TMEM accumulators
//   warps 2-9    epilogue: thread = output row: tcgen05.ld -> + bias -> bf16 / fp16 rows (the residual stream) to global
// The A tile of a whole row tile (128 x 4 Cin bf16 = 80 KB at Cin = 80) is double buffered.
#include <string.h>

#include "tc_common.cuh"

namespace btsb {

int make_tmap_bf16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, int swizzle_bytes);
int num_sms();

namespace {
constexpr int DM = 128;                  // output pixels per tile
constexpr int kWSlots = 3;
constexpr int kEpiD = 8, kProdD = 16;
constexpr int kThreadsD = (2 + kEpiD + kProdD) * 32;          // 832 -> 72 registers per thread

template <int CIN, int N>
struct DsPlan {
  static constexpr int K = 4 * CIN;
  static constexpr int KB = K / 64;                       // 64-column K blocks
  static constexpr int kABlock = DM * 128;                // [128 x 64] bf16
  static constexpr int kATile = KB * kABlock;
  static constexpr int kWBlock = N * 128;                 // [N x 64] bf16
  static constexpr int kOffW = 2 * kATile;
  static constexpr int kOffBar = kOffW + kWSlots * kWBlock;
  static constexpr int kOffVec = kOffBar + 256;
  static constexpr int kSmem = kOffVec + (N + 2 * CIN) * 4 + 1024;
  static_assert(K % 64 == 0 && CIN % 8 == 0 && N % 16 == 0 && N <= 256, "shape");
  static_assert(kWBlock % 1024 == 0, "weight slots stay 1024-byte aligned");
  static_assert(kSmem <= 227 * 1024, "shared-memory plan");
};

template <int CIN, int N, bool XF16IN, bool XF16OUT>
__global__ void __launch_bounds__(kThreadsD, 1)
down_fused_kernel(const uint16_t* __restrict__ x, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b, uint16_t* __restrict__ out,
                  int M, int H, int W, int ho, int wo) {
  using namespace tc;
  using P = DsPlan<CIN, N>;
  constexpr int NCH = CIN / 8;                                 // 16-byte chunks per input pixel
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t bar0 = sbase + P::kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };                       // 2, count kProdD (one arrival per producer warp)
  auto a_empty = [&](int s) { return bar0 + 8u * (2 + s); };                // 2, count 1 (commit)
  auto w_full = [&](int s) { return bar0 + 8u * (4 + s); };                 // kWSlots
  auto w_empty = [&](int s) { return bar0 + 8u * (4 + kWSlots + s); };      // kWSlots
  auto tfull_bar = [&](int s) { return bar0 + 8u * (4 + 2 * kWSlots + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (6 + 2 * kWSlots + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + P::kOffBar + 8 * (8 + 2 * kWSlots));
  float* bias_s = reinterpret_cast<float*>(sal + P::kOffVec);
  float4* lwb_s = reinterpret_cast<float4*>(bias_s + N);      // per channel pair: (ln_w[2i], ln_w[2i+1], ln_b[2i], ln_b[2i+1])

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int num_tiles = (M + DM - 1) / DM;
  for (int i = threadIdx.x; i < N; i += kThreadsD) bias_s[i] = __ldg(bias + i);
  for (int i = threadIdx.x; i < CIN / 2; i += kThreadsD)
    lwb_s[i] = make_float4(__ldg(ln_w + 2 * i), __ldg(ln_w + 2 * i + 1), __ldg(ln_b + 2 * i), __ldg(ln_b + 2 * i + 1));
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full(s), kProdD); mbar_init(a_empty(s), 1);
      mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4);
    }
    for (int s = 0; s < kWSlots; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight stream: [N x 64] K blocks, KB per tile =====================
    int slot = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < P::KB; ++kb) {
        mbar_wait_spin(w_empty(slot), phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(w_full(slot), (uint32_t)P::kWBlock);
          tma_load_2d(sbase + P::kOffW + slot * P::kWBlock, &tmW, w_full(slot), kb * 64, 0);
        }
        __syncwarp();
        if (++slot == kWSlots) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_bf16_f32(DM, N);
    int slot = 0; uint32_t wphase = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1; const uint32_t aphase = (uint32_t)(lt >> 1) & 1u;
      mbar_wait_spin(tempty_bar(ab), aphase ^ 1u);             // the epilogue has read this accumulator's previous tile
      mbar_wait_spin(a_full(ab), aphase);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(ab * 256);
      for (int kb = 0; kb < P::KB; ++kb) {
        mbar_wait_spin(w_full(slot), wphase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = smem_desc_sw128(sbase + ab * P::kATile + kb * P::kABlock);
          const uint64_t bdesc = smem_desc_sw128(sbase + P::kOffW + slot * P::kWBlock);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0 ? 1u : 0u);
          umma_commit(w_empty(slot));
          if (kb == P::KB - 1) { umma_commit(a_empty(ab)); umma_commit(tfull_bar(ab)); }
        }
        __syncwarp();
        if (++slot == kWSlots) { slot = 0; wphase ^= 1u; }
      }
    }
  } else if (warp >= 2 + kEpiD) {
    // ===================== A producers: thread = one input pixel (output pixel p, dy, dx) of a 128-row tile ==============
    // 16 warps x 32 = the 512 input pixels of a tile: one batch of Cin / 8 16-byte loads per thread and tile (the four
    // threads of an output pixel read two runs of 2 Cin contiguous elements), packed-pair arithmetic (FADD2 / FFMA2)
    const int pt = (warp - 2 - kEpiD) * 32 + lane;             // 0 .. 511
    const int p = pt >> 2, dy = (pt >> 1) & 1, dx = pt & 1;
    const int hw = ho * wo;
    constexpr float invC = 1.0f / (float)CIN;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1; const uint32_t aphase = (uint32_t)(lt >> 1) & 1u;
      const int m = tile * DM + p;
      uint4 v[NCH];
      if (m < M) {
        const int b = m / hw, r = m - b * hw;
        const int oy = r / wo, ox = r - oy * wo;
        const uint4* src = reinterpret_cast<const uint4*>(x + (((size_t)b * H + (2 * oy + dy)) * W + (2 * ox + dx)) * CIN);
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] = __ldg(src + j);
      } else {
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] = make_uint4(0, 0, 0, 0);
      }
      f32x2_t s2 = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) s2 = add_f32x2(s2, x2_to_f32x2<XF16IN>(u[k]));
      }
      const float2 ss = unpack_f32x2(s2);
      const float mean = (ss.x + ss.y) * invC;
      const f32x2_t nm2 = pack_f32x2(-mean, -mean);
      f32x2_t q2 = pack_f32x2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const f32x2_t d = add_f32x2(x2_to_f32x2<XF16IN>(u[k]), nm2);
          q2 = fma3_f32x2(d, d, q2);
        }
      }
      const float2 qq = unpack_f32x2(q2);
      const float rstd = rsqrtf((qq.x + qq.y) * invC + kLnEps);
      const f32x2_t r2 = pack_f32x2(rstd, rstd), nmr2 = pack_f32x2(-mean * rstd, -mean * rstd);
      mbar_wait_spin(a_empty(ab), aphase ^ 1u);                // the MMAs that read this A buffer two tiles ago have retired
      unsigned char* tileA = sal + ab * P::kATile;
      const int c0 = (dy * 2 + dx) * NCH;                      // first 16-byte chunk of this segment in the K = 4 Cin row
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 wb = lwb_s[4 * j + k];
          const f32x2_t y = fma3_f32x2(x2_to_f32x2<XF16IN>(u[k]), r2, nmr2);            // (x - mean) * rstd
          const float2 z = unpack_f32x2(fma3_f32x2(y, pack_f32x2(wb.x, wb.y), pack_f32x2(wb.z, wb.w)));
          o[k] = m < M ? pack_bf16x2(z.x, z.y) : 0u;
        }
        const int c = c0 + j;                                  // chunk index in the row: block c >> 3, chunk c & 7 of it
        unsigned char* rowp = tileA + (c >> 3) * P::kABlock + (p >> 3) * 1024 + (p & 7) * 128;
        *reinterpret_cast<uint4*>(rowp + (((c & 7) ^ (p & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async();                                     // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(ab));
    }
  } else {
    // ===================== epilogue: thread = one output row =====================
    const int group = (warp - 2) >> 2, quarter = warp & 3;
    constexpr int chunks = N / 16;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if ((lt & 1) != group) continue;
      const int ab = group; const uint32_t aphase = (uint32_t)(lt >> 1) & 1u;
      mbar_wait_spin(tfull_bar(ab), aphase);
      tc_fence_after();
      const int row = tile * DM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * 256);
      uint32_t ra[16], rb[16];
      tmem_ld16(taddr, ra);
#pragma unroll
      for (int ch = 0; ch < chunks; ++ch) {
        uint32_t (&r)[16] = (ch & 1) ? rb : ra;
        tmem_ld_wait();
        if (ch + 1 < chunks) {
          tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), (ch & 1) ? ra : rb);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(ab));          // accumulator is in registers: the next tile may overwrite it
        }
        if (row < M) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bias_s[ch * 16 + i];
          uint4 o0, o1;
          o0.x = pack_x2<XF16OUT>(v[0], v[1]); o0.y = pack_x2<XF16OUT>(v[2], v[3]);
          o0.z = pack_x2<XF16OUT>(v[4], v[5]); o0.w = pack_x2<XF16OUT>(v[6], v[7]);
          o1.x = pack_x2<XF16OUT>(v[8], v[9]); o1.y = pack_x2<XF16OUT>(v[10], v[11]);
          o1.z = pack_x2<XF16OUT>(v[12], v[13]); o1.w = pack_x2<XF16OUT>(v[14], v[15]);
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)row * N + ch * 16);
          op[0] = o0; op[1] = o1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int CIN, int N, bool XF16IN, bool XF16OUT>
static int launch_down(const void* x, const CUtensorMap& tmW, const float* bias, const float* ln_w, const float* ln_b,
                       void* out, int64_t M, int H, int W, int ho, int wo, cudaStream_t st) {
  using P = DsPlan<CIN, N>;
  auto kern = down_fused_kernel<CIN, N, XF16IN, XF16OUT>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kSmem), "down_fused attr");
  const int tiles = (int)((M + DM - 1) / DM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kThreadsD, P::kSmem, st>>>((const uint16_t*)x, tmW, bias, ln_w, ln_b, (uint16_t*)out, (int)M, H, W, ho, wo);
  return launch_done("down_fused");
}

template <int CIN, int N>
static int dispatch_down(const void* x, bool xin16, const CUtensorMap& tmW, const float* bias, const float* ln_w,
                         const float* ln_b, void* out, bool xout16, int64_t M, int H, int W, int ho, int wo, cudaStream_t st) {
  if (xin16 && xout16) return launch_down<CIN, N, true, true>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
  if (xin16) return launch_down<CIN, N, true, false>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
  if (xout16) return launch_down<CIN, N, false, true>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
  return launch_down<CIN, N, false, false>(x, tmW, bias, ln_w, ln_b, out, M, H, W, ho, wo, st);
}

}  // namespace
}  // namespace btsb

using namespace btsb;

// K5 fused entry point: timm stages.i.downsample (LayerNorm2d + Conv2d k2 s2), called at
// /root/reference/btsbot/architectures.py:132 (oracle/convnext_oracle.py trunk_features).
// x: [B*H*W, Cin] rows in in_dtype (BTSB_BF16, or BTSB_BF16_XF16 = the fp16 residual stream); Wt: [N, 4 Cin] bf16 with column
// (dy*2+dx)*Cin + c (the lnpatch order); out: [B*Ho*Wo, N] rows in out_dtype (BF16 | BF16_XF16).  (Cin, N) = (80, 160) or
// (64, 128); other shapes: btsb_convnext_lnpatch_fwd + btsb_gemm_fwd.
extern "C" int btsb_convnext_down_fused_fwd(const void* x, int in_dtype, int64_t B, int H, int W, int Cin, const float* ln_w,
                                            const float* ln_b, const void* Wt, const float* bias, int N, void* out,
                                            int out_dtype, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(B >= 0 && H >= 2 && W >= 2, "down_fused: bad shape");
  BTSB_REQUIRE((in_dtype == BTSB_BF16 || in_dtype == BTSB_BF16_XF16) && (out_dtype == BTSB_BF16 || out_dtype == BTSB_BF16_XF16),
               "down_fused: dtypes must be BF16 or BF16_XF16");
  BTSB_REQUIRE((Cin == 80 && N == 160) || (Cin == 64 && N == 128),
               "down_fused: (Cin, N) = (%d, %d) is not instantiated ((80, 160), (64, 128)); use lnpatch + gemm", Cin, N);
  if (B == 0) return BTSB_OK;
  BTSB_REQUIRE(x && ln_w && ln_b && Wt && bias && out, "down_fused: null pointer");
  BTSB_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)Wt % 16) == 0,
               "down_fused: x / out / weights must be 16-byte aligned");
  const int ho = (H - 2) / 2 + 1, wo = (W - 2) / 2 + 1;
  const int64_t M = B * ho * wo;
  BTSB_REQUIRE(M < (1ll << 31), "down_fused: too many output pixels");
  CUtensorMap tmW;
  if (int e = make_tmap_bf16_2d_sw(&tmW, Wt, (uint64_t)N, (uint64_t)(4 * Cin), (uint32_t)N, 64, 128)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const bool xi = in_dtype == BTSB_BF16_XF16, xo = out_dtype == BTSB_BF16_XF16;
  if (Cin == 80) return dispatch_down<80, 160>(x, xi, tmW, bias, ln_w, ln_b, out, xo, M, H, W, ho, wo, st);
  return dispatch_down<64, 128>(x, xi, tmW, bias, ln_w, ln_b, out, xo, M, H, W, ho, wo, st);
}
