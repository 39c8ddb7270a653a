// MaxViT stem.conv2 (3x3, stride 1, pad 1, 32 -> 64 channels, bf16 NHWC) as an implicit GEMM on the tensor cores.
//
// The explicit path (maxvit im2col3 + gemm_tc) wrote and re-read a [B*H*W, 288] patch matrix: 7.4 GB per 1024 images of
// 112 x 112, 4.4 ms of the 43 ms chunk.  Here nothing is gathered by threads at all: an output tile is an 8 x 16 pixel
// rectangle of one image, and the A operand of tap (ky, kx) is simply the 8 x 16 x 32-channel box of the INPUT shifted by
// (ky-1, kx-1) -- one 4-D bulk tensor copy (cp.async.bulk.tensor.4d) whose out-of-bounds rows / columns the TMA unit
// zero-fills, which is exactly the conv padding.  The box lands as [128 rows x 64 B] with the 64-byte swizzle, i.e. the
// K-major SW64 operand layout UMMA reads (the layout of the fused MLP's 32-column K tails).
//   warp 0      TMA: the nine [N x 32] weight slabs once, then nine input boxes per tile through a 12-deep ring
//   warp 1      MMA: per tap two tcgen05.mma (K = 16) into one of two TMEM accumulators (M 128 x N 64)
//   warps 2-9   epilogue: tcgen05.ld -> + bias -> bf16 -> one 128-byte row store per thread (thread = output pixel)
#include <cuda.h>
#include <string.h>

#include "tc_common.cuh"

namespace btsb {
int num_sms();

namespace {
constexpr int TY = 8, TX = 16;                 // output tile: 8 rows x 16 columns = 128 pixels = the UMMA M
constexpr int CIN = 32;                        // K slab per tap: 32 bf16 = 64 B rows (SWIZZLE_64B)
constexpr int kSlab = TY * TX * CIN * 2;       // 8 KB
constexpr int kRing = 12;
constexpr int kThreadsC = (2 + 8) * 32;
constexpr int kMaxN = 64;
constexpr int kOffWc = kRing * kSlab;                    // 9 weight slabs [N x 32] bf16
constexpr int kOffBarC = kOffWc + 9 * kMaxN * CIN * 2;
constexpr int kOffVecC = kOffBarC + 512;
constexpr int kSmemC = kOffVecC + kMaxN * 4 + 1024;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::atomic<uint64_t> tried{0};
  if (first_use_on_device(tried)) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// K-major operand whose rows are 64 bytes (32 bf16), 64-byte swizzle, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                       // SWIZZLE_64B
  return d;
}

template <int N>
__global__ void __launch_bounds__(kThreadsC, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t bar0 = sbase + kOffBarC;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kRing + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kRing + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kRing + 2 + s); };
  const uint32_t w_bar = bar0 + 8u * (2 * kRing + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + kOffBarC + 8 * (2 * kRing + 5));
  float* bias_s = reinterpret_cast<float*>(sal + kOffVecC);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int tiles_x = W / TX, tiles_y = H / TY;
  const int tiles_img = tiles_x * tiles_y;
  const int num_tiles = B * tiles_img;
  for (int i = threadIdx.x; i < N; i += kThreadsC) bias_s[i] = bias ? __ldg(bias + i) : 0.f;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kRing; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(w_bar, (uint32_t)(9 * N * CIN * 2));
      for (int t = 0; t < 9; ++t) tma_load_2d(sbase + kOffWc + t * N * CIN * 2, &tmW, w_bar, t * CIN, 0);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_img, r = tile - b * tiles_img;
      const int y0 = (r / tiles_x) * TY, x0 = (r % tiles_x) * TX;
      for (int t = 0; t < 9; ++t) {
        mbar_wait_spin(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)kSlab);
          // box {32 ch, 16 x, 8 y, 1 image} at (x0 + kx - 1, y0 + ky - 1): rows / columns outside the image are zero-filled
          tma_load_4d(sbase + stage * kSlab, &tmX, full_bar(stage), 0, x0 + (t % 3) - 1, y0 + (t / 3) - 1, b);
        }
        __syncwarp();
        if (++stage == kRing) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait_spin(w_bar, 0);
    constexpr uint32_t idesc = idesc_bf16_f32(128, N);
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_spin(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * 64);
      for (int t = 0; t < 9; ++t) {
        mbar_wait_spin(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = smem_desc_sw64(sbase + stage * kSlab);
          const uint64_t bdesc = smem_desc_sw64(sbase + kOffWc + t * N * CIN * 2);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)                       // 16 channels (32 B) per step inside the 64-byte row
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, (t | kk) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (t == 8) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == kRing) { stage = 0; phase ^= 1u; }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ===================== epilogue: thread = output pixel =====================
    const int group = (warp - 2) >> 2, quarter = warp & 3;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if ((lt & 1) != group) continue;
      const int as = group; const uint32_t aphase = (uint32_t)(lt >> 1) & 1u;
      const int b = tile / tiles_img, r = tile - b * tiles_img;
      const int y0 = (r / tiles_x) * TY, x0 = (r % tiles_x) * TX;
      const int p = quarter * 32 + lane;                       // row of the tile = TMEM lane
      const size_t pix = ((size_t)b * H + (y0 + p / TX)) * W + (x0 + p % TX);
      mbar_wait_spin(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * 64);
      uint32_t rr[N / 16][16];
#pragma unroll
      for (int ch = 0; ch < N / 16; ++ch) tmem_ld16(taddr + (uint32_t)(ch * 16), rr[ch]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      uint4* op = reinterpret_cast<uint4*>(out + pix * N);
#pragma unroll
      for (int ch = 0; ch < N / 16; ++ch) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[ch][i]) + bias_s[ch * 16 + i];
        op[2 * ch] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        op[2 * ch + 1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                    pack_bf16x2(v[14], v[15]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

}  // namespace
}  // namespace btsb

using namespace btsb;

// x [B,H,W,32] bf16 NHWC; w [N, 288] bf16 with k = (ky*3+kx)*32 + c; bias [N] fp32 or NULL; out [B*H*W, N] bf16.
// Returns BTSB_EINVAL for shapes the implicit-GEMM kernel does not cover (the caller then uses im2col3 + gemm).
extern "C" int btsb_conv3x3_c32_fwd(const void* x, const void* w, const float* bias, void* out, int64_t B, int H, int W,
                                    int N, void* stream) {
  if (int e = check_device()) return e;
  if (B <= 0) return BTSB_OK;
  BTSB_REQUIRE(x && w && out, "conv3x3: null pointer");
  BTSB_REQUIRE(N == 64 && H % TY == 0 && W % TX == 0 && B * (int64_t)(H / TY) * (W / TX) < (1ll << 31),
               "conv3x3: only N = 64 output channels and maps that tile into 8 x 16 rectangles (H=%d W=%d N=%d)", H, W, N);
  BTSB_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)out % 16) == 0, "conv3x3: misaligned pointer");
  EncodeTiledFn enc = encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return BTSB_ECUDA; }
  CUtensorMap tmX, tmW;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, (cuuint64_t)H * W * CIN * 2};
    const cuuint32_t box[4] = {CIN, TX, TY, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv3x3: cuTensorMapEncodeTiled (input) failed with CUresult %d", (int)r); return BTSB_ECUDA; }
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)(9 * CIN), (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)(9 * CIN) * 2};
    const cuuint32_t box[2] = {CIN, (cuuint32_t)N};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv3x3: cuTensorMapEncodeTiled (weights) failed with CUresult %d", (int)r); return BTSB_ECUDA; }
  }
  static std::atomic<uint64_t> attr_done{0};
  if (first_use_on_device(attr_done)) {
    BTSB_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemC), "conv3x3 attr");
  }
  const int64_t tiles = B * (H / TY) * (W / TX);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv3x3_tc_kernel<64><<<grid, kThreadsC, kSmemC, (cudaStream_t)stream>>>(tmX, tmW, bias, (__nv_bfloat16*)out, (int)B, H, W);
  return launch_done("conv3x3_tc");
}
