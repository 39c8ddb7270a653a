// fp32 GEMMs of the 1e-4 ("fp32") mode on the tensor cores: out[M,N] = epi(A[M,K] . Wt[N,K]^T + bias), fp32 in / out,
// computed as THREE tcgen05.mma kind::tf32 products per K step (the 3xTF32 split):
//
//     a = a_hi + a_lo,  w = w_hi + w_lo      (hi = the top 19 bits of the fp32 word, lo = a - hi, again cut to 19 bits:
//                                              every operand the tensor core sees is exactly representable in tf32, so
//                                              the result does not depend on how the hardware rounds its inputs)
//     a.w ~= a_hi.w_hi + a_hi.w_lo + a_lo.w_hi (the dropped a_lo.w_lo term and the second cut are <= 2^-21 relative)
//
// The tensor core's fp32 accumulator truncates on every accumulation step: over a long K the error grows linearly (measured
// 2-7e-5 absolute on O(1) outputs at K = 640 .. 2560 when a whole tile's K was accumulated in TMEM, profiles/r02f).  So TMEM
// only ever holds the sum over ONE chunk of kChunkKB = 2 K blocks (K = 64: 24 accumulation steps; with 4 blocks the GEMM error was 6e-6, the logits moved by 2e-5); the epilogue warps drain
// every chunk into fp32 registers (round-to-nearest adds on the CUDA cores) while the next chunk accumulates in the other
// TMEM buffer.  Replaces the CUDA-core sgemm_kernel (convnext_simt.cu) for the trunk's fc1 / fc2 /
// downsample GEMMs (timm mlp.fc1, mlp.fc2, downsample.1; MaxViT 1x1 convs / Linears) whenever K % 4 == 0, N % 16 == 0.
//
//   warp 0        TMA producer: fp32 tiles A [128 x 32] and W [BN x 32] (128-byte rows, 128B swizzle) into a 4-slot ring
//   warps 2..9    splitters: rewrite each landed tile IN PLACE as its hi part and write the lo part to a twin tile at the
//                 same (swizzled) offsets of a 2-slot lo ring -- an element-wise pass, the swizzle is never decoded
//   warp 1        MMA issuer: per K step of 8 three UMMAs (M128 x N=BN x K8), accumulators double buffered in TMEM
//   warps 10..17  epilogue: tcgen05.ld -> bias / erf-GELU / SiLU / gamma*(acc+b)+res in fp32 -> 64-byte row pieces
#include <string.h>

#include "tc_common.cuh"

namespace btsb {
namespace {
constexpr int TM = 128;                // rows per tile
constexpr int TK = 32;                 // fp32 columns per K block (128 bytes)
constexpr int TBN = 128;               // max tile width
constexpr int kStagesT = 4;            // raw (TMA) ring: fp32 A + W tiles, rewritten in place as their hi parts
constexpr int kLoStages = 2;           // lo ring: the twin tiles exist only between the splitters and the MMAs
constexpr int kChunkKB = 2;            // K blocks accumulated in TMEM before the sum moves to registers
constexpr int kATile = TM * 128;       // 16 KB
constexpr int kWTile = TBN * 128;      // 16 KB
constexpr int kStageT = kATile + kWTile;                // one raw slot: A, W (-> A_hi, W_hi); one lo slot: A_lo, W_lo
constexpr int kSplitWarps = 8, kEpiWarpsT = 8;
constexpr int kThreadsT = 64 + 32 * (kSplitWarps + kEpiWarpsT);
constexpr int kMaxNT = 2560;
constexpr int kOffLo = kStagesT * kStageT;
constexpr int kOffBarT = kOffLo + kLoStages * kStageT;
constexpr int kOffVecT = kOffBarT + 256;
constexpr int kSmemT = kOffVecT + 2 * kMaxNT * 4 + 1024;
static_assert(kSmemT <= 227 * 1024, "shared-memory plan");
static_assert(kWTile <= kATile, "the splitters handle at most as many W pieces as A pieces");

__host__ __device__ constexpr uint32_t idesc_tf32_f32(int M, int N) {
  // D fp32 (1 << 4), A / B tf32 (format 2 at bits 7 / 10), both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float cut19(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// hi / lo split of one 16-byte piece, in place + twin
__device__ __forceinline__ void split4(float4* hi_p, float4* lo_p) {
  const float4 v = *hi_p;
  const float4 h = make_float4(cut19(v.x), cut19(v.y), cut19(v.z), cut19(v.w));
  *hi_p = h;
  *lo_p = make_float4(cut19(v.x - h.x), cut19(v.y - h.y), cut19(v.z - h.z), cut19(v.w - h.w));
}
}  // namespace

template <int EPI>
__global__ void __launch_bounds__(kThreadsT, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ res,
                   float* __restrict__ out, int M, int N, int K, int BN) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const uint32_t bar0 = sbase + kOffBarT;
  // The TMA ring is deeper than the lo ring: with one 64 KB stage holding all four tiles only three stages fitted and a
  // single one was ever in flight from L2 (ncu: the MMA warp waited on the split barrier, tensor pipe 35 %)
  auto full_bar = [&](int s) { return bar0 + 8u * s; };                       // raw slot: TMA bytes landed
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStagesT + s); };         // raw slot: the MMAs reading its hi tiles retired
  auto split_bar = [&](int l) { return bar0 + 8u * (2 * kStagesT + l); };     // lo slot (and its raw slot's hi rewrite) written
  auto lo_empty = [&](int l) { return bar0 + 8u * (2 * kStagesT + kLoStages + l); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStagesT + 2 * kLoStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStagesT + 2 * kLoStages + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + kOffBarT + 8 * (2 * kStagesT + 2 * kLoStages + 4));
  float* bias_s = reinterpret_cast<float*>(sal + kOffVecT);
  float* gamma_s = bias_s + kMaxNT;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + TM - 1) / TM, n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + TK - 1) / TK;

  for (int i = threadIdx.x; i < N; i += kThreadsT) {
    bias_s[i] = __ldg(bias + i);
    if (EPI == BTSB_EPI_SCALE_RES) gamma_s[i] = __ldg(gamma + i);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStagesT; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int l = 0; l < kLoStages; ++l) { mbar_init(split_bar(l), kSplitWarps); mbar_init(lo_empty(l), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarpsT); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 2 * TBN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t tx_bytes = (uint32_t)(TM + BN) * 128u;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TM, n0 = (tile % n_tiles) * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait_spin(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t sa = sbase + stage * kStageT;
          mbar_expect_tx(full_bar(stage), tx_bytes);
          tma_load_2d(sa, &tmA, full_bar(stage), kb * TK, m0);                    // rows / columns past the tensor: zeros
          tma_load_2d(sa + kATile, &tmB, full_bar(stage), kb * TK, n0);
        }
        __syncwarp();
        if (++stage == kStagesT) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    int lo = 0; uint32_t lphase = 0;
    uint32_t cidx = 0;                                           // running chunk count: TMEM buffer = cidx & 1
    const uint32_t idesc = idesc_tf32_f32(TM, BN);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int as = (int)(cidx & 1u);
        const bool first = (kb % kChunkKB) == 0, last = (kb % kChunkKB) == kChunkKB - 1 || kb == k_blocks - 1;
        if (first) {
          mbar_wait_spin(tempty_bar(as), ((cidx >> 1) & 1u) ^ 1u);      // the epilogue has drained this buffer
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * TBN);
        mbar_wait_spin(split_bar(lo), lphase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + stage * kStageT, sl = sbase + kOffLo + lo * kStageT;
          const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sl);
          const uint64_t w_hi = smem_desc_sw128(sa + kATile), w_lo = smem_desc_sw128(sl + kATile);
          const int kmax = (min(TK, K - kb * TK) + 7) / 8;
          for (int kk = 0; kk < kmax; ++kk) {
            // 8 fp32 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            const uint64_t o = (uint64_t)(2 * kk);
            umma_tf32(tmem_d, a_lo + o, w_hi + o, idesc, (first && kk == 0) ? 0u : 1u);      // small terms first
            umma_tf32(tmem_d, a_hi + o, w_lo + o, idesc, 1u);
            umma_tf32(tmem_d, a_hi + o, w_hi + o, idesc, 1u);
          }
          umma_commit(empty_bar(stage));
          umma_commit(lo_empty(lo));
          if (last) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (last) ++cidx;
        if (++stage == kStagesT) { stage = 0; phase ^= 1u; }
        if (++lo == kLoStages) { lo = 0; lphase ^= 1u; }
      }
    }
  } else if (warp < 2 + kSplitWarps) {
    // ===================== splitters: fp32 tile -> (hi in place, lo twin) =====================
    const int st = (warp - 2) * 32 + lane;                       // 0 .. 255
    int stage = 0; uint32_t phase = 0;
    int lo = 0; uint32_t lphase = 0;
    const int a_vec = kATile / 16, w_vec = BN * 128 / 16;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait_spin(full_bar(stage), phase);
        mbar_wait_spin(lo_empty(lo), lphase ^ 1u);             // the MMAs that read this lo slot two K blocks ago retired
        unsigned char* sa = sal + stage * kStageT;
        unsigned char* sl = sal + kOffLo + lo * kStageT;
        float4* ah = reinterpret_cast<float4*>(sa);
        float4* al = reinterpret_cast<float4*>(sl);
        float4* wh = reinterpret_cast<float4*>(sa + kATile);
        float4* wl = reinterpret_cast<float4*>(sl + kATile);
        // all of this thread's 16-byte pieces are loaded before any is processed (the loop was a chain of dependent
        // LDS -> LOP/FADD -> STS round trips: ncu's top stall of the kernel, with the MMA warp waiting on split_bar)
        constexpr int kPer = kATile / 16 / (kSplitWarps * 32);     // 4 pieces of A and up to 4 of W per thread
        float4 va[kPer], vw[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) va[i] = ah[st + i * kSplitWarps * 32];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = st + i * kSplitWarps * 32;
          vw[i] = idx < w_vec ? wh[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = st + i * kSplitWarps * 32;
          const float4 v = va[i];
          const float4 h = make_float4(cut19(v.x), cut19(v.y), cut19(v.z), cut19(v.w));
          ah[idx] = h;
          al[idx] = make_float4(cut19(v.x - h.x), cut19(v.y - h.y), cut19(v.z - h.z), cut19(v.w - h.w));
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = st + i * kSplitWarps * 32;
          if (idx < w_vec) {
            const float4 v = vw[i];
            const float4 h = make_float4(cut19(v.x), cut19(v.y), cut19(v.z), cut19(v.w));
            wh[idx] = h;
            wl[idx] = make_float4(cut19(v.x - h.x), cut19(v.y - h.y), cut19(v.z - h.z), cut19(v.w - h.w));
          }
        }
        (void)a_vec;
        fence_proxy_async();                                     // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar(lo));
        if (++stage == kStagesT) { stage = 0; phase ^= 1u; }
        if (++lo == kLoStages) { lo = 0; lphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2 - kSplitWarps;                       // 0 .. 7
    const int quarter = warp & 3;                                // TMEM lanes this warp may touch
    const int part = ew >> 2;                                    // which half of the tile's 16-column chunks
    const int chunks = BN / 16;
    const int c_lo = (chunks * part) / 2, c_hi = (chunks * (part + 1)) / 2;
    constexpr int kMaxCh = TBN / 16 / 2;                         // <= 4 column chunks of 16 per warp
    const int n_kchunks = (k_blocks + kChunkKB - 1) / kChunkKB;
    uint32_t cidx = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TM, n0 = (tile % n_tiles) * BN;
      const int row = m0 + quarter * 32 + lane;
      float acc[kMaxCh][16];
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c][i] = 0.f;
      for (int kc = 0; kc < n_kchunks; ++kc, ++cidx) {
        const int as = (int)(cidx & 1u);
        mbar_wait_spin(tfull_bar(as), (cidx >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * TBN);
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) {
          if (c_lo + c < c_hi) {
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)((c_lo + c) * 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[c][i] += __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));                // this warp is done with the buffer
      }
      if (row < M) {
#pragma unroll
        for (int c = 0; c < kMaxCh; ++c) {
          if (c_lo + c >= c_hi) continue;
          const int n = n0 + (c_lo + c) * 16;
          float* op = out + (size_t)row * N + n;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n + i);
            float v[4] = {acc[c][i] + b4.x, acc[c][i + 1] + b4.y, acc[c][i + 2] + b4.z, acc[c][i + 3] + b4.w};
            if (EPI == BTSB_EPI_BIAS_GELU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
            } else if (EPI == BTSB_EPI_BIAS_SILU) {
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = v[j] / (1.0f + expf(-v[j]));
            } else if (EPI == BTSB_EPI_SCALE_RES) {
              const float4 g4 = *reinterpret_cast<const float4*>(gamma_s + n + i);
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(res + (size_t)row * N + n + i));
              v[0] = r4.x + g4.x * v[0]; v[1] = r4.y + g4.y * v[1]; v[2] = r4.z + g4.z * v[2]; v[3] = r4.w + g4.w * v[3];
            }
            *reinterpret_cast<float4*>(op + i) = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * TBN); }
}

int num_sms();
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

// returns 1 when the shape / alignment is not handled here (the caller falls back to the CUDA-core GEMM)
int gemm_tf32x3(const float* A, const float* Wt, const float* bias, const float* gamma, const float* res, float* out,
                int64_t M, int N, int K, int epilogue, cudaStream_t st) {
  if (K % 4 != 0 || N % 16 != 0 || N > kMaxNT || M >= (1ll << 31) || M < 1) return 1;
  if (((uintptr_t)A % 16) || ((uintptr_t)Wt % 16) || ((uintptr_t)out % 16) || ((uintptr_t)bias % 16)) return 1;
  if (epilogue == BTSB_EPI_SCALE_RES && (((uintptr_t)res % 16) || ((uintptr_t)gamma % 16))) return 1;
  int BN = 16;
  for (int bn = TBN; bn >= 16; bn -= 16)
    if (N % bn == 0) { BN = bn; break; }
  CUtensorMap tmA, tmB;
  if (int e = make_tmap_f32_2d(&tmA, A, (uint64_t)M, (uint64_t)K, TM)) return e;
  if (int e = make_tmap_f32_2d(&tmB, Wt, (uint64_t)N, (uint64_t)K, (uint32_t)BN)) return e;
  const int64_t tiles = ((M + TM - 1) / TM) * (int64_t)(N / BN);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
#define BTSB_TF32_LAUNCH(E)                                                                                        \
  do {                                                                                                             \
    BTSB_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemT),    \
              "gemm tf32 attr");                                                                                   \
    gemm_tf32x3_kernel<E><<<grid, kThreadsT, kSmemT, st>>>(tmA, tmB, bias, gamma, res, out, (int)M, N, K, BN);    \
  } while (0)
  if (epilogue == BTSB_EPI_BIAS) BTSB_TF32_LAUNCH(BTSB_EPI_BIAS);
  else if (epilogue == BTSB_EPI_BIAS_GELU) BTSB_TF32_LAUNCH(BTSB_EPI_BIAS_GELU);
  else if (epilogue == BTSB_EPI_BIAS_SILU) BTSB_TF32_LAUNCH(BTSB_EPI_BIAS_SILU);
  else BTSB_TF32_LAUNCH(BTSB_EPI_SCALE_RES);
#undef BTSB_TF32_LAUNCH
  return launch_done("gemm_tf32x3");
}

}  // namespace btsb
