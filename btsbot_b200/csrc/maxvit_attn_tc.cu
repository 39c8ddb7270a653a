// MaxViT 7x7 window / grid attention on the tcgen05 tensor cores (bf16 path): timm AttentionCl inside
// PartitionAttentionCl (called through /root/reference/btsbot/architectures.py:31,62-63).
//
// A CTA of 128 threads works on PAIRS of (window, head) problems: thread r owns row r of a 128-row tile -- rows 0..63 are
// the 49 tokens (+15 zero rows) of problem A, rows 64..127 those of problem B.  Per pair:
//   gather   every thread pulls its token's 192 contiguous bytes (q | k | v of the head; the window / grid partition is
//            index arithmetic) and writes q and k as rows of two K-major SW64 tiles [128 x 32] and v TRANSPOSED into two
//            K-major SW128 tiles [32 x 64] (one per problem)
//   S        = Q K^T : two tcgen05.mma (M128 x N128 x K16) into 128 TMEM columns; only the diagonal 64 x 64 blocks are
//            meaningful (row r reads the 64 columns of its own problem)
//   softmax  thread = row: tcgen05.ld 64 columns, scale + relative-position bias + mask, max / exp2 / sum entirely in
//            registers (no shuffles), P (unnormalised, bf16) back into TMEM columns that S no longer needs
//   O        = P V : per problem four tcgen05.mma (A = P from TMEM, M128 x N32 x K16) into its own 32 columns -- rows of
//            the other problem accumulate garbage there and never read it
//   scatter  tcgen05.ld 32 columns, * 1/sum, bf16, 64 contiguous bytes per token back to image order
// TMEM: 128 columns per CTA (P and O alias S), 25 KB of shared memory: four CTAs per SM overlap each other's phases.
#include "tc_common.cuh"

namespace btsb {
namespace {
constexpr int kWinT = 7, kTokT = 49, kDhT = 32;
constexpr int kThreadsA = 128;
constexpr int kQOff = 0, kKOff = 8192, kVOff = 16384;        // [128 x 64 B] q, [128 x 64 B] k, 2 x [32 x 128 B] v^T
constexpr int kTabOff = 24576;                               // 2 x 176 floats: relative-position bias column of each head
constexpr int kBarOff = kTabOff + 2 * 176 * 4;
constexpr int kSmemA = kBarOff + 64 + 1024;                  // + alignment slack
constexpr int kColS = 0, kColP = 0, kColO = 64;              // TMEM columns: S [0,128); P [0,32) and O_A / O_B [64,128) alias it

__device__ __forceinline__ uint64_t desc_k64(uint32_t saddr) {      // K-major tile with 64-byte rows, 64B swizzle
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;                                  // 8 rows x 64 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                                           // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void tmem_ld32a(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32a(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreadsA, 4)
mv_attn_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int W, int C, int heads,
                  int grid_mode, const float* __restrict__ table, int64_t nitems) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  float* tab = reinterpret_cast<float*>(sal + kTabOff);
  const uint32_t bar = sbase + kBarOff;                      // one mbarrier: MMA batch retired
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + kBarOff + 16);
  const int r = threadIdx.x, warp = r >> 5;
  const int half = r >> 6, t = r & 63;                       // problem A / B of the pair, token index (>= 49: padding row)
  const bool tok = t < kTokT;
  if (r == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32((const void*)tmem_slot), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);

  const int nwx = W / kWinT, nwy = H / kWinT;
  constexpr float kLog2e = 1.4426950408889634f;
  const float sc = 0.17677669529663687f * kLog2e;            // dim_head^-0.5, in the exp2 domain
  // row-side part of the relative-position index (Swin formula): idx = (yi - yj + 6) * 13 + (xi - xj + 6)
  const int io = tok ? (t / kWinT) * 13 + t % kWinT + 84 : 84;
  constexpr uint32_t idesc_s = idesc_bf16_f32(128, 128);
  constexpr uint32_t idesc_o = idesc_bf16_f32(128, 32);
  uint32_t phase = 0;
  const int64_t npairs = (nitems + 1) / 2;

  // image-order row of this thread's token in problem `item`, and the head
  auto locate = [&](int64_t item, int& h, int64_t& row_g) {
    h = (int)(item % heads);
    int64_t win = item / heads;
    const int wx = (int)(win % nwx); win /= nwx;
    const int wy = (int)(win % nwy);
    const int64_t b = win / nwy;
    const int ty = t / kWinT, tx = t - ty * kWinT;
    const int y = grid_mode ? ty * nwy + wy : wy * kWinT + ty;
    const int x = grid_mode ? tx * nwx + wx : wx * kWinT + tx;
    row_g = (b * H + y) * (int64_t)W + x;
  };
  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const int64_t item = 2 * pair + half;
    const bool live = item < nitems;
    int h = 0;
    int64_t row_g = 0;
    if (live) locate(item, h, row_g);
    {
      // the NEXT pair's 192 bytes go to L2 now: with four CTAs per SM and a strictly serial gather -> MMA -> softmax ->
      // MMA -> scatter chain per CTA, the DRAM latency of the gather is otherwise fully exposed once per pair
      const int64_t nitem = 2 * (pair + gridDim.x) + half;
      if (tok && nitem < nitems) {
        int nh; int64_t nrow;
        locate(nitem, nh, nrow);
        const char* np = reinterpret_cast<const char*>(qkv + nrow * 3 * C + nh * 3 * kDhT);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(np));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(np + 128));
      }
    }
    // ---- gather: 192 contiguous bytes per token -> q / k rows (SW64) and the v column of the transposed tile (SW128)
    uint4 v4[4];
    {
      uint4 q4[4], k4[4];
      if (live && tok) {
        const uint4* src = reinterpret_cast<const uint4*>(qkv + row_g * 3 * C + h * 3 * kDhT);
#pragma unroll
        for (int i = 0; i < 4; ++i) { q4[i] = __ldg(src + i); k4[i] = __ldg(src + 4 + i); v4[i] = __ldg(src + 8 + i); }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) q4[i] = k4[i] = v4[i] = make_uint4(0, 0, 0, 0);       // padding rows: exact zeros
      }
      const int xr = (r >> 1) & 3;                           // 64B swizzle: 16-byte chunk index ^ address bits [7,9)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<uint4*>(sal + kQOff + r * 64 + ((i ^ xr) << 4)) = q4[i];
        *reinterpret_cast<uint4*>(sal + kKOff + r * 64 + ((i ^ xr) << 4)) = k4[i];
      }
    }
    {
      // v^T tile of this problem: element (d, t) at sw128_offset(d, t) -- 32 two-byte stores per thread
      unsigned char* vt = sal + kVOff + half * 4096;
      const uint32_t w[16] = {v4[0].x, v4[0].y, v4[0].z, v4[0].w, v4[1].x, v4[1].y, v4[1].z, v4[1].w,
                              v4[2].x, v4[2].y, v4[2].z, v4[2].w, v4[3].x, v4[3].y, v4[3].z, v4[3].w};
#pragma unroll
      for (int d = 0; d < kDhT; ++d) {
        const uint16_t e = (uint16_t)((d & 1) ? (w[d >> 1] >> 16) : (w[d >> 1] & 0xffffu));
        *reinterpret_cast<uint16_t*>(vt + sw128_offset(d, t)) = e;
      }
    }
    // this problem's column of the bias table, in the exp2 domain (both halves' 64 threads fill their own copy)
    for (int i = t; i < 169; i += 64) tab[half * 176 + i] = live ? __ldg(table + i * heads + h) * kLog2e : 0.f;
    fence_proxy_async();                                     // generic-proxy writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- S = Q K^T (K = 32: two k16 steps)
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t qd = desc_k64(sbase + kQOff), kd = desc_k64(sbase + kKOff);
        umma_bf16(tmem_base + kColS, qd, kd, idesc_s, 0u);
        umma_bf16(tmem_base + kColS, qd + 2, kd + 2, idesc_s, 1u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait_spin(bar, phase); phase ^= 1u;
    tc_fence_after();
    // ---- softmax of this thread's row over the 49 tokens of its own problem
    uint32_t sr[2][32];
    tmem_ld32a(lane_addr + (uint32_t)(kColS + half * 64), sr[0]);
    tmem_ld32a(lane_addr + (uint32_t)(kColS + half * 64 + 32), sr[1]);
    tmem_ld_wait();
    const float* tb = tab + half * 176;
    float p[kTokT];
    float mx = -3.0e38f;
#pragma unroll
    for (int j = 0; j < kTokT; ++j) {
      const int jo = (j / kWinT) * 13 + j % kWinT;           // compile-time
      p[j] = fmaf(__uint_as_float(sr[j >> 5][j & 31]), sc, tb[io - jo]);
      mx = fmaxf(mx, p[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kTokT; ++j) { p[j] = ex2a(p[j] - mx); sum += p[j]; }
    uint32_t pr[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float lo = 2 * j < kTokT ? p[2 * j] : 0.f, hi = 2 * j + 1 < kTokT ? p[2 * j + 1] : 0.f;
      pr[j] = pack_bf16x2(lo, hi);
    }
    // every thread must have pulled its S row out of TMEM before P / O overwrite those columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_st32a(lane_addr + (uint32_t)kColP, pr);             // 64 tokens = 32 packed columns
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- O_A = P V_A, O_B = P V_B (K = 64 tokens: four k16 steps each)
    if (warp == 0) {
      if (elect_one()) {
#pragma unroll
        for (int pb = 0; pb < 2; ++pb) {
          const uint64_t vd = smem_desc_sw128(sbase + kVOff + pb * 4096);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem_base + (uint32_t)(kColO + pb * 32), tmem_base + (uint32_t)(kColP + 8 * kk), vd + (uint64_t)(2 * kk),
                    idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait_spin(bar, phase); phase ^= 1u;
    tc_fence_after();
    uint32_t orow[32];
    tmem_ld32a(lane_addr + (uint32_t)(kColO + half * 32), orow);
    tmem_ld_wait();
    if (live && tok) {
      const float inv = 1.0f / sum;
      uint4* dst = reinterpret_cast<uint4*>(out + row_g * C + h * kDhT);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(orow[8 * i]) * inv, __uint_as_float(orow[8 * i + 1]) * inv),
                            pack_bf16x2(__uint_as_float(orow[8 * i + 2]) * inv, __uint_as_float(orow[8 * i + 3]) * inv),
                            pack_bf16x2(__uint_as_float(orow[8 * i + 4]) * inv, __uint_as_float(orow[8 * i + 5]) * inv),
                            pack_bf16x2(__uint_as_float(orow[8 * i + 6]) * inv, __uint_as_float(orow[8 * i + 7]) * inv));
    }
    // the next pair's gather overwrites shared memory and its S overwrites TMEM: everyone is done with both
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}
}  // namespace

int num_sms();

int maxvit_attn_bf16_tc(const void* qkv, void* out, int64_t B, int H, int W, int C, int grid_mode, const float* table,
                        cudaStream_t st) {
  const int heads = C / kDhT;
  const int64_t nitems = B * (H / kWinT) * (W / kWinT) * heads;
  BTSB_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "maxvit attn: qkv/out must be 16-byte aligned");
  BTSB_CUDA(cudaFuncSetAttribute(mv_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemA), "attn tc attr");
  const int64_t npairs = (nitems + 1) / 2;
  const int64_t cap = (int64_t)num_sms() * 4;
  const unsigned grid = (unsigned)(npairs < cap ? npairs : cap);
  mv_attn_tc_kernel<<<grid, kThreadsA, kSmemA, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, H, W, C, heads, grid_mode,
                                                    table, nitems);
  return launch_done("maxvit_attn_tc");
}

}  // namespace btsb
