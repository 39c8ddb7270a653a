// K4 fused (bf16 mode): out = res + gamma * (fc2(GELU(fc1(y) + b1)) + b2) in ONE kernel -- the 4C-wide hidden
// activation lives only in TMEM / shared memory, so per 128-row tile HBM sees y, res (read) and out (write).
// Replaces timm blocks.j.mlp.fc1 -> act -> mlp.fc2 -> *gamma -> +shortcut for C <= 160 (stages 0/1, where the
// hidden tensor would otherwise be 4x the activation traffic).
//
// Per CTA (persistent over 128-row tiles), hidden processed in chunks of 64 columns:
//   warp 0      TMA producer : y tile (KB1 k-blocks of [128x64]), then per chunk W1_j ([64 x C]) and W2_j ([C x 64])
//   warp 1      MMA issuer   : G1_j: D1[j%2] = y . W1_j^T          (M128 x N64, K = C)
//                              G2_j: D2     += H[j%2] . W2_j^T      (M128 x N=C, K = 64), software-pipelined one
//                              chunk behind G1 so the tensor pipe never waits for the GELU warps
//   warps 2..17 epilogue     : D1 chunk: tcgen05.ld -> +b1 -> GELU -> bf16 -> H[j%2] in the SW128 K-major layout
//                              UMMA reads; after the last chunk: D2 -> +b2 -> *gamma + res -> bf16 -> global
#include <stdlib.h>

#include "tc_common.cuh"

namespace btsb {
namespace {
constexpr int FM = 128;            // rows per tile
constexpr int NH = 64;             // hidden chunk
constexpr int kWStages = 3;
constexpr int kEpiWarpsF = 16;
constexpr int kThreadsF = 64 + kEpiWarpsF * 32;
constexpr int kYBlockBytes = FM * 128;     // one [128 x 64] bf16 k-block
constexpr int kHBytes = FM * 128;          // [128 x 64] bf16
constexpr int kD2Col = 2 * NH;             // TMEM column of D2 (D1 buffers occupy [0,128))

struct FusedLayout {
  int kb1;            // k-blocks of the fc1 contraction (ceil(C/64))
  int ny;             // y tile buffers (1 or 2)
  int w_stage_bytes;  // kb1 * 8 KB + C * 128 B
  int y_bytes;        // kb1 * 16 KB
  int off_w, off_h, off_bar, off_b1, total;
};
__host__ __device__ inline FusedLayout fused_layout(int C) {
  FusedLayout L;
  L.kb1 = (C + 63) / 64;
  L.w_stage_bytes = L.kb1 * (NH * 128) + ((C * 128 + 1023) / 1024) * 1024;
  L.y_bytes = L.kb1 * kYBlockBytes;
  const int fixed = kWStages * L.w_stage_bytes + 2 * kHBytes + 1024 + 512 + 4 * C * 4;
  L.ny = (fixed + 2 * L.y_bytes <= 227 * 1024) ? 2 : 1;
  L.off_w = L.ny * L.y_bytes;
  L.off_h = L.off_w + kWStages * L.w_stage_bytes;
  L.off_bar = L.off_h + 2 * kHBytes;
  L.off_b1 = L.off_bar + 512;
  L.total = L.off_b1 + 4 * C * 4 + 1024;
  return L;
}
}  // namespace

__global__ void __launch_bounds__(kThreadsF, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const float* __restrict__ b1,
                 const float* __restrict__ b2, const float* __restrict__ gamma,
                 const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ out, int M, int C) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const FusedLayout L = fused_layout(C);
  const int NJ = (4 * C) / NH;
  const int m_tiles = (M + FM - 1) / FM;
  const uint32_t bar0 = sbase + L.off_bar;
  // barrier indices
  auto y_full = [&](int i) { return bar0 + 8u * i; };
  auto y_empty = [&](int i) { return bar0 + 8u * (2 + i); };
  auto w_full = [&](int i) { return bar0 + 8u * (4 + i); };
  auto w_empty = [&](int i) { return bar0 + 8u * (4 + kWStages + i); };
  auto d1_full = [&](int i) { return bar0 + 8u * (4 + 2 * kWStages + i); };
  auto d1_empty = [&](int i) { return bar0 + 8u * (6 + 2 * kWStages + i); };
  auto h_full = [&](int i) { return bar0 + 8u * (8 + 2 * kWStages + i); };
  auto h_empty = [&](int i) { return bar0 + 8u * (10 + 2 * kWStages + i); };
  auto d2_full = [&](int i) { return bar0 + 8u * (12 + 2 * kWStages + i); };
  auto d2_empty = [&](int i) { return bar0 + 8u * (14 + 2 * kWStages + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + L.off_bar + 8 * (16 + 2 * kWStages));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int i = 0; i < 2; ++i) { mbar_init(y_full(i), 1); mbar_init(y_empty(i), 1); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(w_full(i), 1); mbar_init(w_empty(i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full(i), 1); mbar_init(d1_empty(i), kEpiWarpsF);
      mbar_init(h_full(i), kEpiWarpsF); mbar_init(h_empty(i), 1);
      mbar_init(d2_full(i), 1); mbar_init(d2_empty(i), kEpiWarpsF);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 512); tmem_relinquish(); }
  float* b1s = reinterpret_cast<float*>(sal + L.off_b1);            // fc1 bias staged once per CTA
  for (int i = threadIdx.x; i < 4 * C; i += kThreadsF) b1s[i] = __ldg(b1 + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    // ============================== TMA producer ==============================
    uint32_t ycount = 0, wcount = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      const int yb = ycount % L.ny; const uint32_t yph = (ycount / L.ny) & 1u; ++ycount;
      mbar_wait(y_empty(yb), yph ^ 1u);
      mbar_expect_tx(y_full(yb), (uint32_t)L.y_bytes);
      for (int kb = 0; kb < L.kb1; ++kb)
        tma_load_2d(sbase + yb * L.y_bytes + kb * kYBlockBytes, &tmY, y_full(yb), kb * 64, tile * FM);
      for (int j = 0; j < NJ; ++j) {
        const int ws = wcount % kWStages; const uint32_t wph = (wcount / kWStages) & 1u; ++wcount;
        mbar_wait(w_empty(ws), wph ^ 1u);
        mbar_expect_tx(w_full(ws), (uint32_t)(L.kb1 * NH * 128 + C * 128));
        const uint32_t wb = sbase + L.off_w + ws * L.w_stage_bytes;
        for (int kb = 0; kb < L.kb1; ++kb) tma_load_2d(wb + kb * (NH * 128), &tmW1, w_full(ws), kb * 64, j * NH);
        tma_load_2d(wb + L.kb1 * (NH * 128), &tmW2, w_full(ws), j * NH, 0);
      }
    }
  } else if (threadIdx.x == 32) {
    // ============================== MMA issuer ==============================
    // One global chunk sequence over all tiles of this CTA: G1 of chunk g is issued before G2 of chunk g-1, also across
    // tile boundaries (D2 is double buffered), so the tensor pipe never idles behind the epilogue's last chunk.
    uint32_t nt = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) ++nt;
    const uint32_t total = nt * (uint32_t)NJ;
    const uint32_t idesc1 = idesc_bf16_f32(FM, NH);
    const uint32_t idesc2 = idesc_bf16_f32(FM, C);
    auto do_g1 = [&](uint32_t g) {
      const uint32_t ti = g / (uint32_t)NJ; const int j = (int)(g - ti * NJ);
      const int yb = (int)(ti % (uint32_t)L.ny); const uint32_t yph = (ti / (uint32_t)L.ny) & 1u;
      if (j == 0) mbar_wait(y_full(yb), yph);
      const int ws = (int)(g % kWStages); const uint32_t wph = (g / kWStages) & 1u;
      const int b = (int)(g & 1u); const uint32_t bph = (g >> 1) & 1u;
      mbar_wait(w_full(ws), wph);
      mbar_wait(d1_empty(b), bph ^ 1u);
      tc_fence_after();
      const uint32_t ybase = sbase + yb * L.y_bytes;
      const uint32_t wb = sbase + L.off_w + ws * L.w_stage_bytes;
      for (int kb = 0; kb < L.kb1; ++kb) {
        const uint64_t ad = smem_desc_sw128(ybase + kb * kYBlockBytes);
        const uint64_t bd = smem_desc_sw128(wb + kb * (NH * 128));
        const int kmax = min(64, C - kb * 64) / 16;
        for (int kk = 0; kk < kmax; ++kk)
          umma_bf16(tmem_base + (uint32_t)(b * NH), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc1,
                    (kb | kk) != 0 ? 1u : 0u);
      }
      umma_commit(d1_full(b));
      if (j == NJ - 1) umma_commit(y_empty(yb));               // y tile no longer needed once these retire
    };
    auto do_g2 = [&](uint32_t g) {
      const uint32_t ti = g / (uint32_t)NJ; const int j = (int)(g - ti * NJ);
      const int tb = (int)(ti & 1u); const uint32_t tph = (ti >> 1) & 1u;
      const int ws = (int)(g % kWStages);
      const int b = (int)(g & 1u); const uint32_t bph = (g >> 1) & 1u;
      mbar_wait(h_full(b), bph);
      if (j == 0) mbar_wait(d2_empty(tb), tph ^ 1u);
      tc_fence_after();
      const uint32_t wb = sbase + L.off_w + ws * L.w_stage_bytes + L.kb1 * (NH * 128);
      const uint64_t ad = smem_desc_sw128(sbase + L.off_h + b * kHBytes);
      const uint64_t bd = smem_desc_sw128(wb);
      for (int kk = 0; kk < NH / 16; ++kk)
        umma_bf16(tmem_base + (uint32_t)(kD2Col + tb * C), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc2,
                  (j | kk) != 0 ? 1u : 0u);
      umma_commit(h_empty(b));
      umma_commit(w_empty(ws));
      if (j == NJ - 1) umma_commit(d2_full(tb));
    };
    for (uint32_t g = 0; g <= total; ++g) {
      const bool g1 = g < total, g2 = g >= 1;
      // with a single y buffer the next tile's y load only starts after this tile's last G1 retires: do not let
      // the wait for it delay G2 of the previous chunk
      if (g1 && g2 && L.ny == 1 && (g % (uint32_t)NJ) == 0) { do_g2(g - 1); do_g1(g); }
      else { if (g1) do_g1(g); if (g2) do_g2(g - 1); }
    }
  } else if (warp >= 2) {
    // ============================== epilogue warps ==============================
    const int q = warp & 3;                 // TMEM lane quarter
    const int s = (warp - 2) >> 2;          // 16-column slice of a 64-column chunk / D2 column-group stride, 0..3
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int groups2 = C / 16;
    uint32_t ccount = 0, ti = 0;
    uint4 rpre[3][2];                       // residual rows of the tile whose D2 epilogue is pending

    auto prefetch_res = [&](int tile) {
      const int row = tile * FM + r_in_tile;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int gi = s + 4 * u;
        if (gi < groups2 && row < M) {
          const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)row * C + gi * 16);
          rpre[u][0] = __ldg(rp); rpre[u][1] = __ldg(rp + 1);
        } else {
          rpre[u][0] = make_uint4(0, 0, 0, 0); rpre[u][1] = make_uint4(0, 0, 0, 0);
        }
      }
    };
    // D2 -> +b2 -> *gamma + res -> bf16 rows of tile `tile` (local index tl)
    auto d2_epilogue = [&](int tile, uint32_t tl) {
      const int tb = (int)(tl & 1u);
      mbar_wait(d2_full(tb), (tl >> 1) & 1u);
      tc_fence_after();
      const int row = tile * FM + r_in_tile;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int gi = s + 4 * u;
        if (gi >= groups2) break;
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C + gi * 16), r);
        tmem_ld_wait();
        if (gi + 4 >= groups2) {                         // last D2 read of this warp for this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty(tb));
        }
        if (row < M) {
          const int n = gi * 16;
          const uint4 r0 = rpre[u][0], r1 = rpre[u][1];
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n + i));
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n + i));
            v[i] = fmaf(g4.x, __uint_as_float(r[i]) + b4.x, bf16_lo(rr[i / 2]));
            v[i + 1] = fmaf(g4.y, __uint_as_float(r[i + 1]) + b4.y, bf16_hi(rr[i / 2]));
            v[i + 2] = fmaf(g4.z, __uint_as_float(r[i + 2]) + b4.z, bf16_lo(rr[i / 2 + 1]));
            v[i + 3] = fmaf(g4.w, __uint_as_float(r[i + 3]) + b4.w, bf16_hi(rr[i / 2 + 1]));
          }
          uint4 o0, o1;
          o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
          o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
          o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
          o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)row * C + n);
          op[0] = o0; op[1] = o1;
        }
      }
    };

    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++ti) {
      // the previous tile's D2 epilogue is deferred until after this tile's first chunk, so the tail latency
      // (last H hand-off -> G2 -> commit) is hidden behind GELU work; its residual rows are fetched now
      if (prev_tile >= 0) prefetch_res(prev_tile);
      for (int j = 0; j < NJ; ++j, ++ccount) {
        const int b = (int)(ccount & 1u); const uint32_t bph = (ccount >> 1) & 1u;
        mbar_wait(d1_full(b), bph);
        tc_fence_after();
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)(b * NH + s * 16), r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty(b));         // D1[b] may be overwritten by G1 of chunk j+2
        const int hcol = j * NH + s * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(b1s + hcol + i);
          v[i] = gelu_fast(__uint_as_float(r[i]) + b4.x);
          v[i + 1] = gelu_fast(__uint_as_float(r[i + 1]) + b4.y);
          v[i + 2] = gelu_fast(__uint_as_float(r[i + 2]) + b4.z);
          v[i + 3] = gelu_fast(__uint_as_float(r[i + 3]) + b4.w);
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
        o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
        o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
        o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
        mbar_wait(h_empty(b), bph ^ 1u);                 // G2 of chunk j-2 has finished reading H[b]
        unsigned char* hb = sal + L.off_h + b * kHBytes;
        *reinterpret_cast<uint4*>(hb + sw128_offset(r_in_tile, s * 16)) = o0;
        *reinterpret_cast<uint4*>(hb + sw128_offset(r_in_tile, s * 16 + 8)) = o1;
        fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(h_full(b));
        if (j == 0 && prev_tile >= 0) d2_epilogue(prev_tile, ti - 1);
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      prefetch_res(prev_tile);
      d2_epilogue(prev_tile, ti - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int num_sms();
int mlp_fused2_supported(int C);
int mlp_fused2_launch(const void* y, const void* res, const void* W1, const float* b1, const void* W2, const float* b2,
                      const float* gamma, void* out, int64_t M, int C, cudaStream_t st);

// BTSB_MLP_V1=1 selects the first-generation kernel in this file (kept for A/B timing against mlp_fused2_tc.cu)
static bool use_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BTSB_MLP_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

}  // namespace btsb

using namespace btsb;

extern "C" int btsb_convnext_mlp_fused_fwd(const void* y, const void* res, const void* W1, const float* b1,
                                           const void* W2, const float* b2, const float* gamma, void* out, int64_t M,
                                           int C, void* stream) {
  if (int e = check_device()) return e;
  BTSB_REQUIRE(M >= 0 && M < (1ll << 31), "mlp_fused: bad M");
  BTSB_REQUIRE(mlp_fused2_supported(C), "mlp_fused: C=%d unsupported (multiple of 16 in [64,160], 256 or 320)", C);
  if (M == 0) return BTSB_OK;
  BTSB_REQUIRE(y && res && W1 && b1 && W2 && b2 && gamma && out, "mlp_fused: null pointer");
  BTSB_REQUIRE(((uintptr_t)res % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)b1 % 16) == 0 &&
                   ((uintptr_t)b2 % 16) == 0 && ((uintptr_t)gamma % 16) == 0,
               "mlp_fused: pointers must be 16-byte aligned");
  if (!use_v1() && mlp_fused2_supported(C))
    return mlp_fused2_launch(y, res, W1, b1, W2, b2, gamma, out, M, C, (cudaStream_t)stream);
  CUtensorMap tmY, tmW1, tmW2;
  if (int e = make_tmap_bf16_2d(&tmY, y, (uint64_t)M, (uint64_t)C, FM)) return e;
  if (int e = make_tmap_bf16_2d(&tmW1, W1, (uint64_t)(4 * C), (uint64_t)C, NH)) return e;
  if (int e = make_tmap_bf16_2d(&tmW2, W2, (uint64_t)C, (uint64_t)(4 * C), (uint32_t)C)) return e;
  const FusedLayout L = fused_layout(C);
  BTSB_REQUIRE(L.total <= 227 * 1024, "mlp_fused: shared-memory plan %d B exceeds 227 KB", L.total);
  BTSB_CUDA(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "mlp_fused attr");
  const int m_tiles = (int)((M + FM - 1) / FM);
  const int grid = min(m_tiles, num_sms());
  mlp_fused_kernel<<<grid, kThreadsF, L.total, (cudaStream_t)stream>>>(
      tmY, tmW1, tmW2, b1, b2, gamma, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, (int)M, C);
  return launch_done("mlp_fused");
}
