// MaxViT 7x7 window / grid attention on tensor cores (bf16 path).  One warp per (window, head): the 49 x 32 q, k, v
// tiles are gathered from the image-order qkv rows into shared memory (the partition is index arithmetic), then
//   S = q k^T            mma.sync m16n8k16 bf16 (4 m-tiles x 7 n-tiles x 2 k-steps), operands via ldmatrix
//   P = softmax(S/sqrt(32) + rel-pos bias)   in the accumulator registers (exp2, quad shuffles)
//   O = P v              P re-used straight from the S accumulators as the A fragments, v via ldmatrix.trans
// and the 49 x 32 output tile is staged in shared memory and scattered back to image order in 16-byte pieces.
// Why mma.sync and not tcgen05 here: a (window, head) problem is 49x49x32 -- a single M=64 UMMA tile with 23 % of it
// padding, whose operands would still have to be gathered row by row; the kernel is bound by that gather, not by math.
#include "common.cuh"

namespace btsb {
namespace {

constexpr int kWin = 7, kTok = 49, kDh = 32;
constexpr int kRow = 40;                 // smem row stride in bf16 (80 B): ldmatrix rows hit 8 distinct 16-byte bank groups
constexpr int kWarps = 4;
constexpr int kQRows = 64, kKRows = 56, kVRows = 64;
constexpr int kWarpSmem = (kQRows + kKRows + kVRows) * kRow * 2 + 176 * 4;     // + bias table (169 floats, padded)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-byte asynchronous copy global -> shared (L2 only): the whole (window, head) gather is in flight at once instead of
// one load round trip per 32 pieces (the kernel ran at 1.2 TB/s with 12 resident warps waiting on ~19 dependent trips)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kWarps * 32)
mv_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int H, int W, int C, int heads,
                   int grid_mode, const float* __restrict__ table, int64_t nitems) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * kWarps + warp;
  if (item >= nitems) return;                      // warps are independent: no CTA-wide barrier below
  unsigned char* base = smem + warp * kWarpSmem;
  __nv_bfloat16* qs = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* ks = qs + kQRows * kRow;
  __nv_bfloat16* vs = ks + kKRows * kRow;
  float* tb = reinterpret_cast<float*>(vs + kVRows * kRow);

  const int h = (int)(item % heads);
  int64_t win = item / heads;
  const int nwx = W / kWin, nwy = H / kWin;
  const int wx = (int)(win % nwx); win /= nwx;
  const int wy = (int)(win % nwy);
  const int64_t b = win / nwy;
  auto token_row = [&](int t) -> int64_t {
    const int ty = t / kWin, tx = t - ty * kWin;
    const int y = grid_mode ? ty * nwy + wy : wy * kWin + ty;
    const int x = grid_mode ? tx * nwx + wx : wx * kWin + tx;
    return (b * H + y) * (int64_t)W + x;
  };
  constexpr float kLog2e = 1.4426950408889634f;
  // gather: 49 tokens x 12 pieces of 16 B (q: 4, k: 4, v: 4), all issued before anything is waited for
  for (int i = lane; i < kTok * 12; i += 32) {
    const int t = i / 12, piece = i - t * 12;
    __nv_bfloat16* dst = (piece < 4 ? qs : piece < 8 ? ks : vs) + t * kRow + (piece & 3) * 8;
    cp_async16(dst, reinterpret_cast<const uint4*>(qkv + token_row(t) * 3 * C + h * 3 * kDh) + piece);
  }
  for (int i = lane; i < 169; i += 32) tb[i] = table[i * heads + h] * kLog2e;
  // zero the padding rows (q 49..63, k 49..55, v 49..63): P is exactly 0 there, but 0 * NaN garbage would poison O
  for (int i = lane; i < (15 + 7 + 15) * 4; i += 32) {
    const int r = i >> 2, piece = i & 3;
    __nv_bfloat16* dst = r < 15 ? qs + (kTok + r) * kRow : r < 22 ? ks + (kTok + r - 15) * kRow : vs + (kTok + r - 22) * kRow;
    *reinterpret_cast<uint4*>(dst + piece * 8) = make_uint4(0, 0, 0, 0);
  }
  cp_async_wait_all();
  __syncwarp();

  const int g = lane >> 2, t4 = lane & 3;
  const float sc = 0.17677669529663687f * kLog2e;          // dim_head^-0.5, in the exp2 domain
  const uint32_t qs_u = smem_u32(qs), ks_u = smem_u32(ks), vs_u = smem_u32(vs);
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    uint32_t aq[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      ldsm_x4(qs_u + (uint32_t)(((mt * 16 + (lane & 15)) * kRow + kk * 16 + (lane >> 4) * 8) * 2), aq[kk][0], aq[kk][1],
              aq[kk][2], aq[kk][3]);
    float s[7][4];
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(ks_u + (uint32_t)(((nt * 8 + (lane & 7)) * kRow + (lane >> 3) * 8) * 2), b0, b1, b2, b3);
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      mma_bf16(s[nt], aq[0], b0, b1);
      mma_bf16(s[nt], aq[1], b2, b3);
    }
    // scale + relative-position bias + mask, row max
    const int r0 = min(mt * 16 + g, kTok - 1), r1 = min(mt * 16 + g + 8, kTok - 1);
    const int io0 = (r0 / kWin) * 13 + r0 % kWin + 84, io1 = (r1 / kWin) * 13 + r1 % kWin + 84;
    float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = nt * 8 + 2 * t4 + e;
        const int jo = (c / kWin) * 13 + c % kWin;
        const bool ok = c < kTok;
        const float v0 = ok ? fmaf(s[nt][e], sc, tb[ok ? io0 - jo : 0]) : -3.0e38f;
        const float v1 = ok ? fmaf(s[nt][2 + e], sc, tb[ok ? io1 - jo : 0]) : -3.0e38f;
        s[nt][e] = v0; s[nt][2 + e] = v1;
        m0 = fmaxf(m0, v0); m1 = fmaxf(m1, v1);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
      s[nt][0] = ex2(s[nt][0] - m0); s[nt][1] = ex2(s[nt][1] - m0);
      s[nt][2] = ex2(s[nt][2] - m1); s[nt][3] = ex2(s[nt][3] - m1);
      d0 += s[nt][0] + s[nt][1]; d1 += s[nt][2] + s[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // O = P v : P's accumulator layout is the A-fragment layout (two adjacent 8-column tiles = one 16-wide k-step)
    uint32_t ap[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ap[kk][0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      ap[kk][1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      if (2 * kk + 1 < 7) {
        ap[kk][2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ap[kk][3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      } else {
        ap[kk][2] = 0u; ap[kk][3] = 0u;               // tokens 56..63 do not exist
      }
    }
    float o[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {                // 32 tokens per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vs_u + (uint32_t)(((k2 * 32 + (lane >> 3) * 8 + (lane & 7)) * kRow + dt * 8) * 2), b0, b1, b2, b3);
        mma_bf16(o[dt], ap[2 * k2], b0, b1);
        mma_bf16(o[dt], ap[2 * k2 + 1], b2, b3);
      }
    }
    const float i0 = 1.0f / d0, i1 = 1.0f / d1;
    __syncwarp();                                      // every lane holds its q fragments: the q rows can be overwritten
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      *reinterpret_cast<uint32_t*>(qs + (mt * 16 + g) * kRow + dt * 8 + 2 * t4) = pack2(o[dt][0] * i0, o[dt][1] * i0);
      *reinterpret_cast<uint32_t*>(qs + (mt * 16 + g + 8) * kRow + dt * 8 + 2 * t4) = pack2(o[dt][2] * i1, o[dt][3] * i1);
    }
  }
  __syncwarp();
  for (int i = lane; i < kTok * 4; i += 32) {
    const int t = i >> 2, piece = i & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(qs + t * kRow + piece * 8);
    *reinterpret_cast<uint4*>(out + token_row(t) * C + h * kDh + piece * 8) = v;
  }
}

}  // namespace

int maxvit_attn_bf16_mma(const void* qkv, void* out, int64_t B, int H, int W, int C, int grid_mode, const float* table,
                         cudaStream_t st) {
  const int heads = C / kDh;
  const int64_t nitems = B * (H / kWin) * (W / kWin) * heads;
  BTSB_REQUIRE(nitems / kWarps < (1ll << 31) - 1, "maxvit attn: too many windows");
  BTSB_REQUIRE(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "maxvit attn: qkv/out must be 16-byte aligned");
  static bool attr_done = false;
  if (!attr_done) {
    BTSB_CUDA(cudaFuncSetAttribute(mv_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kWarpSmem), "attn attr");
    attr_done = true;
  }
  const unsigned grid = (unsigned)((nitems + kWarps - 1) / kWarps);
  mv_attn_mma_kernel<<<grid, kWarps * 32, kWarps * kWarpSmem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, H, W, C,
                                                                   heads, grid_mode, table, nitems);
  return launch_done("maxvit_attn_mma");
}

}  // namespace btsb
