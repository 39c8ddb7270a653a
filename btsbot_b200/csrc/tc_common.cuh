// sm_100a building blocks shared by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld) and the UMMA shared-memory + instruction descriptors.  Inline PTX only.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace btsb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);                        // back off: polling warps must not steal issue slots from the math warps
    if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
      printf("btsbot_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Wait without the nanosleep back-off: try_wait already suspends the thread in hardware for a short, system-defined
// interval, and a 64 ns sleep quantum is a visible fraction of a sub-microsecond pipeline step.  Still bounded: a protocol
// bug traps instead of hanging the GPU.  (Parking the warp with a suspend-time hint -- try_wait ..., 20000 ns, which
// compiles to TRYWAIT + NANOSLEEP.SYNCS -- removes the polling instructions, ~20 % of everything the fused MLP issues,
// but wakes later: mlp_fused_80 went from 318 to 352 us, the other kernels did not move; profiles/r02o.  Not kept.)
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) {
        printf("btsbot_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}

// one lane of a converged warp (the single-thread TMA / MMA roles run warp-uniform control flow and predicate only the
// issue on the elected lane, so UTMALDG / UTCHMMA compile without per-lane serialisation loops)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// bulk tensor store shared -> global (tracked by the per-thread bulk async-group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// bulk tensor REDUCTION shared -> global: global[box] += smem[box], element type from the tensor map (bf16 here), performed
// at the L2; completion tracked like a bulk store
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T ; kind::f16 covers bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all prior tcgen05.mma of this thread arrive (once) on `bar` when they retire; implies fence::before_thread_sync
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of one cluster issue ONE UMMA of M = 256 -------------------------------------
// Rank 0 (the leader) issues the MMAs; each CTA stages its own 128 rows of A and HALF of the B tile, and the hardware
// reads both halves -- per CTA that is 8 KB instead of 12 KB of operand reads (and TMA writes) per 128x256x16 UMMA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in THIS CTA's shared memory, the bytes are counted on the barrier at cluster
// address `bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all prior pair-MMAs of this thread arrive on the barrier at this shared-memory offset in BOTH CTAs when they retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers (thread = TMEM lane, register i = column i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layouts) -----------------------------------------------
// K-major operand tile in shared memory, 128-byte swizzle: rows are 128 B (64 bf16), 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows * 128 B      [32,46)
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)           [46,48)
  d |= (uint64_t)2 << 61;                     // layout type SWIZZLE_128B                 [61,64)
  return d;
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row, col) inside a [rows x 64] bf16 SW128 K-major tile (what TMA writes / UMMA reads)
__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
  const int chunk = (col >> 3) ^ (row & 7);
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + chunk * 16 + (col & 7) * 2);
}

// GELU for bf16 outputs: 0.5 x (1 + erf(x/sqrt2)) == 0.5 x (1 + tanh(p(x))) with the odd quintic
// p(x) = x (0.7975078843 + 0.0370056460 x^2 - 0.000351516790 x^4) fitted (minimax over |x| <= 7) to atanh(erf(x/sqrt2)):
// max |error| of the formula = 2.5e-5.  ONE MUFU op (tanh.approx, rel. error 2^-11) + 7 FMA-pipe ops: measured
// 10.3 clk per warp-element and SMSP against 17.1 for the ex2 + rcp sigmoid form (profiles/r01d_micro_gelu.txt) -- the
// GELU epilogues are MUFU-bound (16 lanes/clk/SM), so this is a 1.66x higher ceiling.  The tanh error enters as
// 0.5 |x| 2^-11 absolute (<= 7e-4 at |x| = 3), below the bf16 rounding (2^-9 relative) of the hidden activations it is
// stored into.  The fp32 path keeps erff() (common.cuh gelu_erf).
__device__ __forceinline__ float gelu_fast(float x) {
  // the quintic is only monotone on |x| < 11: clamp x^2 (one op); beyond |x| = 8 the slope is frozen at p(8)/8 > 0, so
  // tanh still saturates to -1 / +1
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(x2, -3.5151679e-4f, 0.037005646f);
  p = fmaf(x2, p, 0.7975078843f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

// Two GELUs per call on packed fp32 pairs (sm_100 FMUL2 / FFMA2): the scalar form issues 9-10 instructions per
// element and is issue-bound at 10.1 clk per warp-element and SMSP, just above the MUFU floor of 8.1
// (profiles/r01d_micro_gelu.txt); packed, the FMA-pipe part is 3 instructions per element.
// (Evaluating both tanh with one tanh.approx.f16x2 was tried twice: round 1 346 vs 317 us at C = 80; again after the D2
// epilogue left the GELU warps and the MUFU pipe read 65 % busy: 280 vs 247 us (profiles/r02h2) -- the pack and the two cvt
// cost more issue slots than the second MUFU costs pipe time.  Not kept.)
__device__ __forceinline__ f32x2_t gelu_fast2(f32x2_t x) {
  const f32x2_t k5 = pack_f32x2(-3.5151679e-4f, -3.5151679e-4f), k3 = pack_f32x2(0.037005646f, 0.037005646f);
  const f32x2_t k1 = pack_f32x2(0.7975078843f, 0.7975078843f), kh = pack_f32x2(0.5f, 0.5f);
  const float2 sq = unpack_f32x2(mul_f32x2(x, x));
  const f32x2_t x2 = pack_f32x2(fminf(sq.x, 64.0f), fminf(sq.y, 64.0f));
  f32x2_t p = fma3_f32x2(x2, k5, k3);
  p = fma3_f32x2(x2, p, k1);
  const float2 a = unpack_f32x2(mul_f32x2(p, x));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a.y));
  const f32x2_t h = mul_f32x2(x, kh);
  return fma3_f32x2(h, pack_f32x2(t0, t1), h);
}

// 2 x GELU for the fused MLP's hidden activations (the factor 1/2 moves into the D2 epilogue's FFMA, where it is free):
// 2 GELU(x) = x + x tanh(x (a + b x^2)) with the odd CUBIC fitted (minimax over |x| <= 9) to the erf form: max |error| of
// GELU itself 2.7e-4 -- above the quintic's 2.5e-5 but below what tanh.approx already contributes (0.5 |x| 2^-11 = 7e-4 at
// |x| = 3) and a tenth of the bf16 rounding of the value it is stored as.  The cubic is monotone, so the clamp of the
// quintic form (two unpacked FMNMX per pair) and one FFMA2 go away, and with the halving folded downstream a pair costs
// 4 packed FMA-pipe instructions + 2 MUFU instead of 8 + 2: the fused MLP's GELU warps are issue-bound (ncu r01m: issue
// slots 73 %, MUFU 45 %, tensor pipe 28 % at C = 80).
__device__ __forceinline__ f32x2_t gelu_twice2(f32x2_t x) {
  const f32x2_t k3 = pack_f32x2(0.03470089f, 0.03470089f), k1 = pack_f32x2(0.80015708f, 0.80015708f);
  const f32x2_t p = fma3_f32x2(mul_f32x2(x, x), k3, k1);
  const float2 a = unpack_f32x2(mul_f32x2(p, x));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a.y));
  return fma3_f32x2(x, pack_f32x2(t0, t1), x);
}

// The quintic form for the fused MLP, bit-identical to gelu_fast2 up to a power of two: the caller supplies x8 = x / 8 (its
// bias FFMA2 does the scaling: x8 = acc * 0.125 + b1 / 8), the clamp min(x^2, 64) becomes u = sat(x8 * x8) through the
// saturating FMUL (one instruction per element instead of the packed square + two unpacked FMNMX), the coefficients
// absorb the powers of two, and the result is GELU(x) / 4 -- the D2 epilogue's FFMA multiplies the accumulator by 4.
// 8 instructions per pair instead of 10.
__device__ __forceinline__ f32x2_t gelu_quarter_quintic2(f32x2_t x8) {
  const f32x2_t k5 = pack_f32x2(-3.5151679e-4f * 4096.0f * 8.0f, -3.5151679e-4f * 4096.0f * 8.0f);
  const f32x2_t k3 = pack_f32x2(0.037005646f * 64.0f * 8.0f, 0.037005646f * 64.0f * 8.0f);
  const f32x2_t k1 = pack_f32x2(0.7975078843f * 8.0f, 0.7975078843f * 8.0f);
  const float2 xs = unpack_f32x2(x8);
  float u0, u1;
  asm("mul.rn.sat.f32 %0, %1, %1;" : "=f"(u0) : "f"(xs.x));
  asm("mul.rn.sat.f32 %0, %1, %1;" : "=f"(u1) : "f"(xs.y));
  const f32x2_t u = pack_f32x2(u0, u1);
  f32x2_t p = fma3_f32x2(u, k5, k3);
  p = fma3_f32x2(u, p, k1);
  const float2 a = unpack_f32x2(mul_f32x2(p, x8));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a.y));
  return fma3_f32x2(x8, pack_f32x2(t0, t1), x8);
}

// SiLU x * sigmoid(x) == 0.5 x (1 + tanh(x / 2)) for bf16 outputs: one MUFU op
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace tc

// host: encode a 2-D bf16 row-major [rows, cols] tensor map with a [box_rows x 64] box, 128B swizzle
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

}  // namespace btsb
