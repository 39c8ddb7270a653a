// K4 fused, second generation (bf16): out = res + gamma * (fc2(GELU(fc1(y) + b1)) + b2) in ONE kernel; the 4C-wide
// hidden activation lives only in TMEM / shared memory.  Replaces timm blocks.j.mlp.fc1 -> act -> mlp.fc2 -> *gamma ->
// +shortcut (and MaxViT's token MLPs) for C <= 160.
//
// Design points (profiles/r01c: with one 3-stage weight ring 20-28 % of all epilogue-warp samples sat on the D1 barrier
// because the TMA producer had only ~1 chunk of look-ahead against ~1 us of L2 latency):
//   * weights RESIDENT in shared memory when they fit (C <= 96: W1 + W2 = 8 C^2 * 2 B <= 147 KB), loaded once per CTA;
//     otherwise two independent rings (W1 slots are released as soon as G1 retires, W2 slots after G2) so both
//     streams run 3-4 hidden chunks ahead of the tensor pipe
//   * K is tiled as 64-column SW128 blocks plus compact 32-column (SW64) / 16-column (SW32) tail blocks instead of
//     zero-padded 64-column blocks: C = 80 moves 20 KB per y tile instead of 32 KB
//   * the 16 epilogue warps form two groups that own alternate hidden chunks (32 columns x 32 rows per warp and
//     chunk): half as many barrier round trips / proxy fences per GELU, and the two groups run out of phase so
//     the MUFU-bound part of one overlaps the TMEM-load / FMA / store part of the other
//   * G1 of chunk g+2 is issued as soon as the epilogue has pulled chunk g out of TMEM (not after its GELU)
//
// Per CTA (persistent over 128-row tiles), hidden processed in chunks of 64 columns, global chunk index g:
//   warp 0      TMA producer : y tiles, W1 chunk [64 x C], W2 chunk [C x 64]
//   warp 1      MMA issuer 1 : G1_g: D1[g&1] = y . W1_g^T (M128 x N64, K = C)
//   warp 18     MMA issuer 2 : G2_g: D2[tile&1] += H[g&1] . W2_g^T     (one warp per stream: see the issuer comment)
//   warps 2..17 epilogue     : group g&1: tcgen05.ld D1 -> +b1 -> GELU -> bf16 -> H[g&1] (SW128 K-major, what UMMA
//                              reads); all 16 warps: D2 -> +b2 -> *gamma + res -> bf16 rows (deferred by one chunk)
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace btsb {
namespace {
constexpr int FM = 128;            // rows per tile
constexpr int NH = 64;             // hidden chunk
constexpr int kEpiWarps2 = 16;
constexpr int kG2Warp = 2 + kEpiWarps2;                 // second MMA-issuing warp (G2 stream)
constexpr int kThreads2 = 64 + kEpiWarps2 * 32 + 32;
// Narrow kernels (C <= 160, two D2 accumulators): four more warps, one per TMEM lane quarter, do nothing but the D2
// epilogue.  With the drain on the GELU warps both warp groups left the GELU loop at every tile boundary for ~1.1-1.3 k
// clocks, and the MUFU pipe -- the resource the C = 80 kernel is on (a warp-wide tanh.approx holds it for 8 clocks; 2560 of
// a tile's ~4950 clocks busy) -- idled meanwhile (profiles/r02t/mlp_trace_80.txt, DESIGN.md lesson 22).
constexpr int kD2Warps = 4;
constexpr int kThreadsDW = kThreads2 + kD2Warps * 32;   // 23 warps: 88 registers per thread

constexpr int kHBytes = FM * 128;  // [128 x 64] bf16
constexpr int kD2Col = 2 * NH;     // TMEM column of D2 (D1 buffers occupy [0,128))
constexpr int kMaxSlots = 16;
constexpr int kSmemMax = 227 * 1024;

struct Plan2 {
  int C = 0, NJ = 0;
  int nfull = 0, t32 = 0, t16 = 0;   // K blocks: nfull x 64 columns (SW128) [+ 32 columns (SW64)] [+ 16 columns (SW32)]
  int y_bytes = 0, w1_bytes = 0, w2_bytes = 0;
  int ny = 0, n1 = 0, n2 = 0, resident = 0;
  int off_w1 = 0, off_w2 = 0, off_h = 0, off_stg = 0, off_slab = 0, off_bar = 0, off_b1 = 0, total = 0;
  int nstg = 1;              // staging tiles (2 for C <= 80: the residual rows of tile t+1 land while tile t drains)
  int slab = 0;              // 1: one 1 KB slab per epilogue warp ([32 rows x 16 columns] bulk tensor copies, wide C)
  int stg = 0;               // 1: residual / output rows move through a [128 x C] bf16 staging tile with bulk tensor copies
  int ht = 0;                // 1: the GELU'd hidden chunk H lives in TMEM (A operand of G2 from TMEM), no shared-memory H tiles
  bool ok = false;
};
__host__ __device__ constexpr Plan2 plan2_for(int C, bool te, bool ht, bool slab = false);

__host__ __device__ constexpr int rup1k(int x) { return (x + 1023) & ~1023; }

__host__ __device__ constexpr bool plan2_try(Plan2& P, int ny, int n1, int n2, int resident) {
  P.ny = ny; P.n1 = n1; P.n2 = n2; P.resident = resident;
  P.off_w1 = ny * P.y_bytes;
  P.off_w2 = P.off_w1 + n1 * P.w1_bytes;
  P.off_h = P.off_w2 + n2 * P.w2_bytes;
  P.off_stg = P.off_h + (P.ht ? 0 : 2 * kHBytes);
  P.off_slab = P.off_stg + (P.stg ? P.nstg * P.C * 256 : 0);
  P.off_bar = P.off_slab + (P.slab ? kEpiWarps2 * 1024 : 0);
  P.off_b1 = P.off_bar + 1024;
  P.total = P.off_b1 + 4 * P.C * 4 + 1024 /*alignment slack*/;
  return P.total <= kSmemMax && n1 <= kMaxSlots && n2 <= kMaxSlots;
}

__host__ __device__ constexpr bool make_plan2(Plan2& P, int C, bool te, bool ht, bool slab) {
  P.C = C; P.NJ = (4 * C) / NH; P.stg = te ? 1 : 0; P.ht = ht ? 1 : 0; P.slab = slab ? 1 : 0;
  P.nstg = (te && C <= 80) ? 2 : 1;         // C = 96 would lose its resident weights to a second tile
  P.nfull = C / 64; P.t32 = (C % 64) >= 32 ? 1 : 0; P.t16 = (C % 32) >= 16 ? 1 : 0;
  P.y_bytes = rup1k(FM * C * 2);
  P.w1_bytes = rup1k(NH * C * 2);
  P.w2_bytes = rup1k(C * 128);
  if (plan2_try(P, 2, P.NJ, P.NJ, 1)) return true;      // everything resident, y double buffered
  if (plan2_try(P, 2, 3, 4, 0)) return true;
  if (plan2_try(P, 1, 3, 4, 0)) return true;
  if (plan2_try(P, 1, 2, 3, 0)) return true;
  if (plan2_try(P, 1, 2, 2, 0)) return true;
  return plan2_try(P, 1, 2, 1, 0);                      // C = 320: 80 KB y tile, 40 KB weight chunks
}

__host__ __device__ constexpr Plan2 plan2_for(int C, bool te, bool ht, bool slab) {
  Plan2 P;
  P.ok = make_plan2(P, C, te, ht, slab);
  return P;
}

// Hidden activation from the fc1 accumulator pair `acc` and the staged bias pair `b` (= kB1Scale * b1).  Both forms leave
// a power-of-two multiple of GELU in H and the D2 epilogue rescales the accumulator inside the FFMA that adds b2 --
// power-of-two scalings commute with the bf16 rounding of H, so the result is that of the GELU form itself.
//   default              : clamp-free cubic-tanh form (max error 2.7e-4 against 0.5 |x| 2^-11 = 7e-4 of tanh.approx itself),
//                          H = 2 GELU (tc_common.cuh gelu_twice2): 136 FMA / MUFU / convert instructions per 32 elements
//   -DBTSB_GELU_QUINTIC  : quintic-tanh form (max error 2.5e-5), H = GELU / 4 (gelu_quarter_quintic2): 160 instructions
// The GELU warps are issue-bound once the D2 epilogue has its own warps (ncu r02y: issue slots 71 %, top stall
// not-selected), so the instruction count is what counts: 228 vs 251 us at C = 80, 175 vs 182 us at C = 160
// (profiles/r02cu); the bf16 logit errors of the two builds differ by noise (3.1-9.2e-3 vs 3.4-12.0e-3 over the cases).
#ifdef BTSB_GELU_QUINTIC
constexpr float kB1Scale = 0.125f, kD2Scale = 4.0f;
__device__ __forceinline__ f32x2_t hidden_act2(f32x2_t acc, f32x2_t b) {
  return tc::gelu_quarter_quintic2(fma3_f32x2(acc, pack_f32x2(0.125f, 0.125f), b));
}
#else
constexpr float kB1Scale = 1.0f, kD2Scale = 0.5f;
__device__ __forceinline__ f32x2_t hidden_act2(f32x2_t acc, f32x2_t b) { return tc::gelu_twice2(add_f32x2(acc, b)); }
#endif

// K-major operand tile descriptor for a block whose rows are `sw` bytes (128 / 64 / 32) with the matching swizzle
__device__ __forceinline__ uint64_t smem_desc_k(uint32_t saddr, int sw) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * sw) >> 4) << 32;                 // stride byte offset: 8 rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw == 128 ? 2 : (sw == 64 ? 4 : 6)) << 61;
  return d;
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]^T : A (128 rows x 16 bf16) read from 8 TMEM columns (two K elements per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The C/16 column chunks of a 32-row quarter are staged as slabs of 4 / 2 / 1 chunks (128- / 64- / 32-byte rows, one
// bulk tensor copy each); chunk gi lives in the slab that starts at chunk s0[gi] and is w[gi] chunks wide.
struct SlabMap { int s0[16] = {}, w[16] = {}; };
__host__ __device__ constexpr SlabMap make_slab_map(int groups) {
  SlabMap m;
  for (int s0 = 0; s0 < groups;) {
    const int left = groups - s0;
    const int w = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
    for (int c = 0; c < w; ++c) { m.s0[s0 + c] = s0; m.w[s0 + c] = w; }
    s0 += w;
  }
  return m;
}

struct Maps2 {
  CUtensorMap y128, y64, y32;      // [M, C]   boxes [128 rows x 64|32|16 cols], swizzle 128|64|32 B
  CUtensorMap a128, a64, a32;      // W1 [4C, C]: boxes [64 rows x 64|32|16 cols]
  CUtensorMap w2;                  // W2 [C, 4C]: box [C rows x 64 cols], swizzle 128 B
  CUtensorMap r128, r64, r32;      // res [M, C]: boxes [32 rows x 64|32|16 cols] (bulk loads into the staging tile)
  CUtensorMap o128, o64, o32;      // out [M, C]: same boxes (bulk stores from the staging tile)
};
}  // namespace

// EP (wide variants only): how the residual / output rows of the D2 epilogue move.  With per-thread 16-byte accesses
// every access is its own L2 request (2.95 M write requests per launch at C = 320) and the single D2 accumulator is
// drained for ~18 k of a tile's 54 k clocks while G2 of the next tile waits (DESIGN.md lesson 9).  Measured per launch
// at M = 73 728, C = 320 (profiles/r02a): per-thread accesses 138 us, 256-bit per-thread accesses 128 us (removed),
//   EP = 2: per-warp 1 KB slabs and [32 rows x 16 columns] bulk tensor loads / stores: 125 us; the single slab per
//           warp serialises load -> update -> store per piece, so the drain still takes ~15 k clocks;
// A third variant (EP = 3: the y tile's own 80 KB as the staging tile, residual rows in by bulk tensor loads, W2 stream on
// its own producer warp) was parity-green and measured 130 us (profiles/r02b): the drain shrank to ~7 k clocks but the
// y tile of the next row tile could only be fetched afterwards (~4 k clocks for 80 KB with every SM at its tile boundary
// at once).  Removed.  In steady state this kernel streams W1 + W2 (1.6 MB per 128-row tile, 80 KB per ~1800-clock hidden
// chunk = 45 B/clk/SM, 7.1 TB/s chip-wide) -- it sits on the L2 -> SM bandwidth, not on the tensor pipe.
template <int C, bool TE, bool HT, int EP = 0, bool TRACE = false, bool XF16 = false>   // XF16: res / out rows are IEEE fp16
__global__ void __launch_bounds__(TE ? kThreadsDW : kThreads2, 1)   // 19 warps: 96 registers; 23 (dedicated D2 warps): 88
mlp_fused2_kernel(const __grid_constant__ Maps2 tm, const float* __restrict__ b1, const float* __restrict__ b2,
                  const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ res,
                  __nv_bfloat16* __restrict__ out, int M, long long* __restrict__ trace) {
  using namespace tc;
  // timeline debugging (btsb_debug_mlp_trace): block 0 records clock64() at the hand-off points of its first
  // kTraceChunks hidden chunks.  A compile-time variant (TRACE): as a run-time branch the clock reads and their predicates
  // were ~2 % of the production kernel's issued instructions (predicated-off CS2R still takes an issue slot).
  constexpr int kTraceChunks = 64, kTraceEv = 8;
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  auto tr = [&](int role, uint32_t g, int ev) {
    if constexpr (TRACE) {
      if (tracing && g < (uint32_t)kTraceChunks) trace[((size_t)role * kTraceChunks + g) * kTraceEv + ev] = clock64();
    }
  };
  constexpr bool TS = EP == 2 || EP == 4, RED = EP == 4;
  constexpr bool DW = TE;                     // dedicated D2-epilogue warps (kG2Warp + 1 .. kG2Warp + 4)
  static_assert(EP == 0 || EP == 2 || EP == 4, "EP");
  static_assert(EP == 0 || !TE, "EP variants belong to the wide (non-staging) kernels");
  constexpr Plan2 P = plan2_for(C, TE, HT, TS);
  static_assert(P.ok, "no shared-memory plan for this C");
  // EP = 4 (in-place: out == res): the update gamma * (acc + b2) leaves through the slab as a bulk tensor REDUCTION
  // (global += slab, bf16 add at the L2) -- no residual load at all, so a piece costs one slab round trip instead of a
  // load -> update -> store chain (the ~15 k-clock drain of EP = 2 holds back the next tile's G2: lesson 9 / 12)
  // wide C (256, 320): one D2 accumulator instead of two (the D2 epilogue of a tile then holds back the first G2 of the
  // next one), and for C > 256 every G2 step is two UMMAs of N = C/2 columns
  constexpr int D2B = (kD2Col + 2 * C + (HT ? 64 : 0) <= 512) ? 2 : 1;
  constexpr int NSPLIT = C > 256 ? 2 : 1;
  constexpr int NC = C / NSPLIT;
  constexpr int kHCol = kD2Col + D2B * C;                        // TMEM columns of H[0], H[1] (32 each) in HT mode
  static_assert(kD2Col + D2B * C + (HT ? 64 : 0) <= 512, "TMEM column budget");
  static_assert(NC % 16 == 0 && NC <= 256, "G2 UMMA width");
  constexpr int NJ = P.NJ;
  constexpr int nkb = P.nfull + P.t32 + P.t16;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t sbase = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sal = smem_dyn + (sbase - smem_u32(smem_dyn));
  const int m_tiles = (M + FM - 1) / FM;
  const uint32_t nt = blockIdx.x < (uint32_t)m_tiles ? (uint32_t)(m_tiles - 1 - blockIdx.x) / gridDim.x + 1u : 0u;
  const uint32_t total = nt * (uint32_t)NJ;                      // hidden chunks this CTA processes

  const uint32_t bar0 = sbase + P.off_bar;
  auto y_full = [&](int i) { return bar0 + 8u * i; };                         // 2
  auto y_empty = [&](int i) { return bar0 + 8u * (2 + i); };                  // 2
  auto d1_full = [&](int i) { return bar0 + 8u * (4 + i); };                  // 2
  auto d1_empty = [&](int i) { return bar0 + 8u * (6 + i); };                 // 2
  auto h_full = [&](int i) { return bar0 + 8u * (8 + i); };                   // 2
  auto h_empty = [&](int i) { return bar0 + 8u * (10 + i); };                 // 2
  auto d2_full = [&](int i) { return bar0 + 8u * (12 + i); };                 // 2
  auto d2_empty = [&](int i) { return bar0 + 8u * (14 + i); };                // 2
  auto w1_full = [&](int i) { return bar0 + 8u * (16 + i); };                 // kMaxSlots
  auto w1_empty = [&](int i) { return bar0 + 8u * (16 + kMaxSlots + i); };
  auto w2_full = [&](int i) { return bar0 + 8u * (16 + 2 * kMaxSlots + i); };
  auto w2_empty = [&](int i) { return bar0 + 8u * (16 + 3 * kMaxSlots + i); };
  auto res_bar = [&](int i) { return bar0 + 8u * (16 + 4 * kMaxSlots + i); };      // kEpiWarps2 (staging path only)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sal + P.off_bar + 8 * (16 + 4 * kMaxSlots + kEpiWarps2));

  // warp index made provably warp-uniform so the role branches below are convergent (the single-thread roles then
  // compile to predicated uniform-datapath instructions instead of per-lane serialisation loops)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm.y128); tma_prefetch_desc(&tm.a128); tma_prefetch_desc(&tm.w2);
    if (P.t32) { tma_prefetch_desc(&tm.y64); tma_prefetch_desc(&tm.a64); }
    if (P.t16) { tma_prefetch_desc(&tm.y32); tma_prefetch_desc(&tm.a32); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(y_full(i), 1); mbar_init(y_empty(i), 1);
      mbar_init(d1_full(i), 1); mbar_init(d1_empty(i), kEpiWarps2 / 2);
      mbar_init(h_full(i), kEpiWarps2 / 2); mbar_init(h_empty(i), 1);
      mbar_init(d2_full(i), 1); mbar_init(d2_empty(i), DW ? kD2Warps : kEpiWarps2);
    }
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(w1_full(i), 1); mbar_init(w1_empty(i), 1); mbar_init(w2_full(i), 1); mbar_init(w2_empty(i), 1);
    }
    for (int i = 0; i < kEpiWarps2; ++i) mbar_init(res_bar(i), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32((const void*)tmem_slot), 512); tmem_relinquish(); }
  float* b1s = reinterpret_cast<float*>(sal + P.off_b1);            // fc1 bias staged once per CTA
  for (int i = threadIdx.x; i < 4 * C; i += (int)blockDim.x) b1s[i] = kB1Scale * __ldg(b1 + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer (whole warp walks the schedule, one elected lane issues) ==============
    // issue order == consumption order of the MMA warp: y_0, W1_0, W1_1, then per chunk g: W1_{g+2} (+ y of its tile), W2_g
    int j1 = 0, s1 = 0, yb = 0; uint32_t ph1 = 0, yph = 0, t1 = 0;     // state of the G1-input stream
    int j2 = 0, s2 = 0; uint32_t ph2 = 0, g2 = 0;                       // state of the W2 stream
    auto load_g1_inputs = [&]() {
      if (j1 == 0) {                                                    // first chunk of a tile: its y rows
        const int tile = blockIdx.x + (int)t1 * gridDim.x;
        mbar_wait_spin(y_empty(yb), yph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(y_full(yb), (uint32_t)(FM * C * 2));
          uint32_t dst = sbase + yb * P.y_bytes;
#pragma unroll
          for (int kb = 0; kb < P.nfull; ++kb) tma_load_2d(dst + kb * FM * 128, &tm.y128, y_full(yb), kb * 64, tile * FM);
          if (P.t32) tma_load_2d(dst + P.nfull * FM * 128, &tm.y64, y_full(yb), P.nfull * 64, tile * FM);
          if (P.t16) tma_load_2d(dst + P.nfull * FM * 128 + P.t32 * FM * 64, &tm.y32, y_full(yb), P.nfull * 64 + P.t32 * 32, tile * FM);
        }
        __syncwarp();
        if (P.ny == 2) { yb ^= 1; if (yb == 0) yph ^= 1u; } else { yph ^= 1u; }
      }
      if (!P.resident || t1 == 0) {
        mbar_wait_spin(w1_empty(s1), ph1 ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(w1_full(s1), (uint32_t)(NH * C * 2));
          const uint32_t dst = sbase + P.off_w1 + s1 * P.w1_bytes;
#pragma unroll
          for (int kb = 0; kb < P.nfull; ++kb) tma_load_2d(dst + kb * NH * 128, &tm.a128, w1_full(s1), kb * 64, j1 * NH);
          if (P.t32) tma_load_2d(dst + P.nfull * NH * 128, &tm.a64, w1_full(s1), P.nfull * 64, j1 * NH);
          if (P.t16) tma_load_2d(dst + P.nfull * NH * 128 + P.t32 * NH * 64, &tm.a32, w1_full(s1), P.nfull * 64 + P.t32 * 32, j1 * NH);
        }
        __syncwarp();
      }
      if (++s1 == P.n1) { s1 = 0; ph1 ^= 1u; }
      if (++j1 == NJ) { j1 = 0; ++t1; if (P.resident) { s1 = 0; } }
    };
    auto load_w2 = [&]() {
      if (!P.resident || g2 < (uint32_t)NJ) {
        mbar_wait_spin(w2_empty(s2), ph2 ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(w2_full(s2), (uint32_t)(C * 128));
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h)
            tma_load_2d(sbase + P.off_w2 + s2 * P.w2_bytes + h * NC * 128, &tm.w2, w2_full(s2), j2 * NH, h * NC);
        }
        __syncwarp();
      }
      if (++s2 == P.n2) { s2 = 0; ph2 ^= 1u; }
      if (++j2 == NJ) { j2 = 0; if (P.resident) s2 = 0; }
      ++g2;
    };
    if (total > 0) load_g1_inputs();
    if (total > 1) load_g1_inputs();
    for (uint32_t g = 0; g < total; ++g) {
      if (g + 2 < total) load_g1_inputs();
      load_w2();
    }
  } else if (warp == 1 || warp == kG2Warp) {
    // ============================== MMA issuers (whole warp walks its schedule, one elected lane issues) ===============
    constexpr uint32_t idesc1 = idesc_bf16_f32(FM, NH);
    constexpr uint32_t idesc2 = idesc_bf16_f32(FM, NC);
    // G1 stream state
    int j1 = 0, s1 = 0, yb = 0, b1i = 0; uint32_t ph1 = 0, yph = 0, dph1 = 0, t1 = 0;
    // G2 stream state
    int j2 = 0, s2 = 0, b2i = 0, tb = 0; uint32_t ph2 = 0, hph = 0, d2ph = 0, t2 = 0;
    auto do_g1 = [&]() {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ya = sbase + yb * P.y_bytes;
        const uint32_t wa = sbase + P.off_w1 + s1 * P.w1_bytes;
        const uint32_t dcol = tmem_base + (uint32_t)(b1i * NH);
#pragma unroll
        for (int kb = 0; kb < nkb; ++kb) {
          const int sw = kb < P.nfull ? 128 : ((kb == P.nfull && P.t32) ? 64 : 32);
          const int yoff = kb < P.nfull ? kb * FM * 128 : (P.nfull * FM * 128 + ((kb == P.nfull || !P.t32) ? 0 : FM * 64));
          const int woff = kb < P.nfull ? kb * NH * 128 : (P.nfull * NH * 128 + ((kb == P.nfull || !P.t32) ? 0 : NH * 64));
          const uint64_t ad = smem_desc_k(ya + yoff, sw), bd = smem_desc_k(wa + woff, sw);
#pragma unroll
          for (int kk = 0; kk < sw / 32; ++kk)
            umma_bf16(dcol, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc1, (kb | kk) != 0 ? 1u : 0u);
        }
        umma_commit(d1_full(b1i));
        if (!P.resident) umma_commit(w1_empty(s1));
        if (j1 == NJ - 1) umma_commit(y_empty(yb));                // y tile no longer needed once these retire
      }
      __syncwarp();
      b1i ^= 1; if (b1i == 0) dph1 ^= 1u;
      if (++s1 == P.n1) { s1 = 0; ph1 ^= 1u; }
      if (++j1 == NJ) {
        j1 = 0; ++t1;
        if (P.resident) s1 = 0;
        if (P.ny == 2) { yb ^= 1; if (yb == 0) yph ^= 1u; } else { yph ^= 1u; }
      }
    };
    auto do_g2 = [&]() {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bd = smem_desc_sw128(sbase + P.off_w2 + s2 * P.w2_bytes);
        const uint32_t dcol = tmem_base + (uint32_t)(kD2Col + tb * C);
        if (HT) {
          const uint32_t acol = tmem_base + (uint32_t)(kHCol + b2i * 32);
#pragma unroll
          for (int kk = 0; kk < NH / 16; ++kk)
#pragma unroll
            for (int h = 0; h < NSPLIT; ++h)      // rows [h NC, (h+1) NC) of the W2 chunk -> D2 columns [h NC, (h+1) NC)
              umma_bf16_ts(dcol + (uint32_t)(h * NC), acol + (uint32_t)(8 * kk), bd + (uint64_t)(h * NC * 8 + 2 * kk), idesc2,
                           (j2 | kk) != 0 ? 1u : 0u);
        } else {
          const uint64_t ad = smem_desc_sw128(sbase + P.off_h + b2i * kHBytes);
#pragma unroll
          for (int kk = 0; kk < NH / 16; ++kk)
#pragma unroll
            for (int h = 0; h < NSPLIT; ++h)
              umma_bf16(dcol + (uint32_t)(h * NC), ad + (uint64_t)(2 * kk), bd + (uint64_t)(h * NC * 8 + 2 * kk), idesc2,
                        (j2 | kk) != 0 ? 1u : 0u);
        }
        umma_commit(h_empty(b2i));
        if (!P.resident) umma_commit(w2_empty(s2));
        if (j2 == NJ - 1) umma_commit(d2_full(tb));
      }
      __syncwarp();
      b2i ^= 1; if (b2i == 0) hph ^= 1u;
      if (++s2 == P.n2) { s2 = 0; ph2 ^= 1u; }
      if (++j2 == NJ) {
        j2 = 0; ++t2;
        if (P.resident) s2 = 0;
        if (++tb == D2B) { tb = 0; d2ph ^= 1u; }
      }
    };
    // Two issuing warps, one per GEMM stream, each with plain blocking waits.  A single warp alternating between the two
    // streams was the pacing element of the whole kernel (scripts/mlp_trace.py, profiles/r01i): issuing one chunk's
    // G1 or G2 holds the warp for 350-400 clk and the next mbarrier poll answers only 200-550 clk later, so one warp
    // delivered a chunk every ~1500 clk while the tensor pipe sat at 21-33 % and the epilogue warps spent half their
    // time waiting for D1.  With the streams on separate warps the two hand-off latencies overlap.
    if (warp == 1) {
      for (uint32_t n1 = 0; n1 < total; ++n1) {
        if (j1 == 0) mbar_wait_spin(y_full(yb), yph);
        if (!P.resident || t1 == 0) mbar_wait_spin(w1_full(s1), ph1);
        mbar_wait_spin(d1_empty(b1i), dph1 ^ 1u);            // the epilogue has pulled chunk g-2 out of D1[b]
        if (lane == 0) tr(0, n1, 2);
        do_g1();
        if (lane == 0) tr(0, n1, 0);
      }
    } else {
      for (uint32_t n2 = 0; n2 < total; ++n2) {
        if (!P.resident || t2 == 0) mbar_wait_spin(w2_full(s2), ph2);
        if (j2 == 0) mbar_wait_spin(d2_empty(tb), d2ph ^ 1u);
        mbar_wait_spin(h_full(b2i), hph);
        if (lane == 0) tr(0, n2, 3);
        do_g2();
        if (lane == 0) tr(0, n2, 1);
      }
    }
  } else if (DW && warp > kG2Warp) {
    // ============================== D2-epilogue warps (narrow kernels only) ==============================
    // warp = one TMEM lane quarter: 32 rows x all C columns of every tile.  Residual rows arrive by bulk tensor loads into
    // the quarter's slabs of a staging tile (two tiles for C <= 80: the rows of tile t+1 land while tile t drains), are
    // updated in place (thread = row) and leave by bulk tensor stores; the TMEM load of chunk gi+1 is in flight during
    // the arithmetic of chunk gi.
   if constexpr (DW) {
    const int q = warp & 3;
    const int dw = warp - (kG2Warp + 1);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int groups2 = C / 16;
    constexpr SlabMap SL = make_slab_map(groups2);
    constexpr int kStgTile = C * 256;                         // [128 rows x C] bf16: four quarters of groups2 KB
    static_assert(groups2 <= 16, "slab map");
    auto stg_off = [&](uint32_t tl) { return (uint32_t)(P.off_stg + (int)(tl % (uint32_t)P.nstg) * kStgTile + q * groups2 * 1024); };
    auto rbar = [&](uint32_t tl) { return res_bar(dw * 2 + (int)(tl % (uint32_t)P.nstg)); };
    auto res_issue = [&](uint32_t tl) {
      if (lane == 0) {
        const int row0 = (int)(blockIdx.x + tl * gridDim.x) * FM + q * 32;
        tma_store_wait_read();                               // earlier bulk stores have read the slabs being refilled
        mbar_expect_tx(rbar(tl), (uint32_t)(groups2 * 1024));
#pragma unroll
        for (int gi = 0; gi < groups2; ++gi) {
          if (SL.s0[gi] != gi) continue;                     // first chunk of a slab issues the slab's copy
          const CUtensorMap* mp = SL.w[gi] == 4 ? &tm.r128 : (SL.w[gi] == 2 ? &tm.r64 : &tm.r32);
          tma_load_2d(sbase + stg_off(tl) + (uint32_t)(gi * 1024), mp, rbar(tl), gi * 16, row0);
        }
      }
      __syncwarp();
    };
    if (nt > 0) res_issue(0);
    if (P.nstg == 2 && nt > 1) res_issue(1);
    for (uint32_t tl = 0; tl < nt; ++tl) {
      const int tb = (int)(tl % (uint32_t)D2B);
      const int row0 = (int)(blockIdx.x + tl * gridDim.x) * FM + q * 32;
      unsigned char* stg = sal + stg_off(tl);
      mbar_wait_spin(d2_full(tb), (tl / (uint32_t)D2B) & 1u);
      tc_fence_after();
      uint32_t ra[16], rb[16];
      tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C), ra);
      mbar_wait_spin(rbar(tl), (tl / (uint32_t)P.nstg) & 1u);
#pragma unroll
      for (int gi = 0; gi < groups2; ++gi) {
        uint32_t (&r)[16] = (gi & 1) ? rb : ra;
        tmem_ld_wait();
        if (gi + 1 < groups2) {
          tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C + (gi + 1) * 16), (gi & 1) ? ra : rb);
        } else {                                             // last D2 read of this warp for this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty(tb));
        }
        const int sw = SL.w[gi] * 32, c = gi - SL.s0[gi];
        const int xr = sw == 128 ? (lane & 7) : (sw == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
        unsigned char* rowp = stg + SL.s0[gi] * 1024 + lane * sw;
        uint4* p0 = reinterpret_cast<uint4*>(rowp + (((2 * c) ^ xr) << 4));
        uint4* p1 = reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ xr) << 4));
        const uint4 r0 = *p0, r1 = *p1;
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        const int n = gi * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n + i));
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n + i));
          v[i] = fmaf(g4.x, fmaf(__uint_as_float(r[i]), kD2Scale, b4.x), x2_lo<XF16>(rr[i / 2]));
          v[i + 1] = fmaf(g4.y, fmaf(__uint_as_float(r[i + 1]), kD2Scale, b4.y), x2_hi<XF16>(rr[i / 2]));
          v[i + 2] = fmaf(g4.z, fmaf(__uint_as_float(r[i + 2]), kD2Scale, b4.z), x2_lo<XF16>(rr[i / 2 + 1]));
          v[i + 3] = fmaf(g4.w, fmaf(__uint_as_float(r[i + 3]), kD2Scale, b4.w), x2_hi<XF16>(rr[i / 2 + 1]));
        }
        *p0 = make_uint4(pack_x2<XF16>(v[0], v[1]), pack_x2<XF16>(v[2], v[3]), pack_x2<XF16>(v[4], v[5]), pack_x2<XF16>(v[6], v[7]));
        *p1 = make_uint4(pack_x2<XF16>(v[8], v[9]), pack_x2<XF16>(v[10], v[11]), pack_x2<XF16>(v[12], v[13]), pack_x2<XF16>(v[14], v[15]));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int gi = 0; gi < groups2; ++gi) {
          if (SL.s0[gi] != gi) continue;
          const CUtensorMap* mp = SL.w[gi] == 4 ? &tm.o128 : (SL.w[gi] == 2 ? &tm.o64 : &tm.o32);
          tma_store_2d(mp, sbase + stg_off(tl) + (uint32_t)(gi * 1024), gi * 16, row0);   // rows beyond M are clipped
        }
        tma_store_commit();
      }
      __syncwarp();
      if (tl + (uint32_t)P.nstg < nt) res_issue(tl + (uint32_t)P.nstg);   // refills the slabs just stored from
    }
    if (lane == 0) tma_store_wait_all();                     // shared memory must outlive the last bulk store's reads
   }
  } else {
    // ============================== epilogue warps (2 .. 17) ==============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int k4 = (warp - 2) >> 2;         // 0..3
    const int grp = k4 & 1;                 // owns hidden chunks g with (g & 1) == grp
    const int half = k4 >> 1;               // 32-column half of the 64-column chunk
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int groups2 = C / 16;
    constexpr int NU = (groups2 + 3) / 4;   // 16-column groups of D2 per warp
    constexpr bool PRE = NU <= 3;           // wide C: 5 groups per warp would cost 40 registers -> load at use instead
    uint4 rpre[PRE ? NU : 1][2];             // residual rows of the tile whose D2 epilogue is pending

    auto prefetch_res = [&](int tile) {
      if (!PRE) return;
      const int row = tile * FM + r_in_tile;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int gi = k4 + 4 * u;
        if (gi < groups2 && row < M) {
          const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)row * C + gi * 16);
          rpre[PRE ? u : 0][0] = __ldg(rp); rpre[PRE ? u : 0][1] = __ldg(rp + 1);
        } else {
          rpre[PRE ? u : 0][0] = make_uint4(0, 0, 0, 0); rpre[PRE ? u : 0][1] = make_uint4(0, 0, 0, 0);
        }
      }
    };
    // D2 -> +b2 -> *gamma + res -> bf16 rows of tile `tile` (local index tl); column groups k4, k4+4, k4+8
    auto d2_epilogue = [&](int tile, uint32_t tl) {
      const int tb = (int)(tl % (uint32_t)D2B);
      mbar_wait_spin(d2_full(tb), (tl / (uint32_t)D2B) & 1u);
      tc_fence_after();
      const int row = tile * FM + r_in_tile;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int gi = k4 + 4 * u;
        if (gi >= groups2) break;
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C + gi * 16), r);
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        if (PRE) { r0 = rpre[PRE ? u : 0][0]; r1 = rpre[PRE ? u : 0][1]; }
        else if (row < M) {
          const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)row * C + gi * 16);
          r0 = __ldg(rp); r1 = __ldg(rp + 1);
        }
        tmem_ld_wait();
        if (gi + 4 >= groups2) {                         // last D2 read of this warp for this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty(tb));
        }
        if (row < M) {
          const int n = gi * 16;
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n + i));
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n + i));
            v[i] = fmaf(g4.x, fmaf(__uint_as_float(r[i]), kD2Scale, b4.x), x2_lo<XF16>(rr[i / 2]));
            v[i + 1] = fmaf(g4.y, fmaf(__uint_as_float(r[i + 1]), kD2Scale, b4.y), x2_hi<XF16>(rr[i / 2]));
            v[i + 2] = fmaf(g4.z, fmaf(__uint_as_float(r[i + 2]), kD2Scale, b4.z), x2_lo<XF16>(rr[i / 2 + 1]));
            v[i + 3] = fmaf(g4.w, fmaf(__uint_as_float(r[i + 3]), kD2Scale, b4.w), x2_hi<XF16>(rr[i / 2 + 1]));
          }
          uint4 o0, o1;
          o0.x = pack_x2<XF16>(v[0], v[1]); o0.y = pack_x2<XF16>(v[2], v[3]);
          o0.z = pack_x2<XF16>(v[4], v[5]); o0.w = pack_x2<XF16>(v[6], v[7]);
          o1.x = pack_x2<XF16>(v[8], v[9]); o1.y = pack_x2<XF16>(v[10], v[11]);
          o1.z = pack_x2<XF16>(v[12], v[13]); o1.w = pack_x2<XF16>(v[14], v[15]);
          uint4* op = reinterpret_cast<uint4*>(out + (size_t)row * C + n);
          op[0] = o0; op[1] = o1;
        }
      }
    };

    // ---- staging path (TE): the warp owns the contiguous D2 column range [c_lo, c_hi) chunks of its 32 rows.  Residual
    // rows arrive by bulk tensor loads into the warp's slab(s) of the staging tile (issued one hidden chunk ahead), are
    // updated IN PLACE (thread = row: it reads and rewrites only its own 32 bytes per chunk) and leave by bulk tensor
    // stores -- no per-thread scattered global access, and nothing for the per-chunk proxy fence (MEMBAR) to wait on.
    const int ew = warp - 2;
    const int c_lo = (groups2 * k4) / 4, c_hi = (groups2 * (k4 + 1)) / 4;
    const uint32_t stg_addr = sbase + P.off_stg + (uint32_t)((q * groups2 + c_lo) * 1024);
    unsigned char* stg = sal + P.off_stg + (q * groups2 + c_lo) * 1024;
    uint32_t rph = 0;                        // phase of this warp's residual barrier
    auto res_issue = [&](int tile) {
      if (lane == 0 && c_lo < c_hi) {
        tma_store_wait_read();                             // the previous tile's bulk stores have drained the slabs
        mbar_expect_tx(res_bar(ew), (uint32_t)((c_hi - c_lo) * 1024));
        uint32_t off = 0;
        for (int s0 = c_lo; s0 < c_hi;) {
          const int left = c_hi - s0;
          const int w = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
          const CUtensorMap* mp = w == 4 ? &tm.r128 : (w == 2 ? &tm.r64 : &tm.r32);
          tma_load_2d(stg_addr + off, mp, res_bar(ew), s0 * 16, tile * FM + q * 32);
          off += (uint32_t)(w * 1024); s0 += w;
        }
      }
      __syncwarp();
    };
    auto d2_epilogue_te = [&](int tile, uint32_t tl) {
      const int tb = (int)(tl % (uint32_t)D2B);
      mbar_wait_spin(d2_full(tb), (tl / (uint32_t)D2B) & 1u);
      tc_fence_after();
      if (c_lo >= c_hi) {                                  // (cannot happen for C >= 64: every warp owns >= 1 chunk)
        tc_fence_before(); __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty(tb));
        return;
      }
      mbar_wait_spin(res_bar(ew), rph); rph ^= 1u;
      uint32_t off = 0;
      for (int s0 = c_lo; s0 < c_hi;) {
        const int left = c_hi - s0;
        const int w = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
        const int sw = w * 32;
        const int xr = sw == 128 ? (lane & 7) : (sw == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
        unsigned char* rowp = stg + off + lane * sw;
        for (int c = 0; c < w; ++c) {
          const int gi = s0 + c;
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C + gi * 16), r);
          tmem_ld_wait();
          if (gi + 1 == c_hi) {                            // last D2 read of this warp for this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty(tb));
          }
          uint4* p0 = reinterpret_cast<uint4*>(rowp + (((2 * c) ^ xr) << 4));
          uint4* p1 = reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ xr) << 4));
          const uint4 r0 = *p0, r1 = *p1;
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          const int n = gi * 16;
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n + i));
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n + i));
            v[i] = fmaf(g4.x, fmaf(__uint_as_float(r[i]), kD2Scale, b4.x), x2_lo<XF16>(rr[i / 2]));
            v[i + 1] = fmaf(g4.y, fmaf(__uint_as_float(r[i + 1]), kD2Scale, b4.y), x2_hi<XF16>(rr[i / 2]));
            v[i + 2] = fmaf(g4.z, fmaf(__uint_as_float(r[i + 2]), kD2Scale, b4.z), x2_lo<XF16>(rr[i / 2 + 1]));
            v[i + 3] = fmaf(g4.w, fmaf(__uint_as_float(r[i + 3]), kD2Scale, b4.w), x2_hi<XF16>(rr[i / 2 + 1]));
          }
          *p0 = make_uint4(pack_x2<XF16>(v[0], v[1]), pack_x2<XF16>(v[2], v[3]), pack_x2<XF16>(v[4], v[5]), pack_x2<XF16>(v[6], v[7]));
          *p1 = make_uint4(pack_x2<XF16>(v[8], v[9]), pack_x2<XF16>(v[10], v[11]), pack_x2<XF16>(v[12], v[13]), pack_x2<XF16>(v[14], v[15]));
        }
        off += (uint32_t)(w * 1024); s0 += w;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        uint32_t o2 = 0;
        for (int s0 = c_lo; s0 < c_hi;) {
          const int left = c_hi - s0;
          const int w = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
          const CUtensorMap* mp = w == 4 ? &tm.o128 : (w == 2 ? &tm.o64 : &tm.o32);
          tma_store_2d(mp, stg_addr + o2, s0 * 16, tile * FM + q * 32);
          o2 += (uint32_t)(w * 1024); s0 += w;
        }
        tma_store_commit();
      }
    };


    // ---- slab path (TS, wide C): the warp's column groups k4, k4+4, ... go one at a time through its private 1 KB slab:
    // bulk tensor load of the residual piece [32 rows x 16 columns] -> in-place update (thread = row) -> bulk tensor
    // store.  The load of the next piece is issued as soon as the store of the current one has read the slab.
    auto d2_epilogue_ts = [&](int tile, uint32_t tl) {
      const int tb = (int)(tl % (uint32_t)D2B);
      const uint32_t slab_addr = sbase + P.off_slab + (uint32_t)(ew * 1024);
      unsigned char* slab = sal + P.off_slab + ew * 1024;
      const int row0 = tile * FM + q * 32;
      auto slab_load = [&](int gi) {
        if (RED) return;                                     // in-place reduction: nothing is loaded
        if (lane == 0) {
          tma_store_wait_read();                             // the previous piece's bulk store has drained the slab
          mbar_expect_tx(res_bar(ew), 1024u);
          tma_load_2d(slab_addr, &tm.r32, res_bar(ew), gi * 16, row0);
        }
        __syncwarp();
      };
      slab_load(k4);                                         // in flight while G2's last commit is awaited
      mbar_wait_spin(d2_full(tb), (tl / (uint32_t)D2B) & 1u);
      tc_fence_after();
      const int xr = (lane >> 2) & 1;                        // 32-byte swizzle: 16-byte chunk index ^ address bit 7
      uint4* p0 = reinterpret_cast<uint4*>(slab + lane * 32 + ((0 ^ xr) << 4));
      uint4* p1 = reinterpret_cast<uint4*>(slab + lane * 32 + ((1 ^ xr) << 4));
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int gi = k4 + 4 * u;
        if (gi >= groups2) break;
        uint32_t r[16];
        tmem_ld16(lane_addr + (uint32_t)(kD2Col + tb * C + gi * 16), r);
        tmem_ld_wait();
        if (gi + 4 >= groups2) {                             // last D2 read of this warp for this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty(tb));
        }
        uint32_t rr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (RED) {
          if (lane == 0) tma_store_wait_read();              // the previous piece's reduction has read the slab
          __syncwarp();
        } else {
          mbar_wait_spin(res_bar(ew), rph); rph ^= 1u;
          const uint4 r0 = *p0, r1 = *p1;
          rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
        }
        const int n = gi * 16;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + n + i));
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + n + i));
          v[i] = fmaf(g4.x, fmaf(__uint_as_float(r[i]), kD2Scale, b4.x), x2_lo<XF16>(rr[i / 2]));
          v[i + 1] = fmaf(g4.y, fmaf(__uint_as_float(r[i + 1]), kD2Scale, b4.y), x2_hi<XF16>(rr[i / 2]));
          v[i + 2] = fmaf(g4.z, fmaf(__uint_as_float(r[i + 2]), kD2Scale, b4.z), x2_lo<XF16>(rr[i / 2 + 1]));
          v[i + 3] = fmaf(g4.w, fmaf(__uint_as_float(r[i + 3]), kD2Scale, b4.w), x2_hi<XF16>(rr[i / 2 + 1]));
        }
        *p0 = make_uint4(pack_x2<XF16>(v[0], v[1]), pack_x2<XF16>(v[2], v[3]), pack_x2<XF16>(v[4], v[5]), pack_x2<XF16>(v[6], v[7]));
        *p1 = make_uint4(pack_x2<XF16>(v[8], v[9]), pack_x2<XF16>(v[10], v[11]), pack_x2<XF16>(v[12], v[13]), pack_x2<XF16>(v[14], v[15]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (RED) tma_reduce_add_2d(&tm.o32, slab_addr, gi * 16, row0);   // rows beyond M are clipped by the tensor map
          else tma_store_2d(&tm.o32, slab_addr, gi * 16, row0);
          tma_store_commit();
        }
        if (gi + 4 < groups2) slab_load(gi + 4);
      }
    };

    int pend_tl = -1;                       // local tile index whose D2 epilogue this warp still owes
    uint32_t prev_tl = 0xffffffffu;
    for (uint32_t g = (uint32_t)grp; g < total; g += 2) {
      const uint32_t tl = g / (uint32_t)NJ; const int j = (int)(g - tl * NJ);
      const uint32_t use = g >> 1;                       // per-buffer use count of D1[grp] / H[grp]
      if (!DW && tl != prev_tl) {
        // first chunk of this warp in a new tile: the previous tile's D2 epilogue is deferred until after this chunk's
        // GELU so its tail latency (last H hand-off -> G2 -> commit) is hidden; fetch its residual rows now
        if (prev_tl != 0xffffffffu) {
          pend_tl = (int)prev_tl;
          if (TE) res_issue(blockIdx.x + pend_tl * gridDim.x); else prefetch_res(blockIdx.x + pend_tl * gridDim.x);
        }
        prev_tl = tl;
      }
      if (lane == 0) tr(1 + ew, g, 0);
      mbar_wait_spin(d1_full(grp), use & 1u);
      tc_fence_after();
      if (lane == 0) tr(1 + ew, g, 1);
      uint32_t r[32];
      tmem_ld32(lane_addr + (uint32_t)(grp * NH + half * 32), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(d1_empty(grp)); tr(1 + ew, g, 2); }   // D1[grp] may be overwritten by G1 of chunk g+2
      const int hcol = j * NH + half * 32;
      uint32_t o[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(b1s + hcol + i);
        const float2 g0 = unpack_f32x2(hidden_act2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])),
                                                   pack_f32x2(b4.x, b4.y)));
        const float2 g1 = unpack_f32x2(hidden_act2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])),
                                                   pack_f32x2(b4.z, b4.w)));
        o[i / 2] = pack_bf16x2(g0.x, g0.y);
        o[i / 2 + 1] = pack_bf16x2(g1.x, g1.y);
      }
      if (lane == 0) tr(1 + ew, g, 3);
      mbar_wait_spin(h_empty(grp), (use & 1u) ^ 1u);         // G2 of chunk g-2 has finished reading H[grp]
      if (lane == 0) tr(1 + ew, g, 4);
      if (HT) {
        // H never touches shared memory: packed bf16 pairs go straight into the TMEM columns G2 reads its A operand
        // from (thread = row / TMEM lane, 16 columns = this warp's 32 hidden values)
        tmem_st16(lane_addr + (uint32_t)(kHCol + grp * 32 + half * 16), o);
        tmem_st_wait();
        tc_fence_before();
      } else {
        unsigned char* hb = sal + P.off_h + grp * kHBytes;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          *reinterpret_cast<uint4*>(hb + sw128_offset(r_in_tile, half * 32 + c8 * 8)) =
              make_uint4(o[4 * c8], o[4 * c8 + 1], o[4 * c8 + 2], o[4 * c8 + 3]);
        fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(h_full(grp)); tr(1 + ew, g, 5); }
      if (pend_tl >= 0) {
        if (TE) d2_epilogue_te(blockIdx.x + pend_tl * gridDim.x, (uint32_t)pend_tl);
        else if (TS) d2_epilogue_ts(blockIdx.x + pend_tl * gridDim.x, (uint32_t)pend_tl);
        else d2_epilogue(blockIdx.x + pend_tl * gridDim.x, (uint32_t)pend_tl);
        pend_tl = -1;
        if (lane == 0) tr(1 + ew, g, 6);
      }
    }
    if (!DW && nt > 0) {
      // every group has at least one chunk in every tile (NJ >= 4), so prev_tl is the CTA's last tile here
      const int last = (int)nt - 1;
      if (TE) {
        res_issue(blockIdx.x + last * gridDim.x);
        d2_epilogue_te(blockIdx.x + last * gridDim.x, (uint32_t)last);
        if (lane == 0) tma_store_wait_all();              // smem must outlive the last bulk store's reads
      } else if (TS) {
        d2_epilogue_ts(blockIdx.x + last * gridDim.x, (uint32_t)last);
        if (lane == 0) tma_store_wait_all();              // smem must outlive the last bulk store's reads
      } else {
        prefetch_res(blockIdx.x + last * gridDim.x);
        d2_epilogue(blockIdx.x + last * gridDim.x, (uint32_t)last);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// A CTA-pair variant of the wide kernel (cta_group::2: one M = 256 UMMA per step issued by the leader, each CTA staging
// half of every W1 / W2 chunk, remote mbarrier arrivals from both CTAs' GELU warps, multicast commits) was written and
// is parity-green on every kernel and model test (git history: "CTA-pair (cta_group::2) wide fused MLP"), but measured
// 128 us against 129 us per launch at C = 320 (profiles/r02e): halving the L2 -> SM weight stream changes nothing, so the
// steady state is NOT bound by that stream but by shared-memory operand reads (G1 re-reads the 80 KB y tile for every
// 64-column hidden chunk: 48 clk per N = 64 UMMA instead of 32) plus the tile-boundary drain.  Removed again.


static long long* g_mlp_trace = nullptr;    // debugging only (btsb_debug_mlp_trace); caller-owned device buffer

int num_sms();
int make_tmap_bf16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, int swizzle_bytes);

int mlp_fused2_supported(int C) {
  return C % 16 == 0 && ((C >= 64 && C <= 160) || C == 256 || C == 320);
}

template <int C, bool TE, bool HT, int EP = 0, bool TRACE = false, bool XF16 = false>
static int launch2(const Maps2& tm, const float* b1, const float* b2, const float* gamma, const void* res, void* out,
                   int64_t M, cudaStream_t st) {
  constexpr Plan2 P = plan2_for(C, TE, HT, EP == 2 || EP == 4);   // must be the kernel's own plan (slabs for EP 2 and 4)
  if constexpr (!TRACE && XF16 && (C == 80 || C == 160 || C == 320)) {   // traced variants: the bench's three widths, fp16 stream
    if (g_mlp_trace != nullptr) return launch2<C, TE, HT, EP, true, XF16>(tm, b1, b2, gamma, res, out, M, st);
  }
  auto kern = mlp_fused2_kernel<C, TE, HT, EP, TRACE, XF16>;
  BTSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax), "mlp_fused2 attr");
  const int m_tiles = (int)((M + FM - 1) / FM);
  const int grid = min(m_tiles, num_sms());
  kern<<<grid, TE ? kThreadsDW : kThreads2, P.total, st>>>(tm, b1, b2, gamma, (const __nv_bfloat16*)res, (__nv_bfloat16*)out, (int)M, g_mlp_trace);
  return launch_done("mlp_fused2");
}


int make_tmap_f16_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                        uint32_t box_cols, int swizzle_bytes);

template <bool XF16>
static int mlp_fused2_dispatch(const Maps2& tm, const float* b1, const float* b2, const float* gamma, const void* res,
                               void* out, int64_t M, int C, cudaStream_t st) {
  // wide C: no room for a separate staging tile next to the 80 KB y tile -> per-warp slabs (EP = 2); called in place
  // (out == res) the update is reduced into the residual stream at the L2 instead (EP = 4)
  if (res == out) {
    if (C == 256) return launch2<256, false, true, 4, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    if (C == 320) return launch2<320, false, true, 4, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
  }
  if (C == 256) return launch2<256, false, true, 2, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
  if (C == 320) return launch2<320, false, true, 2, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
  switch (C) {
    case 64: return launch2<64, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 80: return launch2<80, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 96: return launch2<96, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 112: return launch2<112, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 128: return launch2<128, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 144: return launch2<144, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
    case 160: return launch2<160, true, true, 0, false, XF16>(tm, b1, b2, gamma, res, out, M, st);
  }
  set_error("mlp_fused: C=%d unsupported", C);
  return BTSB_EINVAL;
}

int mlp_fused2_launch(const void* y, const void* res, const void* W1, const float* b1, const void* W2, const float* b2,
                      const float* gamma, void* out, int64_t M, int C, bool xf16, cudaStream_t st) {
  BTSB_REQUIRE(mlp_fused2_supported(C), "mlp_fused: C=%d unsupported", C);
  const int t32 = (C % 64) >= 32, t16 = (C % 32) >= 16;
  Maps2 tm;
  memset(&tm, 0, sizeof(tm));
  if (int e = make_tmap_bf16_2d_sw(&tm.y128, y, (uint64_t)M, (uint64_t)C, FM, 64, 128)) return e;
  if (int e = make_tmap_bf16_2d_sw(&tm.a128, W1, (uint64_t)(4 * C), (uint64_t)C, NH, 64, 128)) return e;
  if (t32) {
    if (int e = make_tmap_bf16_2d_sw(&tm.y64, y, (uint64_t)M, (uint64_t)C, FM, 32, 64)) return e;
    if (int e = make_tmap_bf16_2d_sw(&tm.a64, W1, (uint64_t)(4 * C), (uint64_t)C, NH, 32, 64)) return e;
  }
  if (t16) {
    if (int e = make_tmap_bf16_2d_sw(&tm.y32, y, (uint64_t)M, (uint64_t)C, FM, 16, 32)) return e;
    if (int e = make_tmap_bf16_2d_sw(&tm.a32, W1, (uint64_t)(4 * C), (uint64_t)C, NH, 16, 32)) return e;
  }
  const int NC = C > 256 ? C / 2 : C;                      // W2 chunk rows per bulk copy / per G2 UMMA
  if (int e = make_tmap_bf16_2d_sw(&tm.w2, W2, (uint64_t)C, (uint64_t)(4 * C), (uint32_t)NC, 64, 128)) return e;
  // residual / output rows: bf16 or (xf16) IEEE fp16 -- the element type matters to the in-place reduction only
  auto rmap = xf16 ? make_tmap_f16_2d_sw : make_tmap_bf16_2d_sw;
  if (int e = rmap(&tm.r128, res, (uint64_t)M, (uint64_t)C, 32, 64, 128)) return e;
  if (int e = rmap(&tm.r64, res, (uint64_t)M, (uint64_t)C, 32, 32, 64)) return e;
  if (int e = rmap(&tm.r32, res, (uint64_t)M, (uint64_t)C, 32, 16, 32)) return e;
  if (int e = rmap(&tm.o128, out, (uint64_t)M, (uint64_t)C, 32, 64, 128)) return e;
  if (int e = rmap(&tm.o64, out, (uint64_t)M, (uint64_t)C, 32, 32, 64)) return e;
  if (int e = rmap(&tm.o32, out, (uint64_t)M, (uint64_t)C, 32, 16, 32)) return e;
  if (xf16) return mlp_fused2_dispatch<true>(tm, b1, b2, gamma, res, out, M, C, st);
  return mlp_fused2_dispatch<false>(tm, b1, b2, gamma, res, out, M, C, st);
}

}  // namespace btsb

// Debugging aid (scripts/mlp_trace.py): when `buf` is non-NULL, block 0 of every following fused-MLP launch records
// clock64() timestamps into buf[(role * 64 + chunk) * 8 + event] (17 roles: MMA warp + 16 epilogue warps; >= 69632 B).
// Pass NULL to switch tracing off.  Not part of the reference-facing surface.
extern "C" int btsb_debug_mlp_trace(void* buf) {
  btsb::g_mlp_trace = (long long*)buf;
  return BTSB_OK;
}

// K4 fused entry point: replaces timm blocks.j.mlp.fc1 -> GELU -> mlp.fc2 -> *gamma -> +shortcut
// (called at /root/reference/btsbot/architectures.py:132; oracle/convnext_oracle.py block()).
extern "C" int btsb_convnext_mlp_fused_fwd(const void* y, const void* res, const void* W1, const float* b1,
                                           const void* W2, const float* b2, const float* gamma, void* out, int64_t M,
                                           int C, int dtype, void* stream) {
  using namespace btsb;
  if (int e = check_device()) return e;
  BTSB_REQUIRE(M >= 0 && M < (1ll << 31), "mlp_fused: bad M");
  BTSB_REQUIRE(dtype == BTSB_BF16 || dtype == BTSB_BF16_XF16, "mlp_fused: dtype must be BF16 or BF16_XF16");
  BTSB_REQUIRE(mlp_fused2_supported(C), "mlp_fused: C=%d unsupported (multiple of 16 in [64,160], 256 or 320)", C);
  if (M == 0) return BTSB_OK;
  BTSB_REQUIRE(y && res && W1 && b1 && W2 && b2 && gamma && out, "mlp_fused: null pointer");
  BTSB_REQUIRE(((uintptr_t)res % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)b1 % 16) == 0 &&
                   ((uintptr_t)b2 % 16) == 0 && ((uintptr_t)gamma % 16) == 0,
               "mlp_fused: pointers must be 16-byte aligned");
  return mlp_fused2_launch(y, res, W1, b1, W2, b2, gamma, out, M, C, dtype == BTSB_BF16_XF16, (cudaStream_t)stream);
}
