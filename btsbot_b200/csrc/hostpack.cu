// Host-side marshalling for the end-to-end scoring call (parallel.AlertScorer): float32 -> bf16 (round to nearest even)
// on a pool of host threads, so that a bulk-scoring step moves 195 MB instead of 391 MB over PCIe.
//
// The end-to-end rate of the bf16 mode is the PCIe rate of the reference's fp32 input format (8192 alerts = 391 MB at
// ~55 GB/s = 7.1 ms against 3.9 ms of kernels).  The first thing the bf16 trunk does with a pixel is round it to bf16 (the
// stem's MMA operand), and K1's cast + transpose does no arithmetic, so rounding on the host gives bit-identical logits.
// Not used by the fp32 mode or by the crop / normalise paths of K1.  Host memory only; no CUDA call in here.
#include <immintrin.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace btsb {
namespace {

inline uint16_t rne_bf16(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);   // NaN stays a (quiet) NaN
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

void pack_scalar(const uint32_t* s, uint16_t* d, size_t n) {
  for (size_t i = 0; i < n; ++i) d[i] = rne_bf16(s[i]);
}

__attribute__((target("avx2"))) inline __m256i cvt8_avx2(__m256i u) {
  const __m256i k7fff = _mm256_set1_epi32(0x7fff), one = _mm256_set1_epi32(1);
  const __m256i absmask = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000);
  const __m256i quiet = _mm256_set1_epi32(0x0040);
  const __m256i hi = _mm256_srli_epi32(u, 16);
  const __m256i nan = _mm256_cmpgt_epi32(_mm256_and_si256(u, absmask), inf);
  const __m256i r = _mm256_srli_epi32(_mm256_add_epi32(u, _mm256_add_epi32(k7fff, _mm256_and_si256(hi, one))), 16);
  return _mm256_blendv_epi8(r, _mm256_or_si256(hi, quiet), nan);
}

__attribute__((target("avx2"))) void pack_avx2(const uint32_t* s, uint16_t* d, size_t n) {
  size_t i = 0;
  // the output is written once and next read by the DMA engine: non-temporal stores skip the read-for-ownership of every
  // destination line (a quarter of the job's DRAM traffic) when the slice is 32-byte aligned (pinned buffers are)
  const bool nt = ((uintptr_t)d & 31u) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = cvt8_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i)));
    const __m256i b = cvt8_avx2(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 8)));
    // packus works per 128-bit lane: (a.lo, b.lo | a.hi, b.hi) -> permute the 64-bit quarters back into order
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
    if (nt) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), p);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + i), p);
  }
  if (nt) _mm_sfence();
  pack_scalar(s + i, d + i, n - i);
}

// a small persistent pool: the job is cut into contiguous slices, one per worker; the caller runs slice 0
class Pool {
 public:
  // leaked on purpose: the workers are detached and wait on the pool's condition variable for the life of the process,
  // so the pool must never be destroyed (a static destructor at exit would tear the condition variable down under them)
  static Pool& get() { static Pool* p = new Pool; return *p; }
  void run(const uint32_t* s, uint16_t* d, size_t n, int threads) {
    std::lock_guard<std::mutex> job(job_mu_);                 // one job at a time
    ensure(threads - 1);
    const bool avx2 = __builtin_cpu_supports("avx2");
    const size_t per = ((n + threads - 1) / threads + 63) & ~(size_t)63;
    {
      std::lock_guard<std::mutex> lk(mu_);
      src_ = s; dst_ = d; n_ = n; per_ = per; avx2_ = avx2; active_ = threads - 1; pending_ = threads - 1; ++gen_;
    }
    cv_.notify_all();
    slice(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  void slice(int t) {
    const size_t lo = (size_t)t * per_, hi = lo + per_ < n_ ? lo + per_ : n_;
    if (lo >= hi) return;
    if (avx2_) pack_avx2(src_ + lo, dst_ + lo, hi - lo); else pack_scalar(src_ + lo, dst_ + lo, hi - lo);
  }
  void ensure(int workers) {
    while ((int)th_.size() < workers) {
      const int id = (int)th_.size() + 1;
      th_.emplace_back([this, id] {
        uint64_t seen = 0;
        for (;;) {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return gen_ != seen; });
          seen = gen_;
          const bool mine = id <= active_;
          lk.unlock();
          if (!mine) continue;
          slice(id);
          lk.lock();
          if (--pending_ == 0) done_.notify_one();
        }
      });
      th_.back().detach();
    }
  }
  std::mutex job_mu_, mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
  const uint32_t* src_ = nullptr; uint16_t* dst_ = nullptr;
  size_t n_ = 0, per_ = 0;
  bool avx2_ = false;
  int active_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

}  // namespace
}  // namespace btsb

using namespace btsb;

// src: n float32 values, dst: n bf16 values (uint16 storage), both HOST memory (dst normally pinned).  Round to nearest
// even, NaN -> quiet NaN, +-inf kept.  threads <= 0: one per hardware thread the process may run on, at most 32.
extern "C" int btsb_host_pack_bf16(const float* src, uint16_t* dst, int64_t n, int threads) {
  BTSB_REQUIRE(n >= 0, "host_pack: n < 0");
  if (n == 0) return BTSB_OK;
  BTSB_REQUIRE(src && dst, "host_pack: null pointer");
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 32) nt = 32;
  if ((int64_t)nt * 4096 > n) nt = (int)(n / 4096) + 1;       // small jobs: not worth waking the pool
  if (nt == 1) {
    if (__builtin_cpu_supports("avx2")) pack_avx2(reinterpret_cast<const uint32_t*>(src), dst, (size_t)n);
    else pack_scalar(reinterpret_cast<const uint32_t*>(src), dst, (size_t)n);
    return BTSB_OK;
  }
  Pool::get().run(reinterpret_cast<const uint32_t*>(src), dst, (size_t)n, nt);
  return BTSB_OK;
}
