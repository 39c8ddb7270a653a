// Alert ingest, host side: the step BEFORE K1 in production.  A ZTF alert carries three gzipped
// FITS stamps (cutoutScience / cutoutTemplate / cutoutDifference .stampData); the reference opens each with
// gzip.open + astropy.io.fits.open inside a Python loop (alert_utils.py:137-147).  Here the whole batch is inflated and
// parsed by a pool of host threads straight into the dense float32 staging layout the pad_norm kernel reads
// ([n_stamps, 63*63] + [n_stamps, 2] (rows, cols)); no Python per stamp, no astropy.
//
// FITS as ZTF writes it: one primary HDU, 2880-byte header blocks of 80-character cards (SIMPLE, BITPIX, NAXIS, NAXIS1,
// NAXIS2 [, BSCALE, BZERO], END), then big-endian pixels row-major.  BITPIX -32 / -64 / 16 / 32 / 8 are handled like
// astropy does (physical = BZERO + BSCALE * stored); anything else is an error, never a silent guess.
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace btsb {
namespace {
constexpr int kMaxSide = 63;

struct Card { std::string key, val; };

bool inflate_gzip(const unsigned char* src, size_t n, std::vector<unsigned char>& out, std::string& err) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) { err = "inflateInit2 failed"; return false; }
  zs.next_in = const_cast<unsigned char*>(src);
  zs.avail_in = (uInt)n;
  out.resize(2880 * 8);                              // header block + 63*63 float32 rounded up fits; grown on demand
  size_t have = 0;
  int rc = Z_OK;
  while (rc != Z_STREAM_END) {
    if (have == out.size()) out.resize(out.size() * 2);
    zs.next_out = out.data() + have;
    zs.avail_out = (uInt)(out.size() - have);
    rc = inflate(&zs, Z_NO_FLUSH);
    have = out.size() - zs.avail_out;
    if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); err = "not a gzip stream (zlib error " + std::to_string(rc) + ")"; return false; }
    if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) { inflateEnd(&zs); err = "truncated gzip stream"; return false; }
  }
  inflateEnd(&zs);
  out.resize(have);
  return true;
}

template <typename T>
T load_be(const unsigned char* p) {
  unsigned char b[sizeof(T)];
  for (size_t i = 0; i < sizeof(T); ++i) b[i] = p[sizeof(T) - 1 - i];
  T v;
  memcpy(&v, b, sizeof(T));
  return v;
}

// one stamp: inflated FITS bytes -> dense float32 rows at dst (h*w values), (h, w) out
bool parse_fits(const std::vector<unsigned char>& raw, float* dst, int32_t* hw, std::string& err) {
  size_t off = 0;
  bool end = false;
  int bitpix = 0, naxis = -1, n1 = 0, n2 = 0;
  double bscale = 1.0, bzero = 0.0;
  while (!end) {
    if (raw.size() < off + 2880) { err = "truncated FITS header"; return false; }
    for (int i = 0; i < 2880 && !end; i += 80) {
      const char* card = reinterpret_cast<const char*>(raw.data() + off + i);
      std::string key(card, 8);
      key.erase(key.find_last_not_of(' ') + 1);
      if (key == "END") { end = true; break; }
      if (card[8] != '=' || card[9] != ' ') continue;
      std::string val(card + 10, 70);
      const size_t slash = val.find('/');
      if (slash != std::string::npos) val.erase(slash);
      const char* v = val.c_str();
      if (key == "BITPIX") bitpix = atoi(v);
      else if (key == "NAXIS") naxis = atoi(v);
      else if (key == "NAXIS1") n1 = atoi(v);
      else if (key == "NAXIS2") n2 = atoi(v);
      else if (key == "BSCALE") bscale = atof(v);
      else if (key == "BZERO") bzero = atof(v);
    }
    off += 2880;
  }
  if (naxis != 2) { err = "expected a 2-D FITS image, NAXIS=" + std::to_string(naxis); return false; }
  const int w = n1, h = n2;
  if (h < 1 || w < 1 || h > kMaxSide || w > kMaxSide) {
    err = "cutout has shape (" + std::to_string(h) + ", " + std::to_string(w) + "); expected at most 63x63";
    return false;
  }
  const int bytes = bitpix < 0 ? -bitpix / 8 : bitpix / 8;
  if (!(bitpix == -32 || bitpix == -64 || bitpix == 16 || bitpix == 32 || bitpix == 8)) { err = "unsupported BITPIX " + std::to_string(bitpix); return false; }
  const size_t count = (size_t)h * w;
  if (raw.size() < off + count * bytes) { err = "truncated FITS data"; return false; }
  const unsigned char* p = raw.data() + off;
  for (size_t i = 0; i < count; ++i, p += bytes) {
    double v;
    switch (bitpix) {
      case -32: dst[i] = load_be<float>(p); continue;                       // no scaling for floating-point pixels
      case -64: dst[i] = (float)load_be<double>(p); continue;
      case 16: v = (double)load_be<int16_t>(p); break;
      case 32: v = (double)load_be<int32_t>(p); break;
      default: v = (double)*p; break;
    }
    dst[i] = (float)(v * bscale + bzero);
  }
  hw[0] = h; hw[1] = w;
  return true;
}
}  // namespace
}  // namespace btsb

using namespace btsb;

// blobs[i] / sizes[i]: the gzipped FITS bytes of stamp i (3 per alert: science, template, difference).
// stamps: [n, 63*63] float32 (each stamp dense, row-major with its own width, at the start of its slot; the rest of the
// slot is left untouched), hw: [n, 2] int32.  threads <= 0: one per hardware thread, at most 32.  Host memory only.
extern "C" int btsb_ingest_fits_gz(const unsigned char* const* blobs, const int64_t* sizes, int64_t n, float* stamps,
                                   int32_t* hw, int threads) {
  BTSB_REQUIRE(n >= 0, "ingest: n < 0");
  if (n == 0) return BTSB_OK;
  BTSB_REQUIRE(blobs && sizes && stamps && hw, "ingest: null pointer");
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 32) nt = 32;
  if ((int64_t)nt > n) nt = (int)n;
  std::atomic<int64_t> next{0}, bad{-1};
  std::vector<std::string> errs(nt);
  auto work = [&](int tix) {
    std::vector<unsigned char> raw;
    std::string err;
    for (;;) {
      const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n || bad.load(std::memory_order_relaxed) >= 0) return;
      if (!blobs[i] || sizes[i] <= 0 || !inflate_gzip(blobs[i], (size_t)sizes[i], raw, err) ||
          !parse_fits(raw, stamps + i * (int64_t)(kMaxSide * kMaxSide), hw + 2 * i, err)) {
        if (err.empty()) err = "empty stamp";
        int64_t expect = -1;
        if (bad.compare_exchange_strong(expect, i)) errs[tix] = err;
        return;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  if (bad.load() >= 0) {
    std::string msg;
    for (auto& e : errs) if (!e.empty()) msg = e;
    set_error("ingest: stamp %lld: %s", (long long)bad.load(), msg.c_str());
    return BTSB_EINVAL;
  }
  return BTSB_OK;
}
