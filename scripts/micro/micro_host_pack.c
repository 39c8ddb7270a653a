// Host-side fp32 -> bf16 packing rate (would the e2e path gain from halving the H2D bytes on the host?):
//   gcc -O3 -mavx2 -pthread -o micro_host_pack micro_host_pack.c && ./micro_host_pack
// Converts a 391 MB fp32 buffer (one 8192-alert step) to bf16 (round to nearest even) with 1..N threads and prints GB/s
// (bytes read + written) and ms per step; also a plain memcpy of the same buffer for scale.
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static const uint32_t* g_src;
static uint16_t* g_dst;
static size_t g_n;
static int g_threads;
static int g_mode;   // 0 = pack, 1 = memcpy

static void pack(const uint32_t* s, uint16_t* d, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u = s[i];
    u += 0x7fffu + ((u >> 16) & 1u);
    d[i] = (uint16_t)(u >> 16);
  }
}
static void* worker(void* arg) {
  const long t = (long)arg;
  const size_t per = (g_n / g_threads + 63) & ~(size_t)63;
  const size_t lo = t * per, hi = lo + per < g_n ? lo + per : g_n;
  if (lo >= hi) return NULL;
  if (g_mode == 0) pack(g_src + lo, g_dst + lo, hi - lo);
  else memcpy((uint32_t*)g_dst + lo / 2, g_src + lo / 2, (hi - lo) * 2);
  return NULL;
}
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(void) {
  const size_t n = (size_t)8192 * 63 * 63 * 3;
  uint32_t* src = aligned_alloc(4096, n * 4);
  uint16_t* dst = aligned_alloc(4096, n * 4);
  for (size_t i = 0; i < n; ++i) { float f = 0.016f + 1e-6f * (float)(i % 1000); memcpy(&src[i], &f, 4); }
  memset(dst, 0, n * 4);
  g_src = src; g_dst = dst; g_n = n;
  const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
  printf("online cpus %ld, buffer %.1f MB fp32\n", ncpu, n * 4 / 1e6);
  for (g_mode = 0; g_mode < 2; ++g_mode) {
    for (int T = 1; T <= 64 && T <= 2 * ncpu; T *= 2) {
      g_threads = T;
      double best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        pthread_t th[64];
        const double t0 = now();
        for (long t = 0; t < T; ++t) pthread_create(&th[t], NULL, worker, (void*)t);
        for (long t = 0; t < T; ++t) pthread_join(th[t], NULL);
        const double dt = now() - t0;
        if (dt < best) best = dt;
      }
      const double bytes = g_mode == 0 ? n * 6.0 : n * 4.0;   // pack: 4 read + 2 written; memcpy (half the buffer): 2 + 2
      printf("%s threads %2d: %.2f ms  %.1f GB/s\n", g_mode == 0 ? "pack  " : "memcpy", T, best * 1e3, bytes / best / 1e9);
    }
  }
  return 0;
}
