// Dev micro-benchmark: host-side bit-packing rate (uint8 != 0 -> 1 bit) with T threads, AVX2, with and without
// software prefetch.   gcc -O3 -mavx2 -pthread tools/pack_bench.c -o tools/_bin/pack_bench
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static const uint8_t* g_in; static uint8_t* g_out; static size_t g_n; static int g_t, g_variant, g_dist;
static void* work(void* arg) {
  const int id = (int)(intptr_t)arg;
  const size_t per = (g_n / 64 / g_t) * 64, lo = per * id, hi = id == g_t - 1 ? g_n : lo + per;
  const __m256i z = _mm256_setzero_si256();
  if (g_variant == 0) {
    for (size_t i = lo; i + 32 <= hi; i += 32) {
      const __m256i v = _mm256_loadu_si256((const __m256i*)(g_in + i));
      const uint32_t m = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, z));
      memcpy(g_out + i / 8, &m, 4);
    }
  } else {
    for (size_t i = lo; i + 64 <= hi; i += 64) {
      if (g_variant == 1) _mm_prefetch((const char*)(g_in + i + g_dist), _MM_HINT_NTA);
      else if (g_variant == 2) _mm_prefetch((const char*)(g_in + i + g_dist), _MM_HINT_T0);
      else _mm_prefetch((const char*)(g_in + i + g_dist), _MM_HINT_T2);
      const __m256i a = _mm256_loadu_si256((const __m256i*)(g_in + i));
      const __m256i b = _mm256_loadu_si256((const __m256i*)(g_in + i + 32));
      const uint64_t m = (uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z))) |
                         ((uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z))) << 32);
      memcpy(g_out + i / 8, &m, 8);
    }
  }
  return 0;
}
static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int main(int argc, char** argv) {
  g_n = (size_t)512 * 512 * 512;
  uint8_t* in = aligned_alloc(4096, g_n + 8192); uint8_t* out = aligned_alloc(64, g_n / 8);
  for (size_t i = 0; i < g_n; i++) in[i] = (i * 2654435761u >> 27) == 0;
  memset(out, 0, g_n / 8);
  g_in = in; g_out = out;
  const int ts[] = {6, 8, 12, 15, 16};
  const int variants[][2] = {{0, 0}, {2, 1024}, {2, 2048}, {2, 3072}, {2, 4096}, {2, 8192}, {3, 2048}, {3, 8192}, {3, 16384}};
  for (int v = 0; v < 9; v++) {
    g_variant = variants[v][0]; g_dist = variants[v][1];
    for (int k = 0; k < 5; k++) {
      g_t = ts[k];
      double best = 1e9;
      for (int rep = 0; rep < 8; rep++) {
        pthread_t th[64];
        const double t0 = now();
        for (int i = 0; i < g_t; i++) pthread_create(&th[i], 0, work, (void*)(intptr_t)i);
        for (int i = 0; i < g_t; i++) pthread_join(th[i], 0);
        const double dt = now() - t0;
        if (dt < best) best = dt;
      }
      printf("variant %d dist %4d threads %2d: %.3f ms -> %.1f GB/s\n", g_variant, g_dist, g_t, best * 1e3, g_n / best / 1e9);
    }
  }
  return 0;
}
