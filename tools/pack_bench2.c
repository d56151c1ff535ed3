// Dev micro-benchmark 2: host-side bit-packing rate with T threads — what else could lift it?
//   page = 0: plain pages, 1: MADV_HUGEPAGE on the image before first touch
//   kind = 0: AVX2 + prefetcht2 8 KiB ahead (the library's loop), 1: AVX-512BW vptestmb (when the CPU has it),
//          2: AVX2, each thread walks 4 interleaved streams, 3: AVX2 without any software prefetch (2, 3: skipped now),
//          4: AVX2 + prefetch + non-temporal stores of the bit stream
// gcc -O3 -pthread tools/pack_bench2.c -o tools/_bin/pack_bench2
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
static const uint8_t* g_in; static uint8_t* g_out; static size_t g_n; static int g_t, g_kind;
__attribute__((target("avx2"))) static void pack_avx2(const uint8_t* in, size_t lo, size_t hi, uint8_t* out, int pf) {
  const __m256i z = _mm256_setzero_si256();
  for (size_t i = lo; i + 64 <= hi; i += 64) {
    if (pf) _mm_prefetch((const char*)(in + i + 8192), _MM_HINT_T2);
    const __m256i a = _mm256_loadu_si256((const __m256i*)(in + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(in + i + 32));
    const uint64_t m = (uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z))) |
                       ((uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z))) << 32);
    memcpy(out + i / 8, &m, 8);
  }
}
__attribute__((target("avx2"))) static void pack_avx2_nt(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  const __m256i z = _mm256_setzero_si256();
  for (size_t i = lo; i + 64 <= hi; i += 64) {
    _mm_prefetch((const char*)(in + i + 8192), _MM_HINT_T2);
    const __m256i a = _mm256_loadu_si256((const __m256i*)(in + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(in + i + 32));
    const uint64_t m = (uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z))) |
                       ((uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z))) << 32);
    _mm_stream_si64((long long*)(out + i / 8), (long long)m);
  }
  _mm_sfence();
}
__attribute__((target("avx512bw,avx512f"))) static void pack_avx512(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  for (size_t i = lo; i + 64 <= hi; i += 64) {
    _mm_prefetch((const char*)(in + i + 8192), _MM_HINT_T2);
    const __m512i a = _mm512_loadu_si512((const void*)(in + i));
    const uint64_t m = (uint64_t)_mm512_test_epi8_mask(a, a);
    memcpy(out + i / 8, &m, 8);
  }
}
__attribute__((target("avx2"))) static void pack_avx2_4s(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  const __m256i z = _mm256_setzero_si256();
  const size_t q = ((hi - lo) / 4) & ~(size_t)63;
  for (size_t i = 0; i + 64 <= q; i += 64) {
    for (int s = 0; s < 4; s++) {
      const size_t j = lo + s * q + i;
      _mm_prefetch((const char*)(in + j + 4096), _MM_HINT_T2);
      const __m256i a = _mm256_loadu_si256((const __m256i*)(in + j));
      const __m256i b = _mm256_loadu_si256((const __m256i*)(in + j + 32));
      const uint64_t m = (uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z))) |
                         ((uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z))) << 32);
      memcpy(out + j / 8, &m, 8);
    }
  }
  pack_avx2(in, lo + 4 * q, hi, out, 1);
}
static void* work(void* arg) {
  const int id = (int)(intptr_t)arg;
  const size_t per = (g_n / 64 / g_t) * 64, lo = per * id, hi = id == g_t - 1 ? g_n : lo + per;
  if (g_kind == 0) pack_avx2(g_in, lo, hi, g_out, 1);
  else if (g_kind == 1) pack_avx512(g_in, lo, hi, g_out);
  else if (g_kind == 2) pack_avx2_4s(g_in, lo, hi, g_out);
  else if (g_kind == 3) pack_avx2(g_in, lo, hi, g_out, 0);
  else pack_avx2_nt(g_in, lo, hi, g_out);
  return 0;
}
static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int main(int argc, char** argv) {
  g_n = (size_t)512 * 512 * 512;
  const int have512 = __builtin_cpu_supports("avx512bw");
  printf("avx512bw: %d\n", have512);
  uint8_t* out = aligned_alloc(64, g_n / 8);
  memset(out, 0, g_n / 8);
  g_out = out;
  const int nt = argc > 1 ? atoi(argv[1]) : 15;
  for (int page = 0; page < 2; page++) {
    uint8_t* in = mmap(0, g_n + (4 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    in = (uint8_t*)(((uintptr_t)in + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    if (page) madvise(in, g_n + (1 << 20), MADV_HUGEPAGE);
    for (size_t i = 0; i < g_n; i++) in[i] = (i * 2654435761u >> 27) == 0;
    g_in = in;
    for (g_kind = 0; g_kind < 5; g_kind++) {
      if (g_kind == 2 || g_kind == 3) continue;
      if (g_kind == 1 && !have512) continue;
      const int ts[] = {nt - 3, nt - 1, nt, nt + 1};
      for (int k = 0; k < 4; k++) {
        g_t = ts[k];
        double best = 1e9, sum = 0;
        for (int rep = 0; rep < 16; rep++) {
          pthread_t th[128];
          const double t0 = now();
          for (int i = 0; i < g_t; i++) pthread_create(&th[i], 0, work, (void*)(intptr_t)i);
          for (int i = 0; i < g_t; i++) pthread_join(th[i], 0);
          const double dt = now() - t0;
          if (dt < best) best = dt;
          if (rep >= 2) sum += dt;
        }
        printf("page %d kind %d threads %2d: best %.3f ms mean %.3f ms -> %.1f GB/s\n", page, g_kind, g_t, best * 1e3, sum / 14 * 1e3, g_n / best / 1e9);
      }
    }
  }
  return 0;
}
