// xs_hostpack.cpp — host side of the bit-packed upload (plain C++, compiled by g++; no CUDA here).
//
// A boolean volume crosses PCIe as one BYTE per voxel when uploaded as the reference hands it over (uint8 / bool
// ndarray): 128 MiB for 512^3, 2.4 ms at the ~54 GB/s a B200 host link sustains — three times the 0.78 ms the sweep
// itself takes.  The host cores can turn those bytes into bits at memory speed (~155 GB/s with 12-16 threads on the
// bench box: 0.85 ms for 128 MiB), after which only 16 MiB cross the link.  xs_pack_bits_parts does that with a
// persistent thread pool, reporting finished parts so that their copies can start while the rest is being packed; the
// device then builds its occupancy bytes from the bits (ingest_bits_*_kernel, xs_device.cu).  This is a format
// conversion only: no part of the cross-section algorithm runs on the host.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#include <pthread.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif

namespace {

// bytes [lo, hi) of `in` -> bits [lo, hi) of `out` (bit i of the stream = byte i != 0; lo is a multiple of 64)
void pack_range_scalar(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  size_t i = lo;
  for (; i + 8 <= hi; i += 8) {
    uint64_t x;
    memcpy(&x, in + i, 8);
    // per byte: 0x80 set iff byte != 0, then gather the eight flags into one byte
    x = ((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x;
    x = (x >> 7) & 0x0101010101010101ull;
    out[i >> 3] = (uint8_t)((x * 0x0102040810204080ull) >> 56);
  }
  if (i < hi) {
    uint8_t b = 0;
    for (size_t j = i; j < hi; j++) b |= (uint8_t)((in[j] != 0) << (j - i));
    out[i >> 3] = b;
  }
}

#if defined(__x86_64__)
// 64 voxels per step, with a software prefetch 8 KiB ahead: one core's demand loads (and the hardware streamer behind
// them) keep too few cache lines in flight to fill its share of the memory bandwidth — on the bench box (16 Xeon cores)
// the far prefetch lifts the pool from ~110-120 GB/s to ~165 GB/s (tools/pack_bench.c: distances of 1-4 KiB gain less,
// prefetchnta loses).  Prefetches past the end of the image are harmless (they never fault).
// The bit stream is written with non-temporal stores: it is read next by the copy engine, not by a core, and an ordinary
// store would first READ each output line for ownership — 1/8 more traffic on a loop that is bound by host memory
// (tools/pack_bench2.c on the bench box, 14 threads: 0.77 -> 0.73 ms best, 1.10 -> 0.84 ms mean; huge pages and four
// interleaved streams per thread gain nothing).  The fence at the end makes the stores visible before the part is
// reported done (its host-to-device copy is enqueued right then).
__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  const __m256i z = _mm256_setzero_si256();
  size_t i = lo;
  for (; i + 64 <= hi; i += 64) {
    _mm_prefetch(reinterpret_cast<const char*>(in + i + 8192), _MM_HINT_T2);
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i + 32));
    const uint64_t m = (uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, z))) |
                       ((uint64_t)(~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, z))) << 32);
    _mm_stream_si64(reinterpret_cast<long long*>(out + (i >> 3)), (long long)m);
  }
  _mm_sfence();
  for (; i + 32 <= hi; i += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i));
    const uint32_t m = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, z));
    memcpy(out + (i >> 3), &m, 4);
  }
  if (i < hi) pack_range_scalar(in, i, hi, out);
}
// the same with AVX-512BW where the CPU has it: one load and one vptestmb per 64 voxels
__attribute__((target("avx512bw,avx512f"))) void pack_range_avx512(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  size_t i = lo;
  for (; i + 64 <= hi; i += 64) {
    _mm_prefetch(reinterpret_cast<const char*>(in + i + 8192), _MM_HINT_T2);
    const __m512i a = _mm512_loadu_si512(reinterpret_cast<const void*>(in + i));
    _mm_stream_si64(reinterpret_cast<long long*>(out + (i >> 3)), (long long)_mm512_test_epi8_mask(a, a));
  }
  _mm_sfence();
  if (i < hi) pack_range_scalar(in, i, hi, out);
}
#endif

void pack_range(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
#if defined(__x86_64__)
  static const bool avx512 = __builtin_cpu_supports("avx512bw") && !getenv("XS3D_NO_AVX512") && !getenv("XS3D_NO_AVX2");
  static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("XS3D_NO_AVX2");
  if (avx512) { pack_range_avx512(in, lo, hi, out); return; }
  if (avx2) { pack_range_avx2(in, lo, hi, out); return; }
#endif
  pack_range_scalar(in, lo, hi, out);
}

// A small persistent pool: workers sleep on a condition variable between jobs (creating 16 threads per call would cost
// more than the packing itself).  A job is cut into 256 KiB grains handed out through an atomic cursor, so a worker
// that loses its core for a while (the box also runs the CUDA driver's and the caller's other threads) delays the job
// by one grain, not by its whole share; the calling thread packs grains too.
constexpr size_t GRAIN = (size_t)256 << 10;   // a multiple of 64

typedef void (*PartDone)(void* ctx, size_t lo, size_t hi);

class Pool {
 public:
  static constexpr int MAX_PARTS = 16;
  explicit Pool(int n) : n_(n) {
    for (int i = 0; i < n_ - 1; i++) threads_.emplace_back([this] { loop(); });
#if defined(__linux__)
    // dev knob XS3D_PIN_POOL=first_cpu: worker i runs on cpu first_cpu + i only (keeps the pool off the caller's cores)
    if (const char* e = getenv("XS3D_PIN_POOL")) {
      const int first = atoi(e);
      for (int i = 0; i < (int)threads_.size(); i++) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(first + i, &set);
        pthread_setaffinity_np(threads_[i].native_handle(), sizeof(set), &set);
      }
    }
#endif
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int size() const { return n_; }
  // Pack [lo, hi) as ONE job (one wake-up of the workers) reported in `nparts` consecutive parts: as soon as every grain
  // of the next part is packed, the CALLING thread runs done(ctx, part_lo, part_hi) — the upload enqueues that part's
  // host-to-device copy there, so the transfers run beside the packing of the rest.
  void run(const uint8_t* in, size_t lo, size_t hi, uint8_t* out, int nparts, PartDone done, void* ctx) {
    const size_t grains = (hi - lo + GRAIN - 1) / GRAIN;
    nparts = (int)std::max<size_t>(1, std::min<size_t>({ (size_t)nparts, (size_t)MAX_PARTS, grains }));
    {
      std::lock_guard<std::mutex> lk(mu_);
      in_ = in; out_ = out; lo_ = lo; hi_ = hi; grains_ = grains; nparts_ = nparts;
      for (int p = 0; p < nparts; p++) left_[p].store(part_begin(p + 1) - part_begin(p), std::memory_order_relaxed);
      next_.store(0, std::memory_order_relaxed);
      pending_ = n_ - 1; gen_++;
    }
    cv_.notify_all();
    int reported = 0;
    while (reported < nparts) {
      if (left_[reported].load(std::memory_order_acquire) == 0) {
        if (done) done(ctx, lo + GRAIN * part_begin(reported), std::min(hi, lo + GRAIN * part_begin(reported + 1)));
        reported++;
      } else if (!work_one(in, lo, hi, out)) {
        std::this_thread::yield();     // every grain is handed out; the last ones are still being packed
      }
    }
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  size_t part_begin(int p) const { return grains_ * (size_t)p / (size_t)nparts_; }
  int part_of(size_t g) const {
    int p = (int)(g * (size_t)nparts_ / grains_);
    while (p + 1 < nparts_ && part_begin(p + 1) <= g) p++;
    while (p > 0 && part_begin(p) > g) p--;
    return p;
  }
  bool work_one(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
    const size_t g = next_.fetch_add(1, std::memory_order_relaxed);
    if (g >= grains_) return false;
    const size_t a = lo + GRAIN * g;
    pack_range(in, a, std::min(hi, a + GRAIN), out);
    left_[part_of(g)].fetch_sub(1, std::memory_order_release);
    return true;
  }
  void loop() {
#if defined(__linux__)
    // The workers yield to the caller's other threads: the thread that launches and waits for the sweeps must not lose its
    // core to a packer for a scheduler tick (on a 16-core host with 14 packers its sweep call took 0.81 ms at the median
    // and 1.5 ms at the 90th percentile, 0.75 / 0.77 ms with 8 packers).  Per-thread nice value (Linux: setpriority on the
    // thread id); XS3D_POOL_NICE overrides, 0 switches it off.
    {
      static const int nice_by = getenv("XS3D_POOL_NICE") ? atoi(getenv("XS3D_POOL_NICE")) : 10;
      if (nice_by > 0) setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), nice_by);
    }
#endif
    uint64_t seen = 0;
    for (;;) {
      const uint8_t* in; uint8_t* out; size_t lo, hi;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        in = in_; out = out_; lo = lo_; hi = hi_;
      }
      while (work_one(in, lo, hi, out)) {}
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const uint8_t* in_ = nullptr; uint8_t* out_ = nullptr;
  size_t lo_ = 0, hi_ = 0, grains_ = 0;
  int nparts_ = 1;
  std::atomic<size_t> next_{0};
  std::atomic<size_t> left_[MAX_PARTS];
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

std::mutex g_pool_mu;
Pool* g_pool = nullptr;

// Threads for the pool: the cores this process may run on (affinity mask, not the machine's core count), shared among
// the ranks torchrun started on this node (LOCAL_WORLD_SIZE), minus two (the caller's other threads and the CUDA
// driver's need somewhere to run: on a 16-core B200 box the pipelined sweep takes 1.38 ms with 14 threads, 1.66 ms with
// 16, 1.79 ms with 8), at most 16.  XS3D_HOST_THREADS overrides.
int pool_threads() {
  if (const char* e = getenv("XS3D_HOST_THREADS")) return std::max(1, atoi(e));
  unsigned hc = std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = (unsigned)CPU_COUNT(&set);
#endif
  unsigned ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(e));
  const unsigned share = (hc ? hc : 4u) / ranks;
  return (int)std::min<unsigned>(share > 3u ? share - 2u : 1u, 16u);
}

}  // namespace

extern "C" int xs_pack_threads(void);

// Pack bytes [lo, hi) of `in` (lo a multiple of 64) into the bit stream `out` (bit i <-> byte i != 0), reporting progress
// in `nparts` consecutive parts through done(ctx, part_lo, part_hi), called on the calling thread (may be null).  Thread
// safe: concurrent callers take turns on the pool.  Returns the number of threads used.
extern "C" int xs_pack_bits_parts(const uint8_t* in, size_t lo, size_t hi, uint8_t* out, int nparts,
                                  void (*done)(void* ctx, size_t lo, size_t hi), void* ctx) {
  if (hi <= lo) return 0;
  if (hi - lo < (1u << 20)) {
    pack_range(in, lo, hi, out);
    if (done) done(ctx, lo, hi);
    return 1;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool) g_pool = new Pool(xs_pack_threads());
  g_pool->run(in, lo, hi, out, nparts, done, ctx);
  return g_pool->size();
}

extern "C" int xs_pack_bits(const uint8_t* in, size_t lo, size_t hi, uint8_t* out) {
  return xs_pack_bits_parts(in, lo, hi, out, 1, nullptr, nullptr);
}

// Threads xs_pack_bits would use (decides whether the packed upload is worth it: xs_device.cu, use_packed_upload).
extern "C" int xs_pack_threads(void) {
  static const int n = pool_threads();
  return n;
}
