// xs_slice.cu — xs3d.slice on the GPU: dense validity bitmap -> per-row runs -> 1-D row scan -> gather.
// See xs_slice.cuh for the algorithm and the reference lines it replaces
// (xs3d.hpp:1473-1624, fastxs3d.cpp:119-216).
//
// Two execution shapes:
//   * slice_cta_kernel   : one CTA per plane, window <= 64x64 cells (then <= 256x256 for the planes that pass
//     flags), everything in shared memory — the batched "10k cropped planes" path (BASELINE config 4);
//   * big path           : whole-GPU kernels per stage for windows up to the full lattice
//     (un-cropped planes, up to ~7000^2 cells): valid -> rows -> scan -> gather.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <algorithm>

#include "xs_slice.cuh"
#include "xs_device.h"
#include "../../include/xs3d_b200.h"

using namespace xs;

template <typename T>
__device__ __forceinline__ T load_label(const void* labels, size_t loc) {
  return __ldg(reinterpret_cast<const T*>(labels) + loc);
}

__device__ __forceinline__ void copy_label(void* out, size_t oi, const void* labels, size_t loc, int itemsize) {
  switch (itemsize) {
    case 1: reinterpret_cast<uint8_t*>(out)[oi] = load_label<uint8_t>(labels, loc); break;
    case 2: reinterpret_cast<uint16_t*>(out)[oi] = load_label<uint16_t>(labels, loc); break;
    case 4: reinterpret_cast<uint32_t*>(out)[oi] = load_label<uint32_t>(labels, loc); break;
    default: reinterpret_cast<unsigned long long*>(out)[oi] = load_label<unsigned long long>(labels, loc); break;
  }
}
__device__ __forceinline__ void zero_label(void* out, size_t oi, int itemsize) {
  switch (itemsize) {
    case 1: reinterpret_cast<uint8_t*>(out)[oi] = 0; break;
    case 2: reinterpret_cast<uint16_t*>(out)[oi] = 0; break;
    case 4: reinterpret_cast<uint32_t*>(out)[oi] = 0u; break;
    default: reinterpret_cast<unsigned long long*>(out)[oi] = 0ull; break;
  }
}

// ------------------------------------------------------------------------------------------------
// big path
// ------------------------------------------------------------------------------------------------
// stage 1: one warp per 32 cells of a window row -> one bitmap word
__global__ void __launch_bounds__(256) slice_valid_kernel(SliceCtx s, uint32_t* __restrict__ bits, int words) {
  const size_t total = (size_t)words * s.wh;
  const int lane = threadIdx.x & 31;
  for (size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += ((size_t)gridDim.x * blockDim.x) >> 5) {
    const int row = (int)(wi / words), col = (int)(wi % words) * 32 + lane;
    const bool ok = (col < s.ww) && slice_cell(s, s.wx0 + col, s.wy0 + row, false, nullptr);
    const uint32_t w = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) bits[wi] = w;
  }
}

// stage 2: per-row summary
__global__ void __launch_bounds__(256) slice_rows_kernel(const uint32_t* __restrict__ bits, int words, int nrows, RowRuns* __restrict__ rows) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < nrows) rows[row] = row_runs(bits + (size_t)row * words, words);
}

// stage 3: 1-D scan (one thread; <= ~7000 rows)
__global__ void slice_scan_kernel(const RowRuns* __restrict__ rows, int nrows, int cx, int cy, SliceResult* __restrict__ res) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *res = scan_rows(rows, nrows, cx, cy);
}

// exact fallback: iterate  R <- fill_runs(V, R | vertical neighbours of R)  to a fixpoint
__global__ void __launch_bounds__(256) slice_fill_iter_kernel(const uint32_t* __restrict__ V, uint32_t* __restrict__ R, int words, int nrows, int* __restrict__ changed) {
  const size_t total = (size_t)words * nrows;
  for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < total; wi += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(wi / words), col = (int)(wi % words);
    const uint32_t v = V[wi];
    if (!v) continue;
    const uint32_t old = R[wi];
    uint32_t seeds = old;
    if (row > 0) seeds |= R[wi - words];
    if (row + 1 < nrows) seeds |= R[wi + words];
    if (col > 0) seeds |= R[wi - 1] >> 31;             // carry from the word on the left
    if (col + 1 < words) seeds |= R[wi + 1] << 31;     // and from the right
    seeds &= v;
    // occluded fill inside the word, both directions (Kogge-Stone)
    uint32_t g = seeds, p = v;
    g |= p & (g << 1); p &= p << 1; g |= p & (g << 2); p &= p << 2; g |= p & (g << 4); p &= p << 4;
    g |= p & (g << 8); p &= p << 8; g |= p & (g << 16);
    p = v;
    g |= p & (g >> 1); p &= p >> 1; g |= p & (g >> 2); p &= p >> 2; g |= p & (g >> 4); p &= p >> 4;
    g |= p & (g >> 8); p &= p >> 8; g |= p & (g >> 16);
    if (g != old) { atomicOr(&R[wi], g); *changed = 1; }
  }
}

__global__ void slice_seed_kernel(const uint32_t* V, uint32_t* R, int words, int cx, int cy) {
  const size_t wi = (size_t)cy * words + (cx >> 5);
  R[wi] = V[wi] & (1u << (cx & 31));
}

// bbox of a bitmap from its row summaries (fallback path)
__global__ void slice_bbox_kernel(const RowRuns* __restrict__ rows, int nrows, SliceResult* __restrict__ res) {
  if (threadIdx.x || blockIdx.x) return;
  SliceResult r;
  r.y0 = 0x7fffffff; r.y1 = -1; r.x0 = 0x7fffffff; r.x1 = -1; r.fallback = 1;
  for (int y = 0; y < nrows; y++) {
    if (rows[y].runs == 0) continue;
    if (y < r.y0) r.y0 = y;
    r.y1 = y;
    if (rows[y].a < r.x0) r.x0 = rows[y].a;
    if (rows[y].b > r.x1) r.x1 = rows[y].b;
  }
  *res = r;
}

// stage 4: gather the cut-out (F-order [nx][ny]); bits == reached cells when masked by rows [y0,y1]
__global__ void __launch_bounds__(256) slice_gather_kernel(SliceCtx s, const uint32_t* __restrict__ bits, int words, SliceResult r,
                                                           const void* __restrict__ labels, int itemsize, int c_order, void* __restrict__ out) {
  const int nx = r.x1 - r.x0 + 1, ny = r.y1 - r.y0 + 1;
  const size_t total = (size_t)nx * ny;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = r.x0 + (int)(i % nx), y = r.y0 + (int)(i / nx);
    const bool on = (bits[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u;
    size_t loc;
    if (on && slice_cell(s, s.wx0 + x, s.wy0 + y, c_order != 0, &loc)) copy_label(out, i, labels, loc, itemsize);
    else zero_label(out, i, itemsize);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA path: one CTA per plane, window <= 64 x 64, then <= 256 x 256
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_W = 64, MED_W = 256;

struct SliceBatchArgs {
  const float* pos; const float* nrm;
  float wx, wy, wz;
  int positive; float crop;
  int sx, sy, sz;
  const void* labels; int itemsize; int c_order;
  void* out; unsigned long long pitch_elems;
  long long* shapes; uint8_t* overflow;   // overflow: 1 = cut-out larger than pitch, 2 = window too large for this kernel
  uint32_t n;
};

// One CTA per plane, window <= W x W cells, everything in shared memory.  W = 64 (128 threads): the batched "10k
// cropped planes" shape; W = 256 (256 threads): the planes the first pass flags as too wide (overflow == 2).
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS) slice_cta_kernel(SliceBatchArgs a, int only_flagged) {
  constexpr int WORDS = W / 32;                 // bitmap words per window row
  constexpr int NW = W * WORDS;                 // words in the window
  constexpr int K = (NW + THREADS - 1) / THREADS;
  __shared__ uint32_t V[NW], R[NW];
  __shared__ RowRuns rows[W];
  __shared__ SliceResult res;
  __shared__ int changed;
  for (uint32_t q = blockIdx.x; q < a.n; q += gridDim.x) {
    if (only_flagged && a.overflow[q] != 2) continue;
    SliceCtx s;
    slice_init(s, a.sx, a.sy, a.sz, mk(a.pos[3 * q], a.pos[3 * q + 1], a.pos[3 * q + 2]), mk(a.nrm[3 * q], a.nrm[3 * q + 1], a.nrm[3 * q + 2]),
               mk(a.wx, a.wy, a.wz), a.positive != 0, a.crop);
    char* outq = reinterpret_cast<char*>(a.out) + (size_t)q * a.pitch_elems * a.itemsize;
    __syncthreads();
    if (!s.active) {
      // bbox stays at its initial value: 1x1 cut-out holding 0 (fastxs3d.cpp:179-183)
      if (threadIdx.x == 0) {
        a.shapes[2 * q] = 1; a.shapes[2 * q + 1] = 1; a.overflow[q] = (a.pitch_elems < 1) ? 1 : 0;
        if (a.pitch_elems >= 1) zero_label(outq, 0, a.itemsize);
      }
      continue;
    }
    if (s.ww > W || s.wh > W) {
      if (threadIdx.x == 0) { a.shapes[2 * q] = -1; a.shapes[2 * q + 1] = -1; a.overflow[q] = 2; }
      continue;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int wi = warp; wi < NW; wi += THREADS / 32) {
      const int row = wi / WORDS, col = (wi % WORDS) * 32 + lane;
      const bool ok = (row < s.wh) && (col < s.ww) && slice_cell(s, s.wx0 + col, s.wy0 + row, false, nullptr);
      const uint32_t w = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) V[wi] = w;
    }
    __syncthreads();
    for (int row = threadIdx.x; row < W; row += THREADS) rows[row] = row_runs(V + WORDS * row, WORDS);
    __syncthreads();
    const int cx = (int)(s.c - s.wx0), cy = (int)(s.c - s.wy0);
    if (threadIdx.x == 0) res = scan_rows(rows, s.wh, cx, cy);
    __syncthreads();
    const uint32_t* mask = V;
    if (res.fallback) {
      // exact fixpoint fill in shared memory
      for (int wi = threadIdx.x; wi < NW; wi += THREADS) R[wi] = 0u;
      __syncthreads();
      if (threadIdx.x == 0) R[cy * WORDS + (cx >> 5)] = V[cy * WORDS + (cx >> 5)] & (1u << (cx & 31));
      do {
        __syncthreads();
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        uint32_t gnew[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
          const int wi = threadIdx.x + k * THREADS;
          gnew[k] = 0u;
          if (wi >= NW) continue;
          const int row = wi / WORDS, col = wi % WORDS;
          const uint32_t v = V[wi], old = R[wi];
          uint32_t seeds = old;
          if (row > 0) seeds |= R[wi - WORDS];
          if (row + 1 < W) seeds |= R[wi + WORDS];
          if (col > 0) seeds |= R[wi - 1] >> 31;
          if (col + 1 < WORDS) seeds |= R[wi + 1] << 31;
          seeds &= v;
          uint32_t g = seeds, p = v;
          g |= p & (g << 1); p &= p << 1; g |= p & (g << 2); p &= p << 2; g |= p & (g << 4); p &= p << 4;
          g |= p & (g << 8); p &= p << 8; g |= p & (g << 16);
          p = v;
          g |= p & (g >> 1); p &= p >> 1; g |= p & (g >> 2); p &= p >> 2; g |= p & (g >> 4); p &= p >> 4;
          g |= p & (g >> 8); p &= p >> 8; g |= p & (g >> 16);
          gnew[k] = g;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
          const int wi = threadIdx.x + k * THREADS;
          if (wi < NW && gnew[k] != R[wi]) { R[wi] = gnew[k]; changed = 1; }
        }
        __syncthreads();
      } while (changed);
      for (int row = threadIdx.x; row < W; row += THREADS) rows[row] = row_runs(R + WORDS * row, WORDS);
      __syncthreads();
      if (threadIdx.x == 0) {
        SliceResult r;
        r.y0 = 0x7fffffff; r.y1 = -1; r.x0 = 0x7fffffff; r.x1 = -1; r.fallback = 1;
        for (int y = 0; y < s.wh; y++) {
          if (rows[y].runs == 0) continue;
          if (y < r.y0) r.y0 = y;
          r.y1 = y;
          if (rows[y].a < r.x0) r.x0 = rows[y].a;
          if (rows[y].b > r.x1) r.x1 = rows[y].b;
        }
        res = r;
      }
      __syncthreads();
      mask = R;
    }
    SliceResult r = res;
    if (r.y0 > r.y1) { r.x0 = r.x1 = cx; r.y0 = r.y1 = cy; }   // nothing reached: bbox stays on the centre cell
    const int nx = r.x1 - r.x0 + 1, ny = r.y1 - r.y0 + 1;
    const unsigned long long total = (unsigned long long)nx * ny;
    if (threadIdx.x == 0) { a.shapes[2 * q] = nx; a.shapes[2 * q + 1] = ny; a.overflow[q] = total > a.pitch_elems ? 1 : 0; }
    if (total > a.pitch_elems) continue;
    for (int i = threadIdx.x; i < (int)total; i += THREADS) {
      const int x = r.x0 + i % nx, y = r.y0 + i / nx;
      const bool on = (res.y0 <= res.y1) && ((mask[y * WORDS + (x >> 5)] >> (x & 31)) & 1u) && (res.fallback || (y >= res.y0 && y <= res.y1));
      size_t loc;
      if (on && slice_cell(s, s.wx0 + x, s.wy0 + y, a.c_order != 0, &loc)) copy_label(outq, i, a.labels, loc, a.itemsize);
      else zero_label(outq, i, a.itemsize);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
extern "C" int xs3d_volume_load_labels(xs3d_volume* v, const void* labels, int itemsize, int mem, int c_order) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8) return xs_fail(XS3D_ERR_ARG, "label itemsize must be 1, 2, 4 or 8");
  return xs_labels_set(v, labels, itemsize, mem, c_order);
}

// One plane through the big path.  Leaves the cut-out on the device in buffer 0 (sec_idx) and returns its shape.
static int slice_big(xs3d_volume* v, const float p[3], const float n[3], const float w[3], int positive, float crop,
                     int64_t shape[2], void** d_cut) {
  uint64_t dims[3];
  XS_RET(xs3d_volume_shape(v, dims));
  void* labels; int itemsize, c_order; bool loaded;
  xs_labels(v, &labels, &itemsize, &c_order, &loaded);
  if (!loaded) return xs_fail(XS3D_ERR_ARG, "volume has no labels loaded");
  cudaStream_t st = xs_stream(v);
  SliceCtx s;
  slice_init(s, (int)dims[0], (int)dims[1], (int)dims[2], mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), mk(w[0], w[1], w[2]), positive != 0, crop);
  if (!s.active) {
    shape[0] = 1; shape[1] = 1;
    void* cut;
    XS_RET(xs_buf(v, 0, 64, &cut));
    XS_CK(cudaMemsetAsync(cut, 0, 16, st));
    *d_cut = cut;
    return XS3D_OK;
  }
  const int words = (s.ww + 31) / 32;
  const size_t bitmap_bytes = (size_t)words * s.wh * 4;
  const size_t rows_bytes = (size_t)s.wh * sizeof(RowRuns);
  char* misc;
  XS_RET(xs_buf(v, 2, 2 * bitmap_bytes + rows_bytes + 1024, (void**)&misc));
  uint32_t* V = (uint32_t*)misc;
  uint32_t* R = (uint32_t*)(misc + bitmap_bytes);
  RowRuns* rows = (RowRuns*)(misc + 2 * bitmap_bytes);
  SliceResult* d_res = (SliceResult*)(misc + 2 * bitmap_bytes + rows_bytes);
  int* d_changed = (int*)(d_res + 1);
  const int sms = xs_sm_count(v);
  const size_t nwords = (size_t)words * s.wh;
  int grid = (int)std::min<size_t>((nwords * 32 + 255) / 256, (size_t)sms * 16);
  slice_valid_kernel<<<grid, 256, 0, st>>>(s, V, words);
  slice_rows_kernel<<<(s.wh + 255) / 256, 256, 0, st>>>(V, words, s.wh, rows);
  const int cx = (int)(s.c - s.wx0), cy = (int)(s.c - s.wy0);
  slice_scan_kernel<<<1, 32, 0, st>>>(rows, s.wh, cx, cy, d_res);
  XS_CK(cudaGetLastError());
  SliceResult* h_res = (SliceResult*)xs_pinned(v);
  XS_CK(cudaMemcpyAsync(h_res, d_res, sizeof(SliceResult), cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  SliceResult r = *h_res;
  const uint32_t* mask = V;
  if (r.fallback) {
    XS_CK(cudaMemsetAsync(R, 0, bitmap_bytes, st));
    slice_seed_kernel<<<1, 1, 0, st>>>(V, R, words, cx, cy);
    int* h_changed = (int*)(xs_pinned(v) + 16);
    int fgrid = (int)std::min<size_t>((nwords + 255) / 256, (size_t)sms * 16);
    for (;;) {
      XS_CK(cudaMemsetAsync(d_changed, 0, 4, st));
      slice_fill_iter_kernel<<<fgrid, 256, 0, st>>>(V, R, words, s.wh, d_changed);
      XS_CK(cudaMemcpyAsync(h_changed, d_changed, 4, cudaMemcpyDeviceToHost, st));
      XS_CK(cudaStreamSynchronize(st));
      if (!*h_changed) break;
    }
    slice_rows_kernel<<<(s.wh + 255) / 256, 256, 0, st>>>(R, words, s.wh, rows);
    slice_bbox_kernel<<<1, 32, 0, st>>>(rows, s.wh, d_res);
    XS_CK(cudaMemcpyAsync(h_res, d_res, sizeof(SliceResult), cudaMemcpyDeviceToHost, st));
    XS_CK(cudaStreamSynchronize(st));
    r = *h_res;
    mask = R;
  }
  bool nothing = r.y0 > r.y1;
  if (nothing) { r.x0 = r.x1 = cx; r.y0 = r.y1 = cy; }
  const int nx = r.x1 - r.x0 + 1, ny = r.y1 - r.y0 + 1;
  shape[0] = nx; shape[1] = ny;
  void* cut;
  XS_RET(xs_buf(v, 0, (size_t)nx * ny * itemsize + 64, &cut));
  if (nothing) {
    XS_CK(cudaMemsetAsync(cut, 0, (size_t)itemsize, st));
  } else {
    if (!r.fallback) {
      // rows outside [y0,y1] are not reached: the gather only visits the bbox, whose rows are all reached
    }
    const size_t total = (size_t)nx * ny;
    grid = (int)std::min<size_t>((total + 255) / 256, (size_t)sms * 16);
    slice_gather_kernel<<<grid, 256, 0, st>>>(s, mask, words, r, labels, itemsize, c_order, cut);
    XS_CK(cudaGetLastError());
  }
  *d_cut = cut;
  return XS3D_OK;
}

// fastxs3d.projection drop-in (two-call protocol, see the header)
static struct {
  bool valid = false;
  const void* labels = nullptr;
  uint64_t sx = 0, sy = 0, sz = 0;
  int itemsize = 0, c_order = 0, positive = 0;
  float p[3], n[3], w[3], crop = 0;
  int64_t shape[2];
  void* d_cut = nullptr;
  xs3d_volume* vol = nullptr;
} g_pending;

extern "C" int xs3d_projection(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int itemsize, int c_order,
                               const float pos[3], const float normal[3], const float anisotropy[3],
                               int standardize_basis, float crop_distance, void* out, int64_t shape[2]) {
  std::lock_guard<std::mutex> lk(xs_mutex());
  if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8) return xs_fail(XS3D_ERR_ARG, "label itemsize must be 1, 2, 4 or 8");
  if (crop_distance == 0.0f) { shape[0] = 0; shape[1] = 0; return XS3D_OK; }   // fastxs3d.cpp:160-162
  int ndev = 0;
  XS_CK(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return xs_fail(XS3D_ERR_CUDA, "no CUDA device");
  const bool same = g_pending.valid && g_pending.labels == labels && g_pending.sx == sx && g_pending.sy == sy && g_pending.sz == sz &&
                    g_pending.itemsize == itemsize && g_pending.c_order == c_order && g_pending.positive == standardize_basis &&
                    g_pending.crop == crop_distance && !memcmp(g_pending.p, pos, 12) && !memcmp(g_pending.n, normal, 12) &&
                    !memcmp(g_pending.w, anisotropy, 12);
  if (!(out && same)) {
    xs3d_volume* v;
    XS_RET(xs_default_volume(sx, sy, sz, &v));
    XS_RET(xs_labels_set(v, labels, itemsize, XS3D_HOST, c_order));
    g_pending.valid = false;
    XS_RET(slice_big(v, pos, normal, anisotropy, standardize_basis, crop_distance, g_pending.shape, &g_pending.d_cut));
    g_pending.valid = true; g_pending.labels = labels; g_pending.sx = sx; g_pending.sy = sy; g_pending.sz = sz;
    g_pending.itemsize = itemsize; g_pending.c_order = c_order; g_pending.positive = standardize_basis; g_pending.crop = crop_distance;
    memcpy(g_pending.p, pos, 12); memcpy(g_pending.n, normal, 12); memcpy(g_pending.w, anisotropy, 12);
    g_pending.vol = v;
  }
  shape[0] = g_pending.shape[0]; shape[1] = g_pending.shape[1];
  if (out) {
    const size_t bytes = (size_t)shape[0] * shape[1] * itemsize;
    cudaStream_t st = xs_stream(g_pending.vol);
    XS_CK(cudaMemcpyAsync(out, g_pending.d_cut, bytes, cudaMemcpyDeviceToHost, st));
    XS_CK(cudaStreamSynchronize(st));
    g_pending.valid = false;
  }
  return XS3D_OK;
}

extern "C" int xs3d_projection_batch(xs3d_volume* v, uint64_t n, const float* pos, const float* normal, const float anisotropy[3],
                                     int standardize_basis, float crop_distance, void* out, uint64_t pitch_elems,
                                     int64_t* shapes, uint8_t* overflow, int mem) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  void* labels; int itemsize, c_order; bool loaded;
  xs_labels(v, &labels, &itemsize, &c_order, &loaded);
  if (!loaded) return xs_fail(XS3D_ERR_ARG, "volume has no labels loaded");
  if (n == 0) return XS3D_OK;
  uint64_t dims[3];
  XS_RET(xs3d_volume_shape(v, dims));
  cudaStream_t st = xs_stream(v);
  if (crop_distance == 0.0f) {
    if (mem == XS3D_HOST) { memset(shapes, 0, n * 16); memset(overflow, 0, n); }
    else { XS_CK(cudaMemsetAsync(shapes, 0, n * 16, st)); XS_CK(cudaMemsetAsync(overflow, 0, n, st)); XS_CK(cudaStreamSynchronize(st)); }
    return XS3D_OK;
  }
  // staging for host callers
  const size_t out_bytes = (size_t)n * pitch_elems * itemsize;
  char* stage = nullptr;
  const float *d_pos = pos, *d_nrm = normal;
  void* d_out = out; long long* d_shapes = (long long*)shapes; uint8_t* d_over = overflow;
  if (mem == XS3D_HOST) {
    XS_RET(xs_buf(v, 1, n * 24 + n * 16 + n + out_bytes + 1024, (void**)&stage));
    XS_CK(cudaMemcpyAsync(stage, pos, n * 12, cudaMemcpyHostToDevice, st));
    XS_CK(cudaMemcpyAsync(stage + n * 12, normal, n * 12, cudaMemcpyHostToDevice, st));
    d_pos = (const float*)stage; d_nrm = (const float*)(stage + n * 12);
    d_shapes = (long long*)(stage + n * 24);
    d_out = stage + n * 24 + n * 16;
    d_over = (uint8_t*)(stage + n * 24 + n * 16 + ((out_bytes + 15) / 16) * 16);
  }
  SliceBatchArgs a;
  a.pos = d_pos; a.nrm = d_nrm; a.wx = anisotropy[0]; a.wy = anisotropy[1]; a.wz = anisotropy[2];
  a.positive = standardize_basis; a.crop = crop_distance;
  a.sx = (int)dims[0]; a.sy = (int)dims[1]; a.sz = (int)dims[2];
  a.labels = labels; a.itemsize = itemsize; a.c_order = c_order;
  a.out = d_out; a.pitch_elems = pitch_elems; a.shapes = d_shapes; a.overflow = d_over; a.n = (uint32_t)n;
  const int grid = (int)std::min<uint64_t>(n, (uint64_t)xs_sm_count(v) * 8);
  slice_cta_kernel<SMALL_W, 128><<<grid, 128, 0, st>>>(a, 0);
  XS_CK(cudaGetLastError());
  slice_cta_kernel<MED_W, 256><<<grid, 256, 0, st>>>(a, 1);      // the planes flagged as wider than 64 cells, up to 256
  XS_CK(cudaGetLastError());
  if (mem == XS3D_HOST) {
    XS_CK(cudaMemcpyAsync(shapes, d_shapes, n * 16, cudaMemcpyDeviceToHost, st));
    XS_CK(cudaMemcpyAsync(overflow, d_over, n, cudaMemcpyDeviceToHost, st));
    if (out_bytes) XS_CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    XS_CK(cudaStreamSynchronize(st));
    // planes whose window does not fit the one-CTA kernel go through the whole-GPU path one by one
    for (uint64_t q = 0; q < n; q++) {
      if (overflow[q] != 2) continue;
      int64_t shp[2]; void* d_cut;
      XS_RET(slice_big(v, pos + 3 * q, normal + 3 * q, anisotropy, standardize_basis, crop_distance, shp, &d_cut));
      shapes[2 * q] = shp[0]; shapes[2 * q + 1] = shp[1];
      const uint64_t total = (uint64_t)shp[0] * shp[1];
      overflow[q] = total > pitch_elems ? 1 : 0;
      if (total <= pitch_elems) {
        XS_CK(cudaMemcpyAsync((char*)out + (size_t)q * pitch_elems * itemsize, d_cut, total * itemsize, cudaMemcpyDeviceToHost, st));
        XS_CK(cudaStreamSynchronize(st));
      }
    }
  } else {
    XS_CK(cudaStreamSynchronize(st));
  }
  return XS3D_OK;
}
