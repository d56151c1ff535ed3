// xs_slice.cu — xs3d.slice on the GPU: dense validity bitmap -> per-row runs -> 1-D row scan -> gather.
// See xs_slice.cuh for the algorithm and the reference lines it replaces
// (xs3d.hpp:1473-1624, fastxs3d.cpp:119-216).
//
// Two execution shapes:
//   * batched path (xs3d_projection_plan / _fill / _batch): n planes of any size at once, no bitmap at all — a warp
//     per window row for the shapes, a warp per 512 output elements for the gather (BASELINE configs 4 and 5);
//   * big path (slice_big): whole-GPU kernels per stage with a dense bitmap — valid -> rows -> scan -> gather — for
//     the single-call drop-in and for the planes that need the exact iterative fill.
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
// batched path: PLAN (shapes of n planes) + FILL (ragged or fixed-pitch gather)
//
// Neither step stores a validity bitmap.  PLAN: one thread per plane sets up its SliceCtx; one WARP per lattice row of
// every plane's window evaluates the row's cells (ballots) into a RowRuns summary — first / last valid column and the
// number of runs —; one warp per plane then finds the rows the reference's 4-connected fill reaches: the maximal range
// of consecutive rows around the centre row whose runs overlap column-wise (every link between two consecutive rows is
// an independent test, so the lanes test 32 links at a time).  When every reached row is a single run, the reached
// cells are exactly the valid cells of those rows, so FILL needs no mask either: one warp per 512 output elements
// evaluates slice_cell for its cells and gathers the labels.  A reachable row with more than one run (float noise on a
// row that grazes a volume face) flags the plane for the exact iterative fill (slice_big), one plane at a time.
// Work scales with the planes' windows: a 7 x 7 cropped slice costs a dozen row-warps, an un-cropped plane through a
// 1024^3 volume ~1800 row-warps of ~56 ballots and a gather of the bounding box.
// ------------------------------------------------------------------------------------------------
struct PlanRec {
  int x0, x1, y0, y1;   // bounding box of the reached cells, window coordinates (x0 > x1: nothing reached)
  int flags;            // PLAN_*
};
enum { PLAN_FALLBACK = 1, PLAN_INACTIVE = 2, PLAN_NOTHING = 4 };

struct SlicePlanArgs {
  const float* pos; const float* nrm;
  float wx, wy, wz;
  int positive; float crop;
  int sx, sy, sz;
  uint32_t n;
};

__global__ void __launch_bounds__(256) slice_plan_init_kernel(SlicePlanArgs a, SliceCtx* __restrict__ ctx, uint32_t* __restrict__ nrows) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= a.n) return;
  SliceCtx s;
  slice_init(s, a.sx, a.sy, a.sz, mk(a.pos[3 * q], a.pos[3 * q + 1], a.pos[3 * q + 2]), mk(a.nrm[3 * q], a.nrm[3 * q + 1], a.nrm[3 * q + 2]),
             mk(a.wx, a.wy, a.wz), a.positive != 0, a.crop);
  ctx[q] = s;
  nrows[q] = s.active ? (uint32_t)s.wh : 0u;
}

// exclusive prefix sums of `in` (n values) into `out` (n + 1 values), one CTA
__global__ void __launch_bounds__(1024) scan_u32_kernel(const uint32_t* __restrict__ in, uint32_t n, unsigned long long* __restrict__ out) {
  __shared__ unsigned long long part[1024];
  const uint32_t per = (n + 1023u) / 1024u;
  const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  unsigned long long s = 0;
  for (uint32_t i = lo; i < hi; i++) s += in[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; i++) { const unsigned long long t = part[i]; part[i] = run; run += t; }
    out[n] = run;
  }
  __syncthreads();
  s = part[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) { out[i] = s; s += in[i]; }
}

// largest i in [0, n) with off[i] <= x  (off ascending, off[0] = 0 <= x < off[n])
__device__ __forceinline__ uint32_t find_owner(const unsigned long long* __restrict__ off, uint32_t n, unsigned long long x) {
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// one warp per window row of every plane: the row's RowRuns without storing its bitmap
__global__ void __launch_bounds__(256) slice_plan_rows_kernel(const SliceCtx* __restrict__ ctx, const unsigned long long* __restrict__ row_off, uint32_t n,
                                                              RowRuns* __restrict__ rows) {
  const int lane = threadIdx.x & 31;
  const unsigned long long total = row_off[n];
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long r = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < total; r += nwarps) {
    const uint32_t q = find_owner(row_off, n, r);
    const SliceCtx* sp = ctx + q;
    SliceCtx s = *sp;
    const int row = (int)(r - row_off[q]);
    int a = 0x7fffffff, b = -1, runs = 0;
    uint32_t carry = 0u;
    for (int x0 = 0; x0 < s.ww; x0 += 32) {
      const int col = x0 + lane;
      const bool ok = (col < s.ww) && slice_cell(s, s.wx0 + col, s.wy0 + row, false, nullptr);
      const uint32_t w = __ballot_sync(0xffffffffu, ok);
      if (w) {
        runs += __popc(w & ~((w << 1) | carry));
        if (a == 0x7fffffff) a = x0 + __ffs(w) - 1;
        b = x0 + 31 - __clz(w);
      }
      carry = w >> 31;
    }
    if (lane == 0) { RowRuns rr; rr.a = a; rr.b = b; rr.runs = runs; rows[r] = rr; }
  }
}

// one warp per plane: the reached row range and its column bounding box (scan_rows, 32 rows at a time)
__global__ void __launch_bounds__(256) slice_plan_scan_kernel(const SliceCtx* __restrict__ ctx, const unsigned long long* __restrict__ row_off, uint32_t n,
                                                              const RowRuns* __restrict__ rows_all, PlanRec* __restrict__ plan) {
  const int lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n; q += nwarps) {
    const SliceCtx* s = ctx + q;
    PlanRec pr;
    pr.x0 = 1; pr.x1 = 0; pr.y0 = 1; pr.y1 = 0; pr.flags = 0;
    if (!s->active) {
      pr.flags = PLAN_INACTIVE;
      if (lane == 0) plan[q] = pr;
      continue;
    }
    const RowRuns* rows = rows_all + row_off[q];
    const int nrows = s->wh;
    const int cx = (int)(s->c - s->wx0), cy = (int)(s->c - s->wy0);
    bool dead = cy < 0 || cy >= nrows;
    RowRuns c; c.a = 0; c.b = -1; c.runs = 0;
    if (!dead) c = rows[cy];
    if (dead || c.runs == 0 || cx < c.a || cx > c.b) {            // the centre cell itself is not valid
      pr.flags = PLAN_NOTHING;
      if (lane == 0) plan[q] = pr;
      continue;
    }
    bool fallback = c.runs > 1;
    int lo = c.a, hi = c.b, y0 = cy, y1 = cy;
    for (int dir = -1; dir <= 1 && !fallback; dir += 2) {
      // link k (k = 1, 2, ...): row cy + dir*k is non-empty and overlaps row cy + dir*(k-1); the fill advances while links hold
      int k0 = 1;
      for (;;) {
        const int k = k0 + lane;
        const int y = cy + dir * k;
        bool link = false, multi = false;
        int qa = 0x7fffffff, qb = -1;
        if (y >= 0 && y < nrows) {
          const RowRuns q1 = rows[y], p1 = rows[y - dir];
          link = q1.runs != 0 && !(q1.a > p1.b || q1.b < p1.a);
          multi = q1.runs > 1;
          qa = q1.a; qb = q1.b;
        }
        const uint32_t broken = __ballot_sync(0xffffffffu, !link);
        const int good = broken ? (__ffs(broken) - 1) : 32;        // links k0 .. k0 + good - 1 hold
        const bool mine = lane < good;
        if (__ballot_sync(0xffffffffu, mine && multi)) { fallback = true; break; }
        const int ma = __reduce_min_sync(0xffffffffu, mine ? qa : 0x7fffffff);
        const int mb = __reduce_max_sync(0xffffffffu, mine ? qb : -1);
        if (good > 0) {
          lo = min(lo, ma); hi = max(hi, mb);
          if (dir < 0) y0 = cy - (k0 + good - 1); else y1 = cy + (k0 + good - 1);
        }
        if (good < 32) break;
        k0 += 32;
      }
    }
    pr.x0 = lo; pr.x1 = hi; pr.y0 = y0; pr.y1 = y1;
    pr.flags = fallback ? PLAN_FALLBACK : 0;
    if (lane == 0) plan[q] = pr;
  }
}

// FILL: one warp per unit of FILL_UNIT output elements of one plane.  unit_off[i] = first unit of plane first + i;
// plane i's cut-out (F order [nx][ny]) goes to out + dst_off[i] elements.  Planes flagged skip (exact-fill fallback,
// or larger than the caller's pitch) are left alone.
constexpr int FILL_UNIT = 512;
__global__ void __launch_bounds__(256) slice_fill_kernel(const SliceCtx* __restrict__ ctx, const PlanRec* __restrict__ plan, const unsigned long long* __restrict__ unit_off,
                                                         const unsigned long long* __restrict__ dst_off, const uint8_t* __restrict__ skip, uint32_t count,
                                                         const void* __restrict__ labels, int itemsize, int c_order, void* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const unsigned long long total = unit_off[count];
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long u = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < total; u += nwarps) {
    const uint32_t i = find_owner(unit_off, count, u);
    if (skip[i]) continue;
    const PlanRec pr = plan[i];
    const unsigned long long base = __ldg(dst_off + i);
    if (pr.flags & (PLAN_INACTIVE | PLAN_NOTHING)) {              // 1 x 1 cut-out holding 0 (fastxs3d.cpp:179-183)
      if (lane == 0) zero_label(out, base, itemsize);
      continue;
    }
    const SliceCtx s = ctx[i];
    const int nx = pr.x1 - pr.x0 + 1, ny = pr.y1 - pr.y0 + 1;
    const unsigned long long cells = (unsigned long long)nx * ny;
    const unsigned long long e0 = (u - unit_off[i]) * FILL_UNIT;
#pragma unroll 4
    for (int j = 0; j < FILL_UNIT / 32; j++) {
      const unsigned long long e = e0 + (unsigned long long)(j * 32 + lane);
      if (e >= cells) break;
      int yy, xx;
      if (cells <= 0xffffffffull) { yy = (int)((uint32_t)e / (uint32_t)nx); xx = (int)((uint32_t)e - (uint32_t)yy * (uint32_t)nx); }
      else { yy = (int)(e / (unsigned long long)nx); xx = (int)(e - (unsigned long long)yy * nx); }
      size_t loc;
      if (slice_cell(s, s.wx0 + pr.x0 + xx, s.wy0 + pr.y0 + yy, c_order != 0, &loc)) copy_label(out, base + e, labels, loc, itemsize);
      else zero_label(out, base + e, itemsize);
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

// The device buffer that holds the label volume, (re)allocated for `itemsize` / `c_order` without uploading anything —
// for a host layer that fills it itself (NCCL broadcast from the rank that uploaded the labels, xs3d.dist.broadcast_labels).
// The labels count as loaded afterwards: the caller writes the buffer before the next slice / select_label call.
extern "C" int xs3d_volume_labels_buffer(xs3d_volume* v, int itemsize, int c_order, void** device_ptr, uint64_t* bytes) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8) return xs_fail(XS3D_ERR_ARG, "label itemsize must be 1, 2, 4 or 8");
  XS_RET(xs_set_device(v));
  XS_RET(xs_labels_set(v, nullptr, itemsize, XS3D_DEVICE, c_order));
  void* p; int is, co; bool loaded;
  xs_labels(v, &p, &is, &co, &loaded);
  uint64_t dims[3];
  XS_RET(xs3d_volume_shape(v, dims));
  *device_ptr = p;
  *bytes = dims[0] * dims[1] * dims[2] * (uint64_t)itemsize;
  return XS3D_OK;
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

// ---- batched planes: plan + fill ---------------------------------------------------------------------
struct SliceState {        // what xs3d_projection_plan leaves behind for xs3d_projection_fill (host side; device side: buffers 4-6)
  uint64_t n = 0;
  bool empty = false;      // crop == 0: every cut-out is 0 x 0
  std::vector<PlanRec> plan;
  std::vector<int64_t> shapes;       // 2 per plane
  std::vector<uint64_t> offsets;     // n + 1, elements
  std::vector<float> pos, nrm;       // host copies (the exact-fill fallback runs plane by plane)
  float w[3] = {1, 1, 1}, crop = 0;
  int positive = 0;
};
static void slice_state_free(void* p) { delete (SliceState*)p; }

static SliceState* slice_state(xs3d_volume* v) {
  void** slot = xs_slice_state_slot(v, slice_state_free);
  if (!*slot) *slot = new SliceState();
  return (SliceState*)*slot;
}

// device layout of buffer 4 after a plan of n planes
struct PlanBufs { SliceCtx* ctx; uint32_t* nrows; unsigned long long* row_off; PlanRec* plan; float* pos; float* nrm; };
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static int plan_bufs(xs3d_volume* v, uint64_t n, PlanBufs* b) {
  const size_t s_ctx = align256(n * sizeof(SliceCtx)), s_nr = align256(n * 4), s_ro = align256((n + 1) * 8), s_pl = align256(n * sizeof(PlanRec)), s_q = align256(n * 12);
  char* p;
  XS_RET(xs_buf(v, 4, s_ctx + s_nr + s_ro + s_pl + 2 * s_q + 256, (void**)&p));
  b->ctx = (SliceCtx*)p; p += s_ctx;
  b->nrows = (uint32_t*)p; p += s_nr;
  b->row_off = (unsigned long long*)p; p += s_ro;
  b->plan = (PlanRec*)p; p += s_pl;
  b->pos = (float*)p; p += s_q;
  b->nrm = (float*)p;
  return XS3D_OK;
}

extern "C" int xs3d_projection_plan(xs3d_volume* v, uint64_t n, const float* pos, const float* normal, const float anisotropy[3],
                                    int standardize_basis, float crop_distance, int mem, int64_t* shapes, uint64_t* offsets) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  void* labels; int itemsize, c_order; bool loaded;
  xs_labels(v, &labels, &itemsize, &c_order, &loaded);
  if (!loaded) return xs_fail(XS3D_ERR_ARG, "volume has no labels loaded");
  if (n >= 0x7fffffffull) return xs_fail(XS3D_ERR_ARG, "too many planes in one batch");
  if (n && (!pos || !normal || !anisotropy || !shapes || !offsets)) return xs_fail(XS3D_ERR_ARG, "null argument");
  if (!(crop_distance >= 0.0f)) return xs_fail(XS3D_ERR_ARG, "crop distance must be >= 0");
  XS_RET(xs_set_device(v));
  SliceState* st = slice_state(v);
  st->n = n; st->empty = false;
  st->plan.clear(); st->shapes.assign(2 * n, 0); st->offsets.assign(n + 1, 0);
  if (offsets) offsets[0] = 0;
  if (n == 0) return XS3D_OK;
  if (crop_distance == 0.0f) {                                   // fastxs3d.cpp:160-162
    st->empty = true;
    memset(shapes, 0, n * 16); memset(offsets, 0, (n + 1) * 8);
    return XS3D_OK;
  }
  uint64_t dims[3];
  XS_RET(xs3d_volume_shape(v, dims));
  cudaStream_t sm = xs_stream(v);
  PlanBufs b;
  XS_RET(plan_bufs(v, n, &b));
  st->pos.resize(3 * n); st->nrm.resize(3 * n);
  if (mem == XS3D_HOST) {
    memcpy(st->pos.data(), pos, n * 12); memcpy(st->nrm.data(), normal, n * 12);
    XS_CK(cudaMemcpyAsync(b.pos, pos, n * 12, cudaMemcpyHostToDevice, sm));
    XS_CK(cudaMemcpyAsync(b.nrm, normal, n * 12, cudaMemcpyHostToDevice, sm));
  } else {
    XS_CK(cudaMemcpyAsync(b.pos, pos, n * 12, cudaMemcpyDeviceToDevice, sm));
    XS_CK(cudaMemcpyAsync(b.nrm, normal, n * 12, cudaMemcpyDeviceToDevice, sm));
    XS_CK(cudaMemcpyAsync(st->pos.data(), pos, n * 12, cudaMemcpyDeviceToHost, sm));
    XS_CK(cudaMemcpyAsync(st->nrm.data(), normal, n * 12, cudaMemcpyDeviceToHost, sm));
  }
  memcpy(st->w, anisotropy, 12); st->crop = crop_distance; st->positive = standardize_basis;
  SlicePlanArgs a;
  a.pos = b.pos; a.nrm = b.nrm; a.wx = anisotropy[0]; a.wy = anisotropy[1]; a.wz = anisotropy[2];
  a.positive = standardize_basis; a.crop = crop_distance;
  a.sx = (int)dims[0]; a.sy = (int)dims[1]; a.sz = (int)dims[2]; a.n = (uint32_t)n;
  slice_plan_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, sm>>>(a, b.ctx, b.nrows);
  scan_u32_kernel<<<1, 1024, 0, sm>>>(b.nrows, (uint32_t)n, b.row_off);
  XS_CK(cudaGetLastError());
  unsigned long long* h_total = (unsigned long long*)xs_pinned(v);
  XS_CK(cudaMemcpyAsync(h_total, b.row_off + n, 8, cudaMemcpyDeviceToHost, sm));
  XS_CK(cudaStreamSynchronize(sm));
  const unsigned long long total_rows = *h_total;
  RowRuns* rows;
  XS_RET(xs_buf(v, 5, (size_t)total_rows * sizeof(RowRuns) + 256, (void**)&rows));
  const int sms = xs_sm_count(v);
  if (total_rows) {
    const unsigned rgrid = (unsigned)std::min<unsigned long long>((total_rows + 7) / 8, (unsigned long long)sms * 16);
    slice_plan_rows_kernel<<<rgrid, 256, 0, sm>>>(b.ctx, b.row_off, (uint32_t)n, rows);
  }
  const unsigned sgrid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)sms * 16);
  slice_plan_scan_kernel<<<sgrid, 256, 0, sm>>>(b.ctx, b.row_off, (uint32_t)n, rows, b.plan);
  XS_CK(cudaGetLastError());
  st->plan.resize(n);
  XS_CK(cudaMemcpyAsync(st->plan.data(), b.plan, n * sizeof(PlanRec), cudaMemcpyDeviceToHost, sm));
  XS_CK(cudaStreamSynchronize(sm));
  uint64_t off = 0;
  for (uint64_t i = 0; i < n; i++) {
    const PlanRec& pr = st->plan[i];
    int64_t nx = 1, ny = 1;
    if (pr.flags & PLAN_FALLBACK) {
      int64_t shp[2]; void* d_cut;
      XS_RET(slice_big(v, st->pos.data() + 3 * i, st->nrm.data() + 3 * i, anisotropy, standardize_basis, crop_distance, shp, &d_cut));
      nx = shp[0]; ny = shp[1];
    } else if (!(pr.flags & (PLAN_INACTIVE | PLAN_NOTHING))) {
      nx = pr.x1 - pr.x0 + 1; ny = pr.y1 - pr.y0 + 1;
    }
    st->shapes[2 * i] = nx; st->shapes[2 * i + 1] = ny;
    st->offsets[i] = off;
    off += (uint64_t)nx * (uint64_t)ny;
  }
  st->offsets[n] = off;
  memcpy(shapes, st->shapes.data(), n * 16);
  memcpy(offsets, st->offsets.data(), (n + 1) * 8);
  return XS3D_OK;
}

// gather planes [first, first + count) of the last plan; plane i goes to element dst[i - first] of `out`, planes with
// skip[i - first] != 0 are left out.  `out_elems`: size of the destination in elements (for the host staging copy).
static int fill_impl(xs3d_volume* v, uint64_t first, uint64_t count, const std::vector<uint64_t>& dst, const std::vector<uint8_t>& skip_in,
                     uint64_t out_elems, void* out, int mem) {
  SliceState* st = slice_state(v);
  void* labels; int itemsize, c_order; bool loaded;
  xs_labels(v, &labels, &itemsize, &c_order, &loaded);
  if (!loaded) return xs_fail(XS3D_ERR_ARG, "volume has no labels loaded");
  if (count == 0 || out_elems == 0) return XS3D_OK;
  cudaStream_t sm = xs_stream(v);
  PlanBufs b;
  XS_RET(plan_bufs(v, st->n, &b));
  // per-plane meta for the kernel: unit offsets, destination offsets, skip flags
  std::vector<unsigned long long> meta(2 * count + 1);
  std::vector<uint8_t> skip(skip_in);
  unsigned long long units = 0;
  bool any_fallback = false;
  for (uint64_t i = 0; i < count; i++) {
    const PlanRec& pr = st->plan[first + i];
    if (pr.flags & PLAN_FALLBACK) { any_fallback = any_fallback || !skip[i]; skip[i] = 1; }
    meta[i] = units;
    const uint64_t cells = (uint64_t)st->shapes[2 * (first + i)] * (uint64_t)st->shapes[2 * (first + i) + 1];
    units += skip[i] ? 1 : (cells + FILL_UNIT - 1) / FILL_UNIT;
    meta[count + 1 + i] = dst[i];
  }
  meta[count] = units;
  char* dmeta;
  const size_t meta_bytes = align256(meta.size() * 8);
  XS_RET(xs_buf(v, 6, meta_bytes + align256(count) + 256, (void**)&dmeta));
  XS_CK(cudaMemcpyAsync(dmeta, meta.data(), meta.size() * 8, cudaMemcpyHostToDevice, sm));
  XS_CK(cudaMemcpyAsync(dmeta + meta_bytes, skip.data(), count, cudaMemcpyHostToDevice, sm));
  void* d_out = out;
  if (mem == XS3D_HOST) XS_RET(xs_buf(v, 7, (size_t)out_elems * itemsize + 256, &d_out));
  const unsigned grid = (unsigned)std::min<unsigned long long>((units + 7) / 8, (unsigned long long)xs_sm_count(v) * 32);
  slice_fill_kernel<<<grid, 256, 0, sm>>>(b.ctx + first, b.plan + first, (const unsigned long long*)dmeta, (const unsigned long long*)dmeta + count + 1,
                                          (const uint8_t*)(dmeta + meta_bytes), (uint32_t)count, labels, itemsize, c_order, d_out);
  XS_CK(cudaGetLastError());
  if (any_fallback) {
    for (uint64_t i = 0; i < count; i++) {
      if (!(st->plan[first + i].flags & PLAN_FALLBACK) || skip_in[i]) continue;
      int64_t shp[2]; void* d_cut;
      XS_RET(slice_big(v, st->pos.data() + 3 * (first + i), st->nrm.data() + 3 * (first + i), st->w, st->positive, st->crop, shp, &d_cut));
      XS_CK(cudaMemcpyAsync((char*)d_out + dst[i] * itemsize, d_cut, (size_t)shp[0] * shp[1] * itemsize, cudaMemcpyDeviceToDevice, sm));
    }
  }
  if (mem == XS3D_HOST) XS_CK(cudaMemcpyAsync(out, d_out, (size_t)out_elems * itemsize, cudaMemcpyDeviceToHost, sm));
  XS_CK(cudaStreamSynchronize(sm));
  return XS3D_OK;
}

extern "C" int xs3d_projection_fill(xs3d_volume* v, uint64_t first, uint64_t count, void* out, int mem) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  XS_RET(xs_set_device(v));
  SliceState* st = slice_state(v);
  if (first > st->n || count > st->n - first) return xs_fail(XS3D_ERR_ARG, "plane range outside the last plan");
  if (count == 0 || st->empty) return XS3D_OK;
  const uint64_t base = st->offsets[first], total = st->offsets[first + count] - base;
  if (total && !out) return xs_fail(XS3D_ERR_ARG, "null output");
  std::vector<uint64_t> dst(count);
  for (uint64_t i = 0; i < count; i++) dst[i] = st->offsets[first + i] - base;
  return fill_impl(v, first, count, dst, std::vector<uint8_t>(count, 0), total, out, mem);
}

// fixed-pitch form: plan, then every cut-out that fits at out + i * pitch_elems
extern "C" int xs3d_projection_batch(xs3d_volume* v, uint64_t n, const float* pos, const float* normal, const float anisotropy[3],
                                     int standardize_basis, float crop_distance, void* out, uint64_t pitch_elems,
                                     int64_t* shapes, uint8_t* overflow, int mem) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  if (n == 0) return XS3D_OK;
  if (!shapes || !overflow) return xs_fail(XS3D_ERR_ARG, "null argument");
  std::vector<int64_t> h_shapes(2 * n);
  std::vector<uint64_t> h_off(n + 1);
  XS_RET(xs3d_projection_plan(v, n, pos, normal, anisotropy, standardize_basis, crop_distance, mem, h_shapes.data(), h_off.data()));
  std::vector<uint8_t> over(n, 0);
  std::vector<uint64_t> dst(n);
  for (uint64_t i = 0; i < n; i++) {
    over[i] = (uint64_t)h_shapes[2 * i] * (uint64_t)h_shapes[2 * i + 1] > pitch_elems ? 1 : 0;
    dst[i] = i * pitch_elems;
  }
  cudaStream_t sm = xs_stream(v);
  if (mem == XS3D_HOST) {
    memcpy(shapes, h_shapes.data(), n * 16); memcpy(overflow, over.data(), n);
  } else {
    XS_CK(cudaMemcpyAsync(shapes, h_shapes.data(), n * 16, cudaMemcpyHostToDevice, sm));
    XS_CK(cudaMemcpyAsync(overflow, over.data(), n, cudaMemcpyHostToDevice, sm));
    XS_CK(cudaStreamSynchronize(sm));
  }
  if (crop_distance == 0.0f || pitch_elems == 0) return XS3D_OK;
  if (!out) return xs_fail(XS3D_ERR_ARG, "null output");
  return fill_impl(v, 0, n, dst, over, n * pitch_elems, out, mem);
}
