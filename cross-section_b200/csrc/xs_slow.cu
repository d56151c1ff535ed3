// xs_slow.cu — the reference's whole-volume ("slow") variants on the GPU.
//
// Replaces (file:line in /root/reference/src): cross_section_slow auxiliary.hpp:94-181 (method 1),
// cross_section_slow_2x2x2 auxiliary.hpp:8-90 (method 2), cross_sectional_area_slow auxiliary.hpp:193-266
// (slow_method=True).  All three visit every foreground voxel with no component filtering, so they are
// one HBM-streaming scan of the occupancy bytes (V/8 bytes) plus a polygon evaluation for the voxels
// within sqrt(3)/2 of the plane.  The area variant accumulates in DOUBLE in raster (z,y,x) order; to
// stay bit-identical the non-zero contributions are sorted by voxel index (bitonic sort) and summed
// sequentially in that order (zeros do not change a floating-point sum).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>

#include "xs_device.h"
#include "../../include/xs3d_b200.h"

using namespace xs;

struct SlowArgs {
  VolView vol;
  PlaneGeom g;
  uint64_t* idx; float* val;
  uint32_t cap;
  uint32_t* counters;   // [0] nnz, [1] contact, [2] overflow
  int method;           // 1: per voxel, 2: per block (half resolution)
};

__global__ void __launch_bounds__(256) slow_scan_kernel(SlowArgs a) {
  const VolView& v = a.vol;
  const size_t total = (size_t)v.hx * v.hy * v.hz;
  uint32_t ct = 0u;
  const int lastx = v.sx - 1, lasty = v.sy - 1, lastz = v.sz - 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t cube = __ldcs(v.cubes + i);
    if (!cube) continue;
    const int bx = (int)(i % v.hx);
    const size_t t = i / v.hx;
    const int by = (int)(t % v.hy), bz = (int)(t / v.hy);
    if (a.method == 1) {
      // faces touched by ANY foreground voxel (auxiliary.hpp:161-166)
      if (bx == 0 && (cube & 0x55u)) ct |= 1u;
      if (bx == (lastx >> 1) && (cube & ((lastx & 1) ? 0xaau : 0x55u))) ct |= 2u;
      if (by == 0 && (cube & 0x33u)) ct |= 4u;
      if (by == (lasty >> 1) && (cube & ((lasty & 1) ? 0xccu : 0x33u))) ct |= 8u;
      if (bz == 0 && (cube & 0x0fu)) ct |= 16u;
      if (bz == (lastz >> 1) && (cube & ((lastz & 1) ? 0xf0u : 0x0fu))) ct |= 32u;
    }
    // whole-block early out: every voxel centre is within sqrt(3)/2 of the block centre
    const float cx = 2.0f * bx + 0.5f, cy = 2.0f * by + 0.5f, cz = 2.0f * bz + 0.5f;
    const float d = fabsf((cx - a.g.pos.x) * a.g.n.x + (cy - a.g.pos.y) * a.g.n.y + (cz - a.g.pos.z) * a.g.n.z);
    if (a.method == 1) {
      if (d > 1.8f) continue;
      for (int b = 0; b < 8; b++) {
        if (!((cube >> b) & 1u)) continue;
        const int x = 2 * bx + (b & 1), y = 2 * by + ((b >> 1) & 1), z = 2 * bz + (b >> 2);
        const float area = voxel_area((float)x, (float)y, (float)z, a.g);
        if (!(area == 0.0f)) {
          const uint32_t slot = atomicAdd(&a.counters[0], 1u);
          if (slot < a.cap) { a.idx[slot] = (uint64_t)x + (uint64_t)v.sx * ((uint64_t)y + (uint64_t)v.sy * (uint64_t)z); a.val[slot] = area; }
          else a.counters[2] = 1u;
        }
      }
    } else {
      if (!(cube & 1u) || d > 1.8f) continue;   // auxiliary.hpp:72: only blocks whose origin voxel is set
      const float area = block_polygon_area((float)(2 * bx), (float)(2 * by), (float)(2 * bz), a.g);
      if (!(area == 0.0f)) {
        const uint32_t slot = atomicAdd(&a.counters[0], 1u);
        if (slot < a.cap) { a.idx[slot] = (uint64_t)i; a.val[slot] = area; }
        else a.counters[2] = 1u;
      }
    }
  }
  ct = __reduce_or_sync(0xffffffffu, ct);
  if ((threadIdx.x & 31) == 0 && ct) atomicOr(&a.counters[1], ct);
}

// ---- bitonic sort of (idx, val) pairs by idx, padded to a power of two with idx = ~0 ---------------
__global__ void __launch_bounds__(256) sort_pad_kernel(uint64_t* idx, float* val, uint32_t n, uint32_t np2) {
  for (uint32_t i = n + blockIdx.x * blockDim.x + threadIdx.x; i < np2; i += gridDim.x * blockDim.x) { idx[i] = ~0ull; val[i] = 0.0f; }
}
__global__ void __launch_bounds__(256) bitonic_step_kernel(uint64_t* idx, float* val, uint32_t np2, uint32_t j, uint32_t k) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < np2; i += gridDim.x * blockDim.x) {
    const uint32_t l = i ^ j;
    if (l <= i) continue;
    const uint64_t a = idx[i], b = idx[l];
    const bool up = (i & k) == 0;
    if ((a > b) == up) {
      idx[i] = b; idx[l] = a;
      const float t = val[i]; val[i] = val[l]; val[l] = t;
    }
  }
}
// sequential double accumulation in index order (auxiliary.hpp:220,260), staged through shared memory
__global__ void __launch_bounds__(256) ordered_double_sum_kernel(const float* __restrict__ val, uint32_t n, float* __restrict__ out) {
  __shared__ float buf[4096];
  double acc = 0.0;
  for (uint32_t b = 0; b < n; b += 4096) {
    const uint32_t len = min(4096u, n - b);
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) buf[i] = val[b + i];
    __syncthreads();
    if (threadIdx.x == 0)
      for (uint32_t i = 0; i < len; i++) acc += (double)buf[i];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)acc;
}

static bool slow_seed_ok(const VolView& v, const float p[3], bool check_pos) {
  const float s[3] = { (float)v.sx, (float)v.sy, (float)v.sz };
  for (int a = 0; a < 3; a++) {
    if (check_pos && (p[a] < 0 || p[a] >= s[a])) return false;      // auxiliary.hpp:25-33 / 111-119
    const float r = roundf(p[a]);
    if (r < 0 || r >= s[a]) return false;                            // auxiliary.hpp:37-43 / 124-130 / 206-212
  }
  return true;
}

static int run_scan(xs3d_volume* v, const float p[3], const float n[3], const float w[3], int method,
                    uint64_t** d_idx, float** d_val, uint32_t* nnz, uint8_t* contact, uint32_t* cap_out) {
  const VolView vol = xs_view(v);
  cudaStream_t st = xs_stream(v);
  const unsigned long long nvox = (unsigned long long)vol.sx * vol.sy * vol.sz;
  const unsigned long long m = std::max(vol.sx, std::max(vol.sy, vol.sz));
  unsigned long long cap = std::min<unsigned long long>(nvox, 6ull * (unsigned long long)(1.5 * (double)m * (double)m + 16.0 * m + 4096.0));
  unsigned long long np2 = 1024;
  while (np2 < cap) np2 <<= 1;
  if (np2 > 0x7fffffffull) return xs_fail(XS3D_ERR_UNSUPPORTED, "slow variants: volume too large");
  void *bi, *bv, *bc;
  XS_RET(xs_buf(v, 0, np2 * 8 + 64, &bi));
  XS_RET(xs_buf(v, 1, np2 * 4 + 64, &bv));
  XS_RET(xs_buf(v, 3, 256, &bc));
  uint32_t* d_cnt = (uint32_t*)bc + 40;
  XS_CK(cudaMemsetAsync(d_cnt, 0, 16, st));
  SlowArgs a;
  a.vol = vol;
  const V3 nn = vdivs(mk(n[0], n[1], n[2]), vnorm(mk(n[0], n[1], n[2])));   // host float math: IEEE, same as device
  geom_init(a.g, mk(p[0], p[1], p[2]), nn, mk(w[0], w[1], w[2]));
  a.idx = (uint64_t*)bi; a.val = (float*)bv; a.cap = (uint32_t)cap; a.counters = d_cnt; a.method = method;
  const size_t total = (size_t)vol.hx * vol.hy * vol.hz;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)xs_sm_count(v) * 16);
  slow_scan_kernel<<<grid, 256, 0, st>>>(a);
  XS_CK(cudaGetLastError());
  uint32_t* h = xs_pinned(v);
  XS_CK(cudaMemcpyAsync(h, d_cnt, 16, cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  if (h[2]) return xs_fail(XS3D_ERR_CAPACITY, "slow variant exceeded its output capacity");
  *nnz = h[0]; *contact = (uint8_t)h[1];
  *d_idx = a.idx; *d_val = a.val; *cap_out = (uint32_t)np2;
  return XS3D_OK;
}

int xs_section_slow_sparse(xs3d_volume* v, const float p[3], const float n[3], const float w[3], int method,
                           uint64_t** d_idx, float** d_val, uint64_t* nnz, uint8_t* contact) {
  *nnz = 0; *contact = 0;
  const VolView vol = xs_view(v);
  if (!slow_seed_ok(vol, p, true)) return XS3D_OK;
  uint32_t k = 0, cap = 0;
  uint8_t ct = 0;
  XS_RET(run_scan(v, p, n, w, method >= 2 ? 2 : 1, d_idx, d_val, &k, &ct, &cap));
  *nnz = k;
  *contact = (method >= 2) ? 0 : ct;     // auxiliary.hpp:62 : method 2 reports no contact
  return XS3D_OK;
}

int xs_area_slow(xs3d_volume* v, const float p[3], const float n[3], const float w[3], float* area, uint8_t* contact) {
  *area = 0.0f; *contact = 0;
  const VolView vol = xs_view(v);
  if (!slow_seed_ok(vol, p, false)) return XS3D_OK;
  uint64_t* d_idx; float* d_val;
  uint32_t k = 0, cap = 0;
  uint8_t ct = 0;
  XS_RET(run_scan(v, p, n, w, 1, &d_idx, &d_val, &k, &ct, &cap));
  *contact = ct;
  if (k == 0) return XS3D_OK;
  cudaStream_t st = xs_stream(v);
  uint32_t np2 = 1;
  while (np2 < k) np2 <<= 1;
  const int grid = (int)std::min<uint32_t>((np2 + 255) / 256, (uint32_t)xs_sm_count(v) * 8);
  sort_pad_kernel<<<grid, 256, 0, st>>>(d_idx, d_val, k, np2);
  for (uint32_t kk = 2; kk <= np2; kk <<= 1)
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) bitonic_step_kernel<<<grid, 256, 0, st>>>(d_idx, d_val, np2, j, kk);
  void* bc;
  XS_RET(xs_buf(v, 3, 256, &bc));
  float* d_out = (float*)bc + 48;
  ordered_double_sum_kernel<<<1, 256, 0, st>>>(d_val, k, d_out);
  XS_CK(cudaGetLastError());
  float* h = (float*)xs_pinned(v);
  XS_CK(cudaMemcpyAsync(h, d_out, 4, cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  *area = *h;
  return XS3D_OK;
}
