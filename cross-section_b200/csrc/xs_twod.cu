// xs_twod.cu — the 2-D form of cross_sectional_area (the reference's pure-Python xs3d/twod.py:73-174, reached from
// xs3d/__init__.py:79-80 for 2-D images) on the device.
//
// What the reference computes (oracle/twod_oracle.py restates it with every numpy dtype decision spelled out; numpy >= 2
// scalar promotion: everything derived from the float32 pos / vec stays float32 until the points go through np.unique):
//   mark      : every foreground pixel whose nearest point on the line (slope -vec[0]/vec[1] through pos) lies inside the
//               pixel marks the pixel that point rounds to (half to even), with that pixel's own image value  (:25-71)
//   component : 8-connected component of the seed pixel (int(pos)) among the marks — the reference labels with cc3d and
//               compares labels; an unmarked seed has the background label 0 and selects ALL unmarked pixels   (:71, 93, 112)
//   lengths   : per selected pixel the intersections of the line with slope nhat[1]/nhat[0] with the pixel's sides
//               (float32), exact duplicates dropped, two distinct points -> anisotropy-scaled distance (float64);
//               more than two -> ValueError.  Sum in float64 in raster order; contact bits 0-3.              (:96-172)
// Kernels: twod_mark (one thread per pixel) -> twod_init / twod_link (lock-free union-find over the marks, 8-connected)
// -> twod_lengths (one thread per pixel) -> twod_sum (one CTA: the non-zero lengths are compacted in raster order and
// added up by one thread in that order — adding the zeros the reference also adds changes nothing, so the float64 sum
// is the reference's bit for bit).  float32 arithmetic through __f*_rn (no contraction), doubles through __d*_rn.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/xs3d_b200.h"
#include "xs_device.h"
#include "xs_math.cuh"

using namespace xs;

namespace {

struct TwoD {
  int sx, sy;
  float px, py, vx, vy;      // pos, vec
  double wx, wy;             // anisotropy (float32 values widened, as float(anisotropy[i]) does)
};

__device__ __forceinline__ bool f_isinf_pos(float a) { return a == INFINITY; }

// twod.py:25-71
__global__ void twod_mark(const uint8_t* __restrict__ img, uint8_t* __restrict__ marks, TwoD t) {
  const size_t n = (size_t)t.sx * t.sy;
  const float slope = (t.vy != 0.0f) ? fdiv(-t.vx, t.vy) : INFINITY;
  const float b = fsub(t.py, fmul(slope, t.px));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!img[i]) continue;
    const int x = (int)(i % (size_t)t.sx), y = (int)(i / (size_t)t.sx);
    float p_x, p_y;
    if (slope == 0.0f) { p_x = (float)x; p_y = t.py; }
    else if (f_isinf_pos(slope)) { p_x = t.px; p_y = (float)y; }
    else {
      const float m_n = fdiv(-1.0f, slope);
      const float b_n = fsub((float)y, fmul(m_n, (float)x));
      p_x = fdiv(fsub(b, b_n), fsub(m_n, slope));
      p_y = fadd(fmul(slope, p_x), b);
    }
    if (p_x < (float)x - 0.5f || p_x > (float)x + 0.5f || p_y < (float)y - 0.5f || p_y > (float)y + 0.5f) continue;   // (x +- 0.5 is exact)
    if (!(p_x == p_x) || !(p_y == p_y)) continue;       // NaN: the reference fails on int(round(nan)); nothing is marked here
    int tx = (int)rintf(p_x), ty = (int)rintf(p_y);
    if (tx > t.sx - 1) tx = t.sx - 1;
    if (ty > t.sy - 1) ty = t.sy - 1;
    if (tx < 0 || ty < 0) continue;
    const size_t j = (size_t)tx + (size_t)t.sx * ty;
    if (img[j]) marks[j] = 1;
  }
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t i) {
  for (;;) {
    const uint32_t p = *reinterpret_cast<volatile uint32_t*>(parent + i);
    if (p == i) return i;
    const uint32_t g = *reinterpret_cast<volatile uint32_t*>(parent + p);
    if (g != p) parent[i] = g;       // path halving (a stale write only points at another ancestor)
    i = p;
  }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t s = a; a = b; b = s; }
    if (atomicCAS(parent + a, a, b) == a) return;       // the larger root hangs under the smaller
  }
}
__global__ void twod_init(uint32_t* parent, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) parent[i] = (uint32_t)i;
}
__global__ void twod_link(const uint8_t* __restrict__ marks, uint32_t* parent, TwoD t) {
  const size_t n = (size_t)t.sx * t.sy;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!marks[i]) continue;
    const int x = (int)(i % (size_t)t.sx), y = (int)(i / (size_t)t.sx);
    // the four "forward" neighbours of the 8-neighbourhood
    if (x + 1 < t.sx && marks[i + 1]) uf_union(parent, (uint32_t)i, (uint32_t)(i + 1));
    if (y + 1 < t.sy) {
      const size_t r = i + (size_t)t.sx;
      if (marks[r]) uf_union(parent, (uint32_t)i, (uint32_t)r);
      if (x > 0 && marks[r - 1]) uf_union(parent, (uint32_t)i, (uint32_t)(r - 1));
      if (x + 1 < t.sx && marks[r + 1]) uf_union(parent, (uint32_t)i, (uint32_t)(r + 1));
    }
  }
}

// twod.py:96-172 for one pixel
__global__ void twod_lengths(const uint8_t* __restrict__ marks, uint32_t* parent, double* __restrict__ vals, uint32_t* __restrict__ flags, TwoD t) {
  const size_t n = (size_t)t.sx * t.sy;
  const size_t seed = (size_t)(int)t.px + (size_t)t.sx * (size_t)(int)t.py;
  const bool seed_marked = marks[seed] != 0;
  const uint32_t seed_root = seed_marked ? uf_find(parent, (uint32_t)seed) : 0u;
  float n0 = -t.vy, n1 = t.vx;
  const float norm = fsqrt(fadd(fmul(n0, n0), fmul(n1, n1)));
  n0 = fdiv(n0, norm); n1 = fdiv(n1, norm);
  const float slope = (n0 == 0.0f) ? INFINITY : fdiv(n1, n0);
  const float yi = fsub(t.py, fmul(slope, t.px));
  uint32_t contact = 0u, err = 0u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double val = 0.0;
    const bool sel = seed_marked ? (marks[i] && uf_find(parent, (uint32_t)i) == seed_root) : !marks[i];
    if (sel) {
      const int x = (int)(i % (size_t)t.sx), y = (int)(i / (size_t)t.sx);
      contact |= (uint32_t)(x == 0) | ((uint32_t)(x == t.sx - 1) << 1) | ((uint32_t)(y == 0) << 2) | ((uint32_t)(y == t.sy - 1) << 3);
      const float h0 = (float)y - 0.5f, h1 = (float)y + 0.5f, v0 = (float)x - 0.5f, v1 = (float)x + 0.5f;
      double qx[4], qy[4];
      int np = 0;
      if (slope == 0.0f) {
        if ((float)y == t.py) { qx[0] = v0; qy[0] = y; qx[1] = v1; qy[1] = y; np = 2; }
      } else if (slope == INFINITY || slope == -INFINITY) {
        if ((float)x == t.px) { qx[0] = x; qy[0] = h0; qx[1] = x; qy[1] = h1; np = 2; }
      } else {
        const float x0 = fdiv(fsub(h0, yi), slope), x1 = fdiv(fsub(h1, yi), slope);
        const float y0 = fadd(fmul(v0, slope), yi), y1 = fadd(fmul(v1, slope), yi);
        if (y0 >= h0 && y0 <= h1) { qx[np] = v0; qy[np] = y0; np++; }
        if (y1 >= h0 && y1 <= h1) { qx[np] = v1; qy[np] = y1; np++; }
        if (x0 >= v0 && x0 <= v1) { qx[np] = x0; qy[np] = h0; np++; }
        if (x1 >= v0 && x1 <= v1) { qx[np] = x1; qy[np] = h1; np++; }
      }
      // np.unique(pts, axis=0): distinct rows (== on doubles: -0.0 and 0.0 are one value)
      double ux[4], uy[4];
      int nu = 0;
      for (int a = 0; a < np; a++) {
        bool dup = false;
        for (int c = 0; c < nu; c++) dup = dup || (ux[c] == qx[a] && uy[c] == qy[a]);
        if (!dup) { ux[nu] = qx[a]; uy[nu] = qy[a]; nu++; }
      }
      if (nu > 2) err = 1u;
      else if (nu == 2) {
        const double dx = __dmul_rn(t.wx, __dsub_rn(ux[0], ux[1])), dy = __dmul_rn(t.wy, __dsub_rn(uy[0], uy[1]));
        val = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      }
    }
    vals[i] = val;
  }
  contact = __reduce_or_sync(0xffffffffu, contact);
  err = __reduce_or_sync(0xffffffffu, err);
  if ((threadIdx.x & 31) == 0) {
    if (contact) atomicOr(&flags[0], contact);
    if (err) atomicOr(&flags[1], err);
  }
}

// one CTA: the non-zero lengths in raster order, added up in that order by one thread
__global__ void __launch_bounds__(1024) twod_sum(const double* __restrict__ vals, size_t n, double* __restrict__ list, double* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (size_t b = 0; b < n; b += 1024) {
    const size_t i = b + threadIdx.x;
    const double v = i < n ? vals[i] : 0.0;
    const int has = v != 0.0 ? 1 : 0;
    int inc = has;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    if (lane == 31) wsum[wrp] = inc;
    __syncthreads();
    int off = inc - has, tot = 0;
    for (int w = 0; w < 32; w++) { if (w < wrp) off += wsum[w]; tot += wsum[w]; }
    if (has) list[base_s + off] = v;
    __syncthreads();
    if (threadIdx.x == 0) base_s += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int k = 0; k < base_s; k++) total = __dadd_rn(total, list[k]);
    *out = total;
  }
}

int twod_fail(int code, const char* msg) { return xs_fail(code, msg); }

}  // namespace

#define CK2(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = twod_fail(XS3D_ERR_CUDA, cudaGetErrorString(e_)); goto done; } } while (0)

extern "C" int xs3d_area_2d(const uint8_t* binimg, uint64_t sx, uint64_t sy, const float pos[2], const float vec[2],
                            const float anisotropy[2], double* area, uint8_t* contact) {
  if (!pos || !vec || !anisotropy || !area || !contact) return twod_fail(XS3D_ERR_ARG, "null argument");
  if (vec[0] == 0.0f && vec[1] == 0.0f) return twod_fail(XS3D_ERR_ARG, "normal vector must not be a null vector (all zeros).");
  if (sx > 0x7fffffffull || sy > 0x7fffffffull || sx * sy >= 0xffffffffull) return twod_fail(XS3D_ERR_UNSUPPORTED, "2-D image too large");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { cudaGetLastError(); return twod_fail(XS3D_ERR_CUDA, "no CUDA device (there is no CPU fallback)"); }
  // twod.py:83-90: vertex outside the image or on background -> (0.0, 0b00111111)
  *area = 0.0; *contact = 0x3f;
  if (!(pos[0] < (float)sx) || pos[0] < 0.0f || !(pos[1] < (float)sy) || pos[1] < 0.0f) return XS3D_OK;
  if (!binimg) return twod_fail(XS3D_ERR_ARG, "null image");
  const size_t n = (size_t)sx * sy;
  const size_t seed = (size_t)(int)pos[0] + (size_t)sx * (size_t)(int)pos[1];
  if (!binimg[seed]) return XS3D_OK;
  int rc = XS3D_OK;
  uint8_t* d_img = nullptr; uint8_t* d_marks = nullptr; uint32_t* d_parent = nullptr; double* d_vals = nullptr; double* d_list = nullptr;
  uint32_t* d_flags = nullptr; double* d_out = nullptr;
  uint32_t h_flags[2] = { 0u, 0u };
  double h_out = 0.0;
  TwoD t;
  t.sx = (int)sx; t.sy = (int)sy; t.px = pos[0]; t.py = pos[1]; t.vx = vec[0]; t.vy = vec[1];
  t.wx = (double)anisotropy[0]; t.wy = (double)anisotropy[1];
  const int grid = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  CK2(cudaMalloc((void**)&d_img, n)); CK2(cudaMalloc((void**)&d_marks, n)); CK2(cudaMalloc((void**)&d_parent, n * 4));
  CK2(cudaMalloc((void**)&d_vals, n * 8)); CK2(cudaMalloc((void**)&d_list, n * 8)); CK2(cudaMalloc((void**)&d_flags, 8)); CK2(cudaMalloc((void**)&d_out, 8));
  CK2(cudaMemcpy(d_img, binimg, n, cudaMemcpyHostToDevice));
  CK2(cudaMemset(d_marks, 0, n)); CK2(cudaMemset(d_flags, 0, 8));
  twod_mark<<<grid, 256>>>(d_img, d_marks, t);
  twod_init<<<grid, 256>>>(d_parent, n);
  twod_link<<<grid, 256>>>(d_marks, d_parent, t);
  twod_lengths<<<grid, 256>>>(d_marks, d_parent, d_vals, d_flags, t);
  twod_sum<<<1, 1024>>>(d_vals, n, d_list, d_out);
  CK2(cudaGetLastError());
  CK2(cudaMemcpy(h_flags, d_flags, 8, cudaMemcpyDeviceToHost));
  CK2(cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost));
  if (h_flags[1]) { rc = twod_fail(XS3D_ERR_ARG, "more than two distinct intersection points in a pixel (the reference raises ValueError here, twod.py:163-164)"); goto done; }
  *area = h_out; *contact = (uint8_t)h_flags[0];
done:
  cudaFree(d_img); cudaFree(d_marks); cudaFree(d_parent); cudaFree(d_vals); cudaFree(d_list); cudaFree(d_flags); cudaFree(d_out);
  return rc;
}
