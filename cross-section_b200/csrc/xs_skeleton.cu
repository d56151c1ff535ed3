// xs_skeleton.cu — per-vertex tangents of a skeleton on the device: the section planes of a skeleton
// sweep are normal to the skeleton, so a (vertices, edges) skeleton already resident in HBM goes straight into
// xs3d_area_batch without a host hop.  Reference use case: README.md:59-71 (the reference itself has no sweep driver).
//
// tangent(v) = normalise( sum over incident edges e of  s_e * dir_e ),  dir_e = unit vector from v to its neighbour,
// s_e = -1 when dir_e opposes the direction of v's FIRST incident half-edge (smallest index in the list
// [a->b for every edge] + [b->a for every edge]), so that the two edges of a path vertex reinforce instead of cancel.
// Zero-length edges are ignored; a vertex without edges gets (0, 0, 1).  Same definition as the host version
// (xs3d.skeleton_tangents); the sum is accumulated in 2^-40 fixed point with integer atomics, which makes it exact and
// independent of the order in which the half-edges arrive — the result is deterministic.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>

#include "xs_device.h"
#include "../../include/xs3d_b200.h"

namespace {

struct Skel {
  const float* v; const int64_t* e;
  uint64_t nv, ne;
  unsigned long long* first;     // per vertex: smallest incident half-edge index
  long long* acc;                // per vertex: 3 fixed-point sums
  float* out;
};

__device__ __forceinline__ bool half_edge(const Skel& s, uint64_t h, uint64_t* src, double d[3]) {
  const uint64_t k = h < s.ne ? h : h - s.ne;
  const int64_t a = s.e[2 * k], b = s.e[2 * k + 1];
  if (a < 0 || b < 0 || (uint64_t)a >= s.nv || (uint64_t)b >= s.nv) return false;
  const uint64_t from = h < s.ne ? (uint64_t)a : (uint64_t)b, to = h < s.ne ? (uint64_t)b : (uint64_t)a;
  d[0] = (double)s.v[3 * to] - (double)s.v[3 * from];
  d[1] = (double)s.v[3 * to + 1] - (double)s.v[3 * from + 1];
  d[2] = (double)s.v[3 * to + 2] - (double)s.v[3 * from + 2];
  const double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (!(len > 0.0)) return false;
  d[0] /= len; d[1] /= len; d[2] /= len;
  *src = from;
  return true;
}

__global__ void __launch_bounds__(256) skel_init_kernel(Skel s) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.nv; i += (uint64_t)gridDim.x * blockDim.x) {
    s.first[i] = ~0ull; s.acc[3 * i] = 0; s.acc[3 * i + 1] = 0; s.acc[3 * i + 2] = 0;
  }
}
__global__ void __launch_bounds__(256) skel_first_kernel(Skel s) {
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < 2 * s.ne; h += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t src; double d[3];
    if (half_edge(s, h, &src, d)) atomicMin(&s.first[src], (unsigned long long)h);
  }
}
__global__ void __launch_bounds__(256) skel_sum_kernel(Skel s) {
  for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < 2 * s.ne; h += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t src, src2; double d[3], ref[3];
    if (!half_edge(s, h, &src, d)) continue;
    half_edge(s, s.first[src], &src2, ref);
    const double sign = (d[0] * ref[0] + d[1] * ref[1] + d[2] * ref[2]) < 0.0 ? -1.0 : 1.0;
    const double scale = 1099511627776.0;   // 2^40
    for (int a = 0; a < 3; a++) atomicAdd((unsigned long long*)&s.acc[3 * src + a], (unsigned long long)__double2ll_rn(sign * d[a] * scale));
  }
}
__global__ void __launch_bounds__(256) skel_finish_kernel(Skel s) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.nv; i += (uint64_t)gridDim.x * blockDim.x) {
    const double x = (double)s.acc[3 * i], y = (double)s.acc[3 * i + 1], z = (double)s.acc[3 * i + 2];
    const double n = sqrt(x * x + y * y + z * z);
    if (n > 0.0) { s.out[3 * i] = (float)(x / n); s.out[3 * i + 1] = (float)(y / n); s.out[3 * i + 2] = (float)(z / n); }
    else { s.out[3 * i] = 0.0f; s.out[3 * i + 1] = 0.0f; s.out[3 * i + 2] = 1.0f; }
  }
}

}  // namespace

extern "C" int xs3d_skeleton_tangents(xs3d_volume* v, uint64_t n_vertices, const float* vertices, uint64_t n_edges, const int64_t* edges,
                                      int mem, float* normals) {
  if (!v) return xs_fail(XS3D_ERR_ARG, "null volume");
  if (n_vertices == 0) return XS3D_OK;
  if (!vertices || !normals || (n_edges && !edges)) return xs_fail(XS3D_ERR_ARG, "null argument");
  XS_RET(xs_set_device(v));
  cudaStream_t st = xs_stream(v);
  const size_t a256 = 255;
  const size_t s_first = (n_vertices * 8 + a256) & ~a256, s_acc = (n_vertices * 24 + a256) & ~a256;
  const size_t s_v = (n_vertices * 12 + a256) & ~a256, s_e = (n_edges * 16 + a256) & ~a256;
  char* ws;
  XS_RET(xs_buf(v, 9, s_first + s_acc + (mem == XS3D_HOST ? 2 * s_v + s_e : 0) + 256, (void**)&ws));
  Skel s;
  s.nv = n_vertices; s.ne = n_edges;
  s.first = (unsigned long long*)ws; s.acc = (long long*)(ws + s_first);
  s.v = vertices; s.e = edges; s.out = normals;
  if (mem == XS3D_HOST) {
    float* dv = (float*)(ws + s_first + s_acc);
    float* dout = (float*)(ws + s_first + s_acc + s_v);
    int64_t* de = (int64_t*)(ws + s_first + s_acc + 2 * s_v);
    XS_CK(cudaMemcpyAsync(dv, vertices, n_vertices * 12, cudaMemcpyHostToDevice, st));
    if (n_edges) XS_CK(cudaMemcpyAsync(de, edges, n_edges * 16, cudaMemcpyHostToDevice, st));
    s.v = dv; s.e = de; s.out = dout;
  }
  const int sms = xs_sm_count(v);
  const unsigned gv = (unsigned)std::min<uint64_t>((n_vertices + 255) / 256, (uint64_t)sms * 8);
  const unsigned ge = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((2 * n_edges + 255) / 256, (uint64_t)sms * 8));
  skel_init_kernel<<<gv, 256, 0, st>>>(s);
  if (n_edges) {
    skel_first_kernel<<<ge, 256, 0, st>>>(s);
    skel_sum_kernel<<<ge, 256, 0, st>>>(s);
  }
  skel_finish_kernel<<<gv, 256, 0, st>>>(s);
  XS_CK(cudaGetLastError());
  if (mem == XS3D_HOST) XS_CK(cudaMemcpyAsync(normals, s.out, n_vertices * 12, cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  return XS3D_OK;
}
