// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A thin extern "C" window onto the UNMODIFIED reference headers, compiled from where they
// lie under /root/reference/src (see oracle/Makefile).  Nothing from the reference is copied
// into this repository: this file only #includes the reference headers and forwards calls.
// The resulting oracle/_ref/libxs3d_ref.so is git-ignored but travels to the GPU box.
//
// Exposed:
//   * the three public entry points of the hot path (xs3d.hpp:1305 cross_sectional_area,
//     :1367 cross_section, :1473 cross_section_projection, auxiliary.hpp:94/193 slow variants)
//   * the internal geometry helpers, so kernel-level goldens can be pinned
//     (xs3d.hpp:26 compute_cube, :102 check_intersections_2x2x2, :238 check_intersections_1x1x1,
//      :361 calc_area_at_point_2x2x2, :537 is_26_connected, :1428 create_orthonormal_basis,
//      area.hpp:211 points_to_area)
//   * a std::thread fan-out of the single-query reference over a batch (the "all host cores"
//     CPU baseline of BASELINE.md §3; the non-persistent path is re-entrant because its only
//     static scratch is thread_local, area.hpp:74,99,128,181).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>
#include <limits>

#include "xs3d.hpp"   // -I/root/reference/src

extern "C" {

void ref_area(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
              const float* p, const float* n, const float* w,
              int slow_method, int use_persistent_data,
              float* area, uint8_t* contact) {
  std::tuple<float, uint8_t> r;
  if (slow_method) {
    r = xs3d::cross_sectional_area_slow(binimg, sx, sy, sz, p[0], p[1], p[2],
                                        n[0], n[1], n[2], w[0], w[1], w[2]);
  } else {
    r = xs3d::cross_sectional_area(binimg, sx, sy, sz, p[0], p[1], p[2],
                                   n[0], n[1], n[2], w[0], w[1], w[2],
                                   use_persistent_data != 0);
  }
  *area = std::get<0>(r);
  *contact = std::get<1>(r);
}

void ref_set_shape(uint64_t sx, uint64_t sy, uint64_t sz) { xs3d::set_shape(sx, sy, sz); }
void ref_clear_shape() { xs3d::clear_shape(); }

// Fan the single-query reference out over `nthreads` host threads, one contiguous chunk each.
void ref_area_batch(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
                    uint64_t nq, const float* pos, const float* nrm, const float* w,
                    float* area, uint8_t* contact, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  auto work = [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      auto r = xs3d::cross_sectional_area(binimg, sx, sy, sz,
                                          pos[3*i], pos[3*i+1], pos[3*i+2],
                                          nrm[3*i], nrm[3*i+1], nrm[3*i+2],
                                          w[0], w[1], w[2], false);
      area[i] = std::get<0>(r);
      contact[i] = std::get<1>(r);
    }
  };
  if (nthreads == 1) { work(0, nq); return; }
  std::vector<std::thread> pool;
  uint64_t chunk = (nq + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    uint64_t lo = std::min<uint64_t>(nq, t * chunk), hi = std::min<uint64_t>(nq, lo + chunk);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
}

// out must hold sx*sy*sz floats (method 0/1) or ceil-halved dims (method 2) and be zeroed
// by the caller, exactly as fastxs3d.cpp:42-44 does before dispatching.
void ref_section(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
                 const float* p, const float* n, const float* w, int method,
                 float* out, uint8_t* contact) {
  std::tuple<float*, uint8_t> r;
  if (method == 0) {
    r = xs3d::cross_section(binimg, sx, sy, sz, p[0], p[1], p[2], n[0], n[1], n[2],
                            w[0], w[1], w[2], out);
  } else if (method == 1) {
    r = xs3d::cross_section_slow(binimg, sx, sy, sz, p[0], p[1], p[2], n[0], n[1], n[2],
                                 w[0], w[1], w[2], out);
  } else {
    r = xs3d::cross_section_slow_2x2x2(binimg, sx, sy, sz, p[0], p[1], p[2], n[0], n[1], n[2],
                                       w[0], w[1], w[2], out);
  }
  *contact = std::get<1>(r);
}

}  // extern "C"

// Raw projection: `out` holds psx*psx LABELs, zeroed by the caller; bbox = {x_min,x_max,y_min,y_max}
// exactly as cross_section_projection returns it (before the +1 of fastxs3d.cpp:176-177).
template <typename T>
static void projection_t(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int c_order,
                         const float* p, const float* n, const float* w, int positive_basis,
                         float crop, void* out, int64_t* bbox) {
  auto r = xs3d::cross_section_projection<T>(
      reinterpret_cast<const T*>(labels), sx, sy, sz, c_order != 0,
      p[0], p[1], p[2], n[0], n[1], n[2], w[0], w[1], w[2],
      positive_basis != 0, crop, reinterpret_cast<T*>(out));
  xs3d::Bbox2d b = std::get<1>(r);
  bbox[0] = b.x_min; bbox[1] = b.x_max; bbox[2] = b.y_min; bbox[3] = b.y_max;
}

extern "C" {

int ref_projection(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int itemsize,
                   int c_order, const float* p, const float* n, const float* w,
                   int positive_basis, float crop, void* out, int64_t* bbox) {
  switch (itemsize) {
    case 1: projection_t<uint8_t>(labels, sx, sy, sz, c_order, p, n, w, positive_basis, crop, out, bbox); return 0;
    case 2: projection_t<uint16_t>(labels, sx, sy, sz, c_order, p, n, w, positive_basis, crop, out, bbox); return 0;
    case 4: projection_t<uint32_t>(labels, sx, sy, sz, c_order, p, n, w, positive_basis, crop, out, bbox); return 0;
    case 8: projection_t<uint64_t>(labels, sx, sy, sz, c_order, p, n, w, positive_basis, crop, out, bbox); return 0;
  }
  return -1;
}

// ---------------------------------------------------------------- internals (kernel-level goldens)

static void make_proj(const float* n, std::vector<float>& proj, std::vector<float>& inv) {
  Vec3 normal(n[0], n[1], n[2]);
  proj = { ihat.dot(normal), jhat.dot(normal), khat.dot(normal) };
  inv.resize(3);
  for (int i = 0; i < 3; i++) inv[i] = (proj[i] == 0) ? 0 : 1.0 / proj[i];
}

// normal is used as given (callers normalise in float32 first if they want the public-path value)
int ref_check_1x1x1(uint64_t x, uint64_t y, uint64_t z, const float* p, const float* n, float* pts_out) {
  std::vector<float> proj, inv; make_proj(n, proj, inv);
  std::vector<Vec3> pts; pts.reserve(12);
  check_intersections_1x1x1(pts, x, y, z, Vec3(p[0],p[1],p[2]), Vec3(n[0],n[1],n[2]), proj, inv);
  for (size_t i = 0; i < pts.size(); i++) { pts_out[3*i] = pts[i].x; pts_out[3*i+1] = pts[i].y; pts_out[3*i+2] = pts[i].z; }
  return (int)pts.size();
}

int ref_check_2x2x2(uint64_t x, uint64_t y, uint64_t z, const float* p, const float* n, float* pts_out, uint16_t* edges) {
  std::vector<float> proj, inv; make_proj(n, proj, inv);
  std::vector<Vec3> pts; pts.reserve(12);
  *edges = check_intersections_2x2x2(pts, x, y, z, Vec3(p[0],p[1],p[2]), Vec3(n[0],n[1],n[2]), proj, inv);
  for (size_t i = 0; i < pts.size(); i++) { pts_out[3*i] = pts[i].x; pts_out[3*i+1] = pts[i].y; pts_out[3*i+2] = pts[i].z; }
  return (int)pts.size();
}

float ref_points_to_area(const float* pts_in, int npts, const float* w, const float* n) {
  std::vector<Vec3> pts;
  for (int i = 0; i < npts; i++) pts.emplace_back(pts_in[3*i], pts_in[3*i+1], pts_in[3*i+2]);
  return xs3d::area::points_to_area(pts, Vec3(w[0],w[1],w[2]), Vec3(n[0],n[1],n[2]));
}

float ref_voxel_area(uint64_t x, uint64_t y, uint64_t z, const float* p, const float* n, const float* w) {
  std::vector<float> proj, inv; make_proj(n, proj, inv);
  std::vector<Vec3> pts; pts.reserve(12);
  Vec3 normal(n[0], n[1], n[2]);
  check_intersections_1x1x1(pts, x, y, z, Vec3(p[0],p[1],p[2]), normal, proj, inv);
  return xs3d::area::points_to_area(pts, Vec3(w[0],w[1],w[2]), normal);
}

float ref_block_area(uint8_t cube, const float* cur, const float* p, const float* n, const float* w) {
  std::vector<float> proj, inv; make_proj(n, proj, inv);
  std::vector<Vec3> pts; pts.reserve(12);
  return calc_area_at_point_2x2x2(cube, 0, 0, 0, Vec3(cur[0],cur[1],cur[2]), Vec3(p[0],p[1],p[2]),
                                  Vec3(n[0],n[1],n[2]), Vec3(w[0],w[1],w[2]), pts, proj, inv);
}

int ref_is_26_connected(uint8_t center, uint8_t candidate, int x, int y, int z) {
  return is_26_connected(center, candidate, x, y, z) ? 1 : 0;
}

uint8_t ref_compute_cube(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
                         uint64_t x, uint64_t y, uint64_t z) {
  return compute_cube(binimg, sx, sy, sz, x, y, z);
}

void ref_slice_basis(const float* n, int positive_basis, float* b1, float* b2) {
  auto t = xs3d::create_orthonormal_basis(Vec3(n[0],n[1],n[2]), positive_basis != 0);
  Vec3 a = std::get<0>(t), b = std::get<1>(t);
  b1[0] = a.x; b1[1] = a.y; b1[2] = a.z; b2[0] = b.x; b2[1] = b.y; b2[2] = b.z;
}

}  // extern "C"
