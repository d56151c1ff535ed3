// xs_geom.cuh — plane ∩ cube polygon, polygon area, 2x2x2 block value and block adjacency.
//
// Behavioural spec: SURVEY.md Appendix A.  Reference functions replaced (file:line in /root/reference/src):
//   plane_cube_points(S=1)  <- check_intersections_1x1x1   xs3d.hpp:238-359
//   plane_cube_points(S=2)  <- check_intersections_2x2x2   xs3d.hpp:102-236
//   polygon_area          <- xs3d::area::points_to_area   area.hpp:211-228 (+ :14-24, :37-52, :68-143, :147-207)
//   block_value           <- calc_area_at_point_2x2x2     xs3d.hpp:361-440
//   blocks_adjacent       <- is_26_connected              xs3d.hpp:537-646
#pragma once
#include "xs_math.cuh"

namespace xs {

// float32 constants after the reference's C++ promotion rules
#define XS_EPS 1.99999994947575032711029052734375e-05f      /* 0x37a7c5ac : constexpr float epsilon = 2e-5 */
#define XS_FAR1 0.866045415401458740234375f                  /* 0x3f5db527 : float(1.7320508076/2 + eps)   */
#define XS_FAR2 1.73207080364227294921875f                   /* 0x3fddb47f : float(1.7320508076 + eps)     */
#define XS_BOUND1 0.500020027160644531250f                   /* 0x3f000150 : float(0.5 + eps)              */
#define XS_BOUND2 1.00002002716064453125f                    /* 0x3f8000a8 : 1.0f + eps                    */
#define XS_TMAX2 2.0000200271606445312500f                   /* 0x40000054 : 2.0f + eps                    */

struct PlaneGeom {
  V3 pos;      // plane point (unrounded query vertex)
  V3 n;        // unit normal (float32-normalised once)
  V3 w;        // anisotropy
  float inv[3];  // 1/n[a], 0 where n[a]==0   (xs3d.hpp:859-864)
  int max_pts;   // 4 if the normal has a zero component else 6 (xs3d.hpp:299-301)
};

XS_HD void geom_init(PlaneGeom& g, V3 pos, V3 unit_n, V3 w) {
  g.pos = pos; g.n = unit_n; g.w = w;
  g.inv[0] = (unit_n.x == 0.0f) ? 0.0f : fdiv(1.0f, unit_n.x);   // 1.0/p in double then narrowed == 1.0f/p
  g.inv[1] = (unit_n.y == 0.0f) ? 0.0f : fdiv(1.0f, unit_n.y);
  g.inv[2] = (unit_n.z == 0.0f) ? 0.0f : fdiv(1.0f, unit_n.z);
  g.max_pts = ((unit_n.x == 0.0f) + (unit_n.y == 0.0f) + (unit_n.z == 0.0f)) >= 1 ? 4 : 6;
}

// A polygon of up to 6 vertices held in named slots so that, after full unrolling, every access has a
// compile-time index and the whole thing lives in registers (no local-memory stack).
struct Poly {
  V3 p[6];
  int n;
};

XS_HD void poly_push(Poly& q, V3 v) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j == q.n) q.p[j] = v;
  }
  q.n++;
}
XS_HD bool poly_has(const Poly& q, V3 v) {
  bool dup = false;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < q.n) dup = dup || vclose(q.p[j], v);
  }
  return dup;
}

// The plane passes too far from the cube's centre to touch it (xs3d.hpp:270 / :138): the polygon is empty.
XS_HD bool plane_cube_far(int S, float x, float y, float z, V3 pos, V3 n) {
  const V3 low = mk(x, y, z);
  const V3 centre = (S == 1) ? low : vadd(low, mk(0.5f, 0.5f, 0.5f));
  const float dist = fabsf(vdot(vsub(centre, pos), n));
  return dist > (S == 1 ? XS_FAR1 : XS_FAR2);
}

// Conservative pre-filter for the polygon kernel: true only when the plane cannot reach the cube, so that
// plane_cube_points() is certain to find no point (polygon area exactly 0).  A point of the cube is at most
// (S/2)(|nx|+|ny|+|nz|) from the plane through the centre.  The reference accepts hits up to eps = 2e-5 outside the cube
// and computes them in float32: the hit point carries <= 6e-8 of its coordinates, the corner projections <= 3e-7 of the
// distance plane point - cube; the margin below is 0.01 plus 16 and 6 times those bounds (tests/test_hostsim.py probes it
// with planes through and next to corners, edges and faces up to coordinates of 60 000).
// (plane_cube_far — the reference's own test against the circumscribed sphere — lets about 40 % more cubes through.)
#ifndef XS_MISS_MARGIN
#define XS_MISS_MARGIN 0.01f
#endif
XS_HD bool plane_cube_miss(int S, float x, float y, float z, V3 pos, V3 n) {
  const V3 low = mk(x, y, z);
  const V3 centre = (S == 1) ? low : vadd(low, mk(0.5f, 0.5f, 0.5f));
  const V3 rel = vsub(centre, pos);
  const float dist = fabsf(vdot(rel, n));
  const float mag = fadd(fmul(fadd(fadd(x, y), z), 1e-6f), fmul(fadd(fadd(fabsf(rel.x), fabsf(rel.y)), fabsf(rel.z)), 2e-6f));
  const float reach = fadd(fadd(fmul(fadd(fadd(fabsf(n.x), fabsf(n.y)), fabsf(n.z)), S == 1 ? 0.5f : 1.0f), XS_MISS_MARGIN), mag);
  return dist > reach;
}

// Intersect the plane with the 12 edges of the cube of side S whose low voxel is (x,y,z)
// (S=1: the voxel itself, spanning [x-.5,x+.5]; S=2: the block spanning [x-.5,x+1.5]).
// Edge order, corner de-duplication and the 4/6 vertex cap follow xs3d.hpp:183-233 / 307-358.
// S is a run-time value (1 or 2) so that unit voxels and 2x2x2 blocks share one instruction stream in the polygon
// kernel; every constant is selected, never computed, so the arithmetic is that of the two reference functions.
XS_HD void plane_cube_points(int S, float x, float y, float z, const PlaneGeom& g, Poly& q) {
  q.n = 0;
  const float fS = (float)S;
  const V3 low = mk(x, y, z);
  const V3 centre = (S == 1) ? low : vadd(low, mk(0.5f, 0.5f, 0.5f));
  if (plane_cube_far(S, x, y, z, g.pos, g.n)) return;

  const V3 minpt = vadd(low, mk(-0.5f, -0.5f, -0.5f));
  const V3 rel = vsub(g.pos, minpt);
  // the four alternating corners (0,0,0) (0,S,S) (S,0,S) (S,S,0): each edge touches exactly one
  float cproj[4];
  cproj[0] = vdot(rel, g.n);   // rel - (0,0,0) is exact
  cproj[1] = vdot(vsub(rel, mk(0.0f, fS, fS)), g.n);
  cproj[2] = vdot(vsub(rel, mk(fS, 0.0f, fS)), g.n);
  cproj[3] = vdot(vsub(rel, mk(fS, fS, 0.0f)), g.n);
  V3 corner[4];
  corner[0] = minpt;                                   // (0,0,0) + minpt is exact
  corner[1] = vadd(mk(0.0f, fS, fS), minpt);
  corner[2] = vadd(mk(fS, 0.0f, fS), minpt);
  corner[3] = vadd(mk(fS, fS, 0.0f), minpt);

  const float bound = (S == 1) ? XS_BOUND1 : XS_BOUND2;
  const float tmax = (S == 1) ? XS_BOUND2 : XS_TMAX2;   // 1+eps / 2+eps
  bool done = false;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const int k = i & 3, a = i >> 2;
    const float proj = cproj[k];
    const V3 c = corner[k];
    // Evaluate the edge unconditionally and fold the reference's chain of `continue`s (xs3d.hpp:308-345) into
    // predicates: 32 different cubes per warp would otherwise diverge at every one of them.
    const float na = (a == 0) ? g.n.x : (a == 1 ? g.n.y : g.n.z);
    const float t = fmul(proj, g.inv[a]);
    const float at = fabsf(t);
    V3 hit = c;
    if (a == 0) hit.x = fadd(c.x, t);
    else if (a == 1) hit.y = fadd(c.y, t);
    else hit.z = fadd(c.z, t);
    const bool on_plane = (proj == 0.0f);
    const bool inside = !(fabsf(fsub(hit.x, centre.x)) >= bound) & !(fabsf(fsub(hit.y, centre.y)) >= bound) & !(fabsf(fsub(hit.z, centre.z)) >= bound);
    const bool edge_hit = !done & !on_plane & (na != 0.0f) & !(at > tmax) & inside;
    const bool corner_hit = !done & on_plane & (i < 4);
    const bool oncorner = (at < XS_EPS) | (fabsf(fsub(at, fS)) < XS_EPS);
    const V3 cand = on_plane ? c : hit;
    bool take = edge_hit | corner_hit;
    if (take & (corner_hit | oncorner)) {       // rare: only points that may coincide with a cube corner are de-duplicated
      if (poly_has(q, cand)) take = false;
    }
    if (take) {
      poly_push(q, cand);
      if (edge_hit & (q.n >= g.max_pts)) done = true;   // the reference only checks the cap after an edge hit (:354-356)
    }
  }
}

XS_HD void cswap(float* key, V3* v, int a, int b) {
  if (key[a] > key[b]) {
    float tk = key[a]; key[a] = key[b]; key[b] = tk;
    V3 tv = v[a]; v[a] = v[b]; v[b] = tv;
  }
}

// Anisotropy-scaled area of the convex polygon q lying in the plane (area.hpp:211-228).
XS_HD float polygon_area(const Poly& q, const PlaneGeom& g) {
  const int np = q.n;
  if (np < 3) return 0.0f;
  if (np == 3) {
    V3 a = vmul(vsub(q.p[1], q.p[0]), g.w), b = vmul(vsub(q.p[2], q.p[0]), g.w);
    return fmul(vnorm(vcross(a, b)), 0.5f);
  }
  V3 cen = mk(0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    if (i < np) cen = vadd(cen, q.p[i]);
  }
  cen = vdivs(cen, (float)np);
  V3 sp[6];
  float key[6];
#pragma unroll
  for (int i = 0; i < 6; i++) sp[i] = vsub(q.p[i < np ? i : 0], cen);
  const V3 prime = vdivs(sp[0], vnorm(sp[0]));
  V3 basis = vcross(prime, g.n);
  basis = vdivs(basis, vnorm(basis));
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float c = fdiv(vdot(sp[i], prime), vnorm(sp[i]));
    key[i] = (vdot(sp[i], basis) < 0.0f) ? fsub(c, 1.0f) : fsub(1.0f, c);
  }
  // fixed comparator networks of area.hpp:68-143 (strict >, so ties keep their order)
  if (np == 4) {
    cswap(key, sp, 0, 2); cswap(key, sp, 1, 3); cswap(key, sp, 0, 1); cswap(key, sp, 2, 3); cswap(key, sp, 1, 2);
  } else if (np == 5) {
    cswap(key, sp, 0, 3); cswap(key, sp, 1, 4); cswap(key, sp, 0, 2); cswap(key, sp, 1, 3); cswap(key, sp, 0, 1);
    cswap(key, sp, 2, 4); cswap(key, sp, 1, 2); cswap(key, sp, 3, 4); cswap(key, sp, 2, 3);
  } else {
    cswap(key, sp, 0, 5); cswap(key, sp, 1, 3); cswap(key, sp, 2, 4); cswap(key, sp, 1, 2); cswap(key, sp, 3, 4);
    cswap(key, sp, 0, 3); cswap(key, sp, 2, 5); cswap(key, sp, 0, 1); cswap(key, sp, 2, 3); cswap(key, sp, 4, 5);
    cswap(key, sp, 1, 2); cswap(key, sp, 3, 4);
  }
#pragma unroll
  for (int i = 0; i < 6; i++) sp[i] = vmul(sp[i], g.w);
  float area = 0.0f;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (i + 1 < np) area = fadd(area, vnorm(vcross(sp[i], sp[i + 1])));
  }
  const V3 last = (np == 4) ? sp[3] : (np == 5 ? sp[4] : sp[5]);
  area = fadd(area, vnorm(vcross(sp[0], last)));
  return fmul(area, 0.5f);
}

XS_HD float voxel_area(float x, float y, float z, const PlaneGeom& g) {
  Poly q;
  plane_cube_points(1, x, y, z, g, q);
  return polygon_area(q, g);
}
// kind = cube side (1: unit voxel, 2: 2x2x2 block), one code path for both
XS_HD float cube_polygon_area(int S, float x, float y, float z, const PlaneGeom& g) {
  Poly q;
  plane_cube_points(S, x, y, z, g, q);
  return polygon_area(q, g);
}
XS_HD float block_polygon_area(float bx, float by, float bz, const PlaneGeom& g) {
  Poly q;
  plane_cube_points(2, bx, by, bz, g, q);
  return polygon_area(q, g);
}

// Value of the 2x2x2 block with low (even) voxel (bx,by,bz) and occupancy mask `cube`
// (bit = dx + 2dy + 4dz): block polygon minus absent voxels when >=5 voxels are set,
// else the sum of the present voxels, always in ascending bit order.
XS_HD float block_value(uint32_t cube, float bx, float by, float bz, const PlaneGeom& g) {
  const V3 centre = vadd(mk(bx, by, bz), mk(0.5f, 0.5f, 0.5f));
  if (fabsf(vdot(vsub(centre, g.pos), g.n)) > XS_FAR2) return 0.0f;
  float area = 0.0f;
#ifdef __CUDA_ARCH__
  const int pop = __popc(cube & 0xffu);
#else
  const int pop = __builtin_popcount(cube & 0xffu);
#endif
  if (pop >= 5) {
    area = block_polygon_area(bx, by, bz, g);
    if (area == 0.0f) return area;
    for (int b = 0; b < 8; b++) {
      if (!((cube >> b) & 1u))
        area = fsub(area, voxel_area(fadd(bx, (float)(b & 1)), fadd(by, (float)((b >> 1) & 1)), fadd(bz, (float)(b >> 2)), g));
    }
  } else {
    for (int b = 0; b < 8; b++) {
      if ((cube >> b) & 1u)
        area = fadd(area, voxel_area(fadd(bx, (float)(b & 1)), fadd(by, (float)((b >> 1) & 1)), fadd(bz, (float)(b >> 2)), g));
    }
  }
  return area;
}

// Face masks of a block: voxels with coordinate a == v, as bits dx + 2dy + 4dz.
XS_HD uint32_t face_mask(int ox, int oy, int oz, bool candidate_side) {
  // centre side: o<0 -> coordinate 0, o>0 -> coordinate 1; candidate side is the opposite face
  uint32_t m = 0xffu;
  const uint32_t X0 = 0x55u, X1 = 0xaau, Y0 = 0x33u, Y1 = 0xccu, Z0 = 0x0fu, Z1 = 0xf0u;
  if (ox < 0) m &= candidate_side ? X1 : X0;
  if (ox > 0) m &= candidate_side ? X0 : X1;
  if (oy < 0) m &= candidate_side ? Y1 : Y0;
  if (oy > 0) m &= candidate_side ? Y0 : Y1;
  if (oz < 0) m &= candidate_side ? Z1 : Z0;
  if (oz > 0) m &= candidate_side ? Z0 : Z1;
  return m;
}

// Two neighbouring blocks count as connected when each has a set voxel on the side facing the
// other.  One entry of the reference's table deviates from that rule and is reproduced:
// o = (-,+,0) uses the centre mask 0b01000010 (xs3d.hpp:569).
XS_HD bool blocks_adjacent(uint32_t centre, uint32_t cand, int ox, int oy, int oz) {
  if ((ox | oy | oz) == 0) return true;
  uint32_t cm = face_mask(ox, oy, oz, false);
  const uint32_t km = face_mask(ox, oy, oz, true);
  if (ox < 0 && oy > 0 && oz == 0) cm = 0x42u;
  return ((cand & km) != 0u) && ((centre & cm) != 0u);
}

}  // namespace xs
