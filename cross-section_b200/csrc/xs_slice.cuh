// xs_slice.cuh — oblique label resampling (xs3d.slice), shared host/device logic.
//
// Replaces (file:line in /root/reference/src): cross_section_projection<LABEL> xs3d.hpp:1473-1624,
// create_orthonormal_basis :1428-1471, Bbox2d :1274-1293 and the sizing / cut-out logic of
// fastxs3d.cpp:137-196.
//
// The reference flood-fills (4-connected) the lattice cells that are inside the volume and within the
// crop radius, starting from the centre cell, gathering labels[round(cur)] as it goes.  Every validity
// test is monotone along a lattice row, so the valid cells of a row form ONE run [a_j, b_j] (up to
// float noise, which is detected).  The fill then reaches exactly the maximal range of consecutive
// rows around the centre row whose runs overlap column-wise.  That turns the sequential fill into
//   (1) a dense, embarrassingly parallel validity bitmap,  (2) a per-row run summary,
//   (3) a 1-D scan over rows,                               (4) a dense gather into the cut-out.
// Rows with more than one run that the fill can reach switch to an exact iterative fill (fallback).
#pragma once
#include "xs_geom.cuh"

namespace xs {

struct SliceCtx {
  V3 pos, b1, b2;        // basis already multiplied by min(w)/w (xs3d.hpp:1491-1496, 1531-1532)
  float lim[3];          // float(s) - 0.5
  float crop_sq;         // xs3d.hpp:1489
  int sx, sy, sz;
  long long psx;         // lattice side (fastxs3d.cpp:137-154 == xs3d.hpp:1505)
  long long c;           // centre index psx/2
  int active;            // 0: vertex outside the volume or crop <= 0 -> 1x1 zero output (bbox 0,0,0,0 / centre)
  int pos_inside;
  // conservative window of lattice cells that can be valid
  long long wx0, wy0;
  int ww, wh;
};

XS_HD int argmax_abs(V3 a) {   // vec.hpp:102 applied to abs()
  const float x = fabsf(a.x), y = fabsf(a.y), z = fabsf(a.z);
  if (x >= y) return (x >= z) ? 0 : 2;
  return (y >= z) ? 1 : 2;
}

// xs3d.hpp:1428-1471 (n is the unit normal)
XS_HD void slice_basis(V3 n, bool positive, V3* o1, V3* o2) {
  V3 b1 = vcross(n, mk(0.0f, 1.0f, 0.0f));
  if (vnull(b1)) b1 = vcross(n, mk(1.0f, 0.0f, 0.0f));
  b1 = vdivs(b1, vnorm(b1));
  V3 b2 = vcross(n, b1);
  b2 = vdivs(b2, vnorm(b2));
  if (argmax_abs(b2) < argmax_abs(b1)) { V3 t = b1; b1 = b2; b2 = t; }
  if (positive) {
    V3 zone = mk(1.0f, 1.0f, 1.0f);
    if (vdot(n, zone) < 0.0f) zone = mk(-1.0f, -1.0f, -1.0f);
    if (vdot(b1, zone) < 0.0f) b1 = mk(-b1.x, -b1.y, -b1.z);
    if (vdot(b2, zone) < 0.0f) b2 = mk(-b2.x, -b2.y, -b2.z);
  }
  *o1 = b1; *o2 = b2;
}

// Lattice side for a volume / anisotropy / crop (fastxs3d.cpp:137-154).
XS_HD long long slice_plane_size(int sx, int sy, int sz, V3 w, float crop) {
  const float wmin = fminf(fminf(w.x, w.y), w.z);
  const V3 wn = vdivs(w, wmin);
  const float wmax = fmaxf(fmaxf(fabsf(wn.x), fabsf(wn.y)), fabsf(wn.z));
  const long long distortion = (long long)ceil((double)wmax);
  long long largest = sx > sy ? sx : sy;
  if (sz > largest) largest = sz;
  if ((float)largest > crop && crop >= 0.0f) largest = (long long)ceilf(crop);
  return (distortion * 2 * 97 * largest) / 56 + 1;
}

XS_HD void slice_init(SliceCtx& s, int sx, int sy, int sz, V3 pos, V3 nraw, V3 w, bool positive, float crop) {
  s.sx = sx; s.sy = sy; s.sz = sz;
  s.pos = pos;
  const float wmin = fminf(fminf(w.x, w.y), w.z);
  s.crop_sq = fdiv(fdiv(fmul(crop, crop), wmin), wmin);
  const V3 wn = vdivs(w, wmin);
  const V3 scale = mk(fdiv(1.0f, wn.x), fdiv(1.0f, wn.y), fdiv(1.0f, wn.z));
  s.psx = slice_plane_size(sx, sy, sz, w, crop);
  s.c = s.psx / 2;
  s.lim[0] = fsub((float)sx, 0.5f); s.lim[1] = fsub((float)sy, 0.5f); s.lim[2] = fsub((float)sz, 0.5f);
  s.pos_inside = !(pos.x < 0.0f || pos.x >= (float)sx || pos.y < 0.0f || pos.y >= (float)sy || pos.z < 0.0f || pos.z >= (float)sz);
  s.active = s.pos_inside && (crop > 0.0f);
  s.wx0 = s.wy0 = s.c; s.ww = s.wh = 1;
  if (!s.active) return;
  const V3 n = vdivs(nraw, vnorm(nraw));
  V3 b1, b2;
  slice_basis(n, positive, &b1, &b2);
  s.b1 = vmul(b1, scale); s.b2 = vmul(b2, scale);
  // window: valid cells satisfy dx^2 + dy^2 <= min(crop_sq, D^2) / scale_min^2 and lie inside the
  // projection of the (anisotropy-unscaled) volume box on (b1, b2).  Plain float arithmetic + padding.
  const float smin = fminf(fminf(scale.x, scale.y), scale.z);
  float d2 = 0.0f;
  float lo[2] = { 0.0f, 0.0f }, hi[2] = { 0.0f, 0.0f };
  for (int k = 0; k < 8; k++) {
    const float qx = (((k & 1) ? (float)sx - 0.5f : -0.5f) - pos.x);
    const float qy = (((k & 2) ? (float)sy - 0.5f : -0.5f) - pos.y);
    const float qz = (((k & 4) ? (float)sz - 0.5f : -0.5f) - pos.z);
    const float dd = qx * qx + qy * qy + qz * qz;
    d2 = dd > d2 ? dd : d2;
    const float ux = qx * wn.x, uy = qy * wn.y, uz = qz * wn.z;   // undo the per-axis step scale
    const float a = ux * b1.x + uy * b1.y + uz * b1.z;
    const float b = ux * b2.x + uy * b2.y + uz * b2.z;
    lo[0] = a < lo[0] ? a : lo[0]; hi[0] = a > hi[0] ? a : hi[0];
    lo[1] = b < lo[1] ? b : lo[1]; hi[1] = b > hi[1] ? b : hi[1];
  }
  float r2 = d2;
  if (s.crop_sq < r2) r2 = s.crop_sq;
  const float rad = sqrtf(r2) / smin;
  long long x0, x1, y0, y1;
  {
    const float padx = 3.0f + 0.001f * (hi[0] - lo[0]), pady = 3.0f + 0.001f * (hi[1] - lo[1]);
    float ax = lo[0] - padx, bx = hi[0] + padx, ay = lo[1] - pady, by = hi[1] + pady;
    const float rr = rad * 1.001f + 3.0f;
    if (-rr > ax) ax = -rr;
    if (rr < bx) bx = rr;
    if (-rr > ay) ay = -rr;
    if (rr < by) by = rr;
    x0 = s.c + (long long)floorf(ax); x1 = s.c + (long long)ceilf(bx);
    y0 = s.c + (long long)floorf(ay); y1 = s.c + (long long)ceilf(by);
  }
  if (x0 < 0) x0 = 0;
  if (y0 < 0) y0 = 0;
  if (x1 > s.psx - 1) x1 = s.psx - 1;
  if (y1 > s.psx - 1) y1 = s.psx - 1;
  s.wx0 = x0; s.wy0 = y0;
  s.ww = (int)(x1 - x0 + 1); s.wh = (int)(y1 - y0 + 1);
}

// Validity of lattice cell (x,y) (xs3d.hpp:1567-1581) and the linear index of its voxel (:1583-1600).
// A coordinate of exactly -0.5 passes the reference's bounds test and rounds to -1 (UB cast); the
// cell is kept (so output SHAPES match the reference) and gathers from the clamped voxel.
XS_HD bool slice_cell(const SliceCtx& s, long long x, long long y, bool c_order, size_t* loc) {
  if (x < 0 || y < 0 || x >= s.psx || y >= s.psx) return false;
  const float dx = fsub((float)x, (float)s.c), dy = fsub((float)y, (float)s.c);
  const V3 delta = vadd(vscale(s.b1, dx), vscale(s.b2, dy));
  const V3 cur = vadd(s.pos, delta);
  if (cur.x < -0.5f || cur.y < -0.5f || cur.z < -0.5f) return false;
  if (cur.x >= s.lim[0] || cur.y >= s.lim[1] || cur.z >= s.lim[2]) return false;
  if (vnorm2(delta) >= s.crop_sq) return false;
  if (!(cur.x == cur.x) || !(cur.y == cur.y) || !(cur.z == cur.z)) return false;
  if (loc) {
    V3 r = vround(cur);
    if (r.x < 0.0f) r.x = 0.0f;
    if (r.y < 0.0f) r.y = 0.0f;
    if (r.z < 0.0f) r.z = 0.0f;
    const size_t vx = (size_t)r.x, vy = (size_t)r.y, vz = (size_t)r.z;
    *loc = c_order ? (vz + (size_t)s.sz * (vy + (size_t)s.sy * vx)) : (vx + (size_t)s.sx * (vy + (size_t)s.sy * vz));
  }
  return true;
}

// Per-row summary of a bitmap row: first / last set column (window coordinates) and number of runs.
struct RowRuns {
  int a, b, runs;
};

XS_HD RowRuns row_runs(const uint32_t* row, int words) {
  RowRuns r;
  r.a = 0x7fffffff; r.b = -1; r.runs = 0;
  uint32_t carry = 0u;
  for (int i = 0; i < words; i++) {
    const uint32_t w = row[i];
    if (w) {
#ifdef __CUDA_ARCH__
      const int lo = __ffs(w) - 1, hi = 31 - __clz(w);
      r.runs += __popc(w & ~((w << 1) | carry));
#else
      const int lo = __builtin_ffs(w) - 1, hi = 31 - __builtin_clz(w);
      r.runs += __builtin_popcount(w & ~((w << 1) | carry));
#endif
      if (r.a == 0x7fffffff) r.a = i * 32 + lo;
      r.b = i * 32 + hi;
    }
    carry = w >> 31;
  }
  return r;
}

struct SliceResult {
  int y0, y1;        // reached row range (window coordinates), y0 > y1 when nothing is reached
  int x0, x1;        // column bbox of the reached cells
  int fallback;      // a reachable row has more than one run: use the exact iterative fill
};

// 1-D scan over the row summaries from the centre row outwards (window coordinates cx, cy).
XS_HD SliceResult scan_rows(const RowRuns* rows, int nrows, int cx, int cy) {
  SliceResult r;
  r.y0 = 1; r.y1 = 0; r.x0 = cx; r.x1 = cx; r.fallback = 0;
  if (cy < 0 || cy >= nrows) return r;
  const RowRuns c = rows[cy];
  if (c.runs == 0 || cx < c.a || cx > c.b) return r;   // the centre cell itself is not valid
  if (c.runs > 1) { r.fallback = 1; return r; }
  r.y0 = r.y1 = cy; r.x0 = c.a; r.x1 = c.b;
  for (int dir = -1; dir <= 1; dir += 2) {
    int pa = c.a, pb = c.b;
    for (int y = cy + dir; y >= 0 && y < nrows; y += dir) {
      const RowRuns q = rows[y];
      if (q.runs == 0 || q.a > pb || q.b < pa) break;      // no column in common with the previous row
      if (q.runs > 1) { r.fallback = 1; return r; }
      if (dir < 0) r.y0 = y; else r.y1 = y;
      if (q.a < r.x0) r.x0 = q.a;
      if (q.b > r.x1) r.x1 = q.b;
      pa = q.a; pb = q.b;
    }
  }
  return r;
}

}  // namespace xs
