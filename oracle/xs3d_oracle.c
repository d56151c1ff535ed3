/* oracle/xs3d_oracle.c — CPU RESTATEMENT of the xs3d cross-section hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; it is the
 * checker, never the product and never the thing measured.  It is an independent restatement in
 * plain C of the algorithm in /root/reference/src (cited per function), written in the DECOMPOSED
 * form of SURVEY.md §7.3 (component -> DFS preorder rank -> first-touch block ownership ->
 * two-level ordered float32 sums) rather than as the reference's fused flood fill, so pinning it
 * bit-for-bit against the real reference (oracle/_ref, tests/test_oracle_vs_ref.py) also validates
 * the decomposition the CUDA kernels implement.
 *
 * Parity: PINNED — against the compiled reference itself on randomised inputs (when /root/reference
 * is present) and against the committed golden vectors in tests/golden/ (always).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction: the reference binary has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
/* vec.hpp:93 — (x*ox + y*oy) + z*oz, each product rounded to float */
static inline float vdot(v3 a, v3 b) { float p = a.x * b.x; float q = a.y * b.y; float r = a.z * b.z; float s = p + q; return s + r; }
static inline float vnorm2(v3 a) { return vdot(a, a); }
static inline float vnorm(v3 a) { return sqrtf(vnorm2(a)); }                 /* vec.hpp:114 */
static inline v3 vcross(v3 a, v3 b) {                                        /* vec.hpp:123 */
  float x1 = a.y * b.z, x2 = a.z * b.y, y1 = a.z * b.x, y2 = a.x * b.z, z1 = a.x * b.y, z2 = a.y * b.x;
  return V(x1 - x2, y1 - y2, z1 - z2);
}
static inline v3 vround(v3 a) { return V(roundf(a.x), roundf(a.y), roundf(a.z)); }   /* vec.hpp:99 */
static inline int vclose(v3 a, v3 b) { return (double)vnorm2(vsub(a, b)) < 1e-4; }   /* vec.hpp:120 */
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

#define XS_EPS 2e-5f  /* xs3d.hpp:137,269,381: constexpr float epsilon = 2e-5 */

/* Plane ∩ the 12 edges of an axis-aligned cube of side S (S=1: the voxel centred on (x,y,z),
 * xs3d.hpp:238-359; S=2: the 2x2x2 block whose low voxel is (x,y,z), xs3d.hpp:102-236).
 * Returns the number of polygon vertices written to pts (<= 6). */
static int plane_cube_points(int S, uint64_t x, uint64_t y, uint64_t z, v3 pos, v3 n, const float inv[3], v3* pts) {
  const float fS = (float)S;
  v3 lowvox = V((float)x, (float)y, (float)z);
  v3 centre = (S == 1) ? lowvox : vadd(lowvox, V(0.5f, 0.5f, 0.5f));
  /* early out: distance from the cube centre to the plane (xs3d.hpp:270-277 / 138-145).  The
   * thresholds are folded in double and narrowed: float(1.7320508076/2 + 2e-5f), float(1.7320508076 + 2e-5f) */
  const float far1 = (float)(1.7320508076 / 2 + (double)XS_EPS);
  const float far2 = (float)(1.7320508076 + (double)XS_EPS);
  float dist = fabsf(vdot(vsub(centre, pos), n));
  if (dist > (S == 1 ? far1 : far2)) return 0;

  v3 minpt = vadd(lowvox, V(-0.5f, -0.5f, -0.5f));
  v3 rel = vsub(pos, minpt);
  /* the four alternating corners; every edge of the cube touches exactly one of them */
  const v3 corner[4] = { V(0, 0, 0), V(0, fS, fS), V(fS, 0, fS), V(fS, fS, 0) };
  float cproj[4];
  for (int k = 0; k < 4; k++) cproj[k] = vdot(vsub(rel, corner[k]), n);

  const int zero_dims = (n.x == 0) + (n.y == 0) + (n.z == 0);
  const int max_pts = zero_dims >= 1 ? 4 : 6;
  /* in-cube bound: S=1 `0.5 + epsilon` (double add, narrowed); S=2 `1 + epsilon` (float add) */
  const float bound = (S == 1) ? (float)(0.5 + (double)XS_EPS) : (1.0f + XS_EPS);
  const float tmax = fS + XS_EPS;

  int np = 0;
  for (int i = 0; i < 12; i++) {
    const int k = i & 3, a = i >> 2;
    const float proj = cproj[k];
    v3 c = vadd(corner[k], minpt);
    if (proj == 0) {              /* the corner itself lies on the plane */
      if (i < 4) {
        int dup = 0;
        for (int j = 0; j < np; j++) dup |= vclose(pts[j], c);
        if (!dup) pts[np++] = c;
      }
      continue;
    }
    if (comp(n, a) == 0) continue;   /* edge parallel to the plane */
    const float t = proj * inv[a];
    if (fabsf(t) > tmax) continue;
    v3 hit = c;
    if (a == 0) hit.x = c.x + t; else if (a == 1) hit.y = c.y + t; else hit.z = c.z + t;
    if (fabsf(hit.x - centre.x) >= bound) continue;
    if (fabsf(hit.y - centre.y) >= bound) continue;
    if (fabsf(hit.z - centre.z) >= bound) continue;
    const int oncorner = (fabsf(t) < XS_EPS) || (fabsf(fabsf(t) - fS) < XS_EPS);
    int dup = 0;
    if (oncorner) for (int j = 0; j < np; j++) dup |= vclose(pts[j], hit);
    if (!dup) {
      pts[np++] = hit;
      if (np >= max_pts) break;
    }
  }
  return np;
}

/* area.hpp:211-228 points_to_area (+ :14-24 triangle, :167-207 polygon, :37-52 keys, :68-143 networks) */
static float polygon_area(const v3* pts, int np, v3 w, v3 n) {
  if (np < 3) return 0.0f;
  if (np == 3) {
    v3 a = vmul(vsub(pts[1], pts[0]), w), b = vmul(vsub(pts[2], pts[0]), w);
    return vnorm(vcross(a, b)) * 0.5f;
  }
  v3 cen = V(0, 0, 0);
  for (int i = 0; i < np; i++) cen = vadd(cen, pts[i]);
  cen = vdivs(cen, (float)np);
  v3 sp[6];
  float key[6];
  for (int i = 0; i < np; i++) sp[i] = vsub(pts[i], cen);
  v3 prime = vdivs(sp[0], vnorm(sp[0]));
  v3 basis = vcross(prime, n);
  basis = vdivs(basis, vnorm(basis));
  for (int i = 0; i < np; i++) {
    float c = vdot(sp[i], prime) / vnorm(sp[i]);
    key[i] = (vdot(sp[i], basis) < 0) ? (c - 1) : (1 - c);
  }
  /* fixed comparator networks; the trailing comparators move only the vectors (CMP_SWAP_FAST) —
   * harmless because no later comparator reads those keys */
  static const int8_t net4[][2] = {{0,2},{1,3},{0,1},{2,3},{1,2}};
  static const int8_t net5[][2] = {{0,3},{1,4},{0,2},{1,3},{0,1},{2,4},{1,2},{3,4},{2,3}};
  static const int8_t net6[][2] = {{0,5},{1,3},{2,4},{1,2},{3,4},{0,3},{2,5},{0,1},{2,3},{4,5},{1,2},{3,4}};
  const int8_t (*net)[2] = np == 4 ? net4 : (np == 5 ? net5 : net6);
  const int nnet = np == 4 ? 5 : (np == 5 ? 9 : 12);
  for (int s = 0; s < nnet; s++) {
    int a = net[s][0], b = net[s][1];
    if (key[a] > key[b]) {
      float tk = key[a]; key[a] = key[b]; key[b] = tk;
      v3 tv = sp[a]; sp[a] = sp[b]; sp[b] = tv;
    }
  }
  for (int i = 0; i < np; i++) sp[i] = vmul(sp[i], w);
  float area = 0.0f;
  for (int i = 0; i + 1 < np; i++) area += vnorm(vcross(sp[i], sp[i + 1]));
  area += vnorm(vcross(sp[0], sp[np - 1]));
  return area * 0.5f;
}

static float voxel_area(uint64_t x, uint64_t y, uint64_t z, v3 pos, v3 n, const float inv[3], v3 w) {
  v3 pts[8];
  int np = plane_cube_points(1, x, y, z, pos, n, inv, pts);
  return polygon_area(pts, np, w, n);
}

/* xs3d.hpp:361-440 calc_area_at_point_2x2x2: (bx,by,bz) is the block's low (even) voxel */
static float block_area(uint8_t cube, uint64_t bx, uint64_t by, uint64_t bz, v3 pos, v3 n, const float inv[3], v3 w) {
  const float far2 = (float)(1.7320508076 + (double)XS_EPS);
  v3 centre = vadd(V((float)bx, (float)by, (float)bz), V(0.5f, 0.5f, 0.5f));
  if (fabsf(vdot(vsub(centre, pos), n)) > far2) return 0.0f;
  float area = 0.0f;
  int pop = __builtin_popcount(cube);
  if (pop >= 5) {
    v3 pts[8];
    int np = plane_cube_points(2, bx, by, bz, pos, n, inv, pts);
    area = polygon_area(pts, np, w, n);
    if (area == 0) return area;
    for (int b = 0; b < 8; b++)
      if (!((cube >> b) & 1)) area -= voxel_area(bx + (b & 1), by + ((b >> 1) & 1), bz + (b >> 2), pos, n, inv, w);
  } else {
    for (int b = 0; b < 8; b++)
      if ((cube >> b) & 1) area += voxel_area(bx + (b & 1), by + ((b >> 1) & 1), bz + (b >> 2), pos, n, inv, w);
  }
  return area;
}

typedef struct { const uint8_t* img; uint64_t sx, sy, sz; } vol_t;

static inline int fg(const vol_t* v, uint64_t x, uint64_t y, uint64_t z) { return v->img[x + v->sx * (y + v->sy * z)] != 0; }

/* xs3d.hpp:26-48 compute_cube: bit = dx + 2dy + 4dz, out-of-volume reads as 0 */
static uint8_t block_mask(const vol_t* v, uint64_t x, uint64_t y, uint64_t z) {
  uint8_t m = 0;
  for (int b = 0; b < 8; b++) {
    uint64_t xx = x + (b & 1), yy = y + ((b >> 1) & 1), zz = z + (b >> 2);
    if (xx < v->sx && yy < v->sy && zz < v->sz && fg(v, xx, yy, zz)) m |= (uint8_t)(1u << b);
  }
  return m;
}

/* xs3d.hpp:537-646 is_26_connected as a rule: the candidate must have a set voxel on its side
 * facing the centre block and the centre block one on its side facing the candidate.  One table
 * entry of the reference deviates from the rule (o = (-,+,0), centre mask 0b01000010, :569). */
static int blocks_adjacent(uint8_t centre, uint8_t cand, int ox, int oy, int oz) {
  if (ox == 0 && oy == 0 && oz == 0) return 1;
  uint8_t cm = 0, km = 0;
  for (int b = 0; b < 8; b++) {
    int d[3] = { b & 1, (b >> 1) & 1, b >> 2 }, o[3] = { ox, oy, oz };
    int on_c = 1, on_k = 1;
    for (int a = 0; a < 3; a++) {
      if (o[a] < 0) { on_c &= (d[a] == 0); on_k &= (d[a] == 1); }
      if (o[a] > 0) { on_c &= (d[a] == 1); on_k &= (d[a] == 0); }
    }
    if (on_c) cm |= (uint8_t)(1u << b);
    if (on_k) km |= (uint8_t)(1u << b);
  }
  if (ox < 0 && oy > 0 && oz == 0) cm = 0x42;
  return ((cand & km) != 0) & ((centre & cm) != 0);
}

/* ------------------------------------------------------------------ lattice on the plane */
typedef struct {
  v3 pos, n, w, b1, b2;
  float inv[3];
  float lim[3];          /* float(s) - 0.5, xs3d.hpp:866-868 */
  int64_t psx, c;        /* lattice side and centre index, xs3d.hpp:823,840 */
  int axis_aligned;
} plane_t;

static void plane_init(plane_t* P, const vol_t* v, v3 pos, v3 nraw, v3 w) {
  P->pos = pos; P->w = w;
  P->n = vdivs(nraw, vnorm(nraw));                           /* xs3d.hpp:1347-1349 */
  uint64_t m = v->sx > v->sy ? v->sx : v->sy; if (v->sz > m) m = v->sz;
  P->psx = (int64_t)(2 * 97 * m / 56 + 1);                   /* xs3d.hpp:823 */
  P->c = P->psx / 2;
  v3 b1 = vcross(P->n, V(1, 0, 0));                          /* xs3d.hpp:831-838 */
  if (b1.x == 0 && b1.y == 0 && b1.z == 0) b1 = vcross(P->n, V(0, 1, 0));
  b1 = vdivs(b1, vnorm(b1));
  v3 b2 = vcross(P->n, b1);
  b2 = vdivs(b2, vnorm(b2));
  P->b1 = b1; P->b2 = b2;
  for (int i = 0; i < 3; i++) { float p = comp(P->n, i); P->inv[i] = (p == 0) ? 0.0f : (float)(1.0 / (double)p); }
  P->lim[0] = (float)v->sx - 0.5f; P->lim[1] = (float)v->sy - 0.5f; P->lim[2] = (float)v->sz - 0.5f;
  P->axis_aligned = ((P->n.x != 0) + (P->n.y != 0) + (P->n.z != 0)) == 1;
}

/* sample point of lattice cell (i,j): (pos + b1*dx) + b2*dy, xs3d.hpp:883-886 */
static inline v3 cell_point(const plane_t* P, int64_t i, int64_t j) {
  float dx = (float)i - (float)P->c, dy = (float)j - (float)P->c;
  return vadd(vadd(P->pos, vscale(P->b1, dx)), vscale(P->b2, dy));
}

/* in bounds (xs3d.hpp:888-893) and foreground at the rounded point (:895-907).  A sample exactly
 * on -0.5 passes the reference's bounds test and then rounds to -1 (UB cast); treated as outside. */
static int cell_fg(const plane_t* P, const vol_t* v, int64_t i, int64_t j, v3* cur_out) {
  if (i < 0 || j < 0 || i >= P->psx || j >= P->psx) return 0;
  v3 cur = cell_point(P, i, j);
  if (cur.x < -0.5f || cur.y < -0.5f || cur.z < -0.5f) return 0;
  if (cur.x >= P->lim[0] || cur.y >= P->lim[1] || cur.z >= P->lim[2]) return 0;
  v3 r = vround(cur);
  if (r.x < 0 || r.y < 0 || r.z < 0) return 0;
  if (cur_out) *cur_out = cur;
  return fg(v, (uint64_t)r.x, (uint64_t)r.y, (uint64_t)r.z);
}

static int seed_ok(const vol_t* v, v3 pos) {   /* xs3d.hpp:1315-1345 */
  if (pos.x < 0 || pos.x >= (float)v->sx) return 0;
  if (pos.y < 0 || pos.y >= (float)v->sy) return 0;
  if (pos.z < 0 || pos.z >= (float)v->sz) return 0;
  v3 r = vround(pos);
  if (r.x < 0 || r.x >= (float)v->sx || r.y < 0 || r.y >= (float)v->sy || r.z < 0 || r.z >= (float)v->sz) return 0;
  return fg(v, (uint64_t)r.x, (uint64_t)r.y, (uint64_t)r.z);
}

/* The 8-connected component of the centre cell in DFS PREORDER, where the recursion tries the
 * neighbours in the reverse of the reference's push order (xs3d.hpp:926-950): (+1,+1) (-1,+1)
 * (+1,-1) (-1,-1) (0,+1) (0,-1) (+1,0) (-1,0).  state: 0 unknown, 1 fg-unvisited, 2 visited/bg. */
typedef struct { int64_t i, j; int next; } frame_t;
static const int DI[8] = { 1, -1, 1, -1, 0, 0, 1, -1 };
static const int DJ[8] = { 1, 1, -1, -1, 1, -1, 0, 0 };

static int64_t lattice_component(const plane_t* P, const vol_t* v, int64_t** order_i, int64_t** order_j) {
  const int64_t psx = P->psx;
  uint8_t* state = (uint8_t*)calloc((size_t)psx * (size_t)psx, 1);
  int64_t cap = 1024, n = 0, depth = 0, scap = 1024;
  int64_t* oi = (int64_t*)malloc(cap * sizeof(int64_t));
  int64_t* oj = (int64_t*)malloc(cap * sizeof(int64_t));
  frame_t* st = (frame_t*)malloc(scap * sizeof(frame_t));
  if (cell_fg(P, v, P->c, P->c, 0)) {
    state[P->c + psx * P->c] = 2;
    oi[n] = P->c; oj[n] = P->c; n++;
    st[depth].i = P->c; st[depth].j = P->c; st[depth].next = 0; depth++;
  }
  while (depth > 0) {
    frame_t* f = &st[depth - 1];
    if (f->next == 8) { depth--; continue; }
    int d = f->next++;
    int64_t i = f->i + DI[d], j = f->j + DJ[d];
    if (i < 0 || j < 0 || i >= psx || j >= psx) continue;
    uint8_t* s = &state[i + psx * j];
    if (*s == 0) *s = cell_fg(P, v, i, j, 0) ? 1 : 2;
    if (*s != 1) continue;
    *s = 2;
    if (n == cap) { cap *= 2; oi = (int64_t*)realloc(oi, cap * sizeof(int64_t)); oj = (int64_t*)realloc(oj, cap * sizeof(int64_t)); }
    oi[n] = i; oj[n] = j; n++;
    if (depth == scap) { scap *= 2; st = (frame_t*)realloc(st, scap * sizeof(frame_t)); }
    st[depth].i = i; st[depth].j = j; st[depth].next = 0; depth++;
  }
  free(st); free(state);
  *order_i = oi; *order_j = oj;
  return n;
}

/* ------------------------------------------------------------------ tiny open-addressing map u64 -> u64 (min) */
typedef struct { uint64_t* key; uint64_t* val; uint64_t mask, used; } map_t;
static void map_init(map_t* m, uint64_t cap) { uint64_t c = 64; while (c < cap) c <<= 1; m->key = (uint64_t*)malloc(c * 8); m->val = (uint64_t*)malloc(c * 8); memset(m->key, 0xff, c * 8); m->mask = c - 1; m->used = 0; }
static void map_free(map_t* m) { free(m->key); free(m->val); }
static uint64_t* map_slot(map_t* m, uint64_t k, int insert) {
  uint64_t h = (k * 0x9E3779B97F4A7C15ull) >> 20;
  for (;; h++) {
    uint64_t s = h & m->mask;
    if (m->key[s] == k) return &m->val[s];
    if (m->key[s] == ~0ull) { if (!insert) return 0; m->key[s] = k; m->val[s] = ~0ull; m->used++; return &m->val[s]; }
  }
}

/* ------------------------------------------------------------------ path A: xs3d.hpp:809-961 + :648-726 */
static void area_fast(const vol_t* v, v3 pos, v3 nraw, v3 w, float* area_out, uint8_t* contact_out) {
  *area_out = 0.0f; *contact_out = 0;
  if (v->sx == 0 || v->sy == 0 || v->sz == 0 || !seed_ok(v, pos)) return;
  plane_t P; plane_init(&P, v, pos, nraw, w);
  int64_t *oi, *oj;
  int64_t n = lattice_component(&P, v, &oi, &oj);
  const uint64_t hx = (v->sx + 1) >> 1, hy = (v->sy + 1) >> 1;

  /* pass 1: contact bits + first-touch ownership.  key = rank*32 + k, k = loop position of the
   * offset in the reference's z,y,x-ascending triple loop (xs3d.hpp:689-691). */
  map_t owner; map_init(&owner, (uint64_t)n * 27 * 2 + 64);
  uint8_t contact = 0;
  for (int64_t r = 0; r < n; r++) {
    v3 cur = vround(cell_point(&P, oi[r], oj[r]));
    contact |= (uint8_t)((cur.x < 1) | ((cur.x >= (float)((double)v->sx - 1.5)) << 1)      /* xs3d.hpp:909-914 */
             | ((cur.y < 1) << 2) | ((cur.y >= (float)((double)v->sy - 1.5)) << 3)
             | ((cur.z < 1) << 4) | ((cur.z >= (float)((double)v->sz - 1.5)) << 5));
    int64_t rx = (int64_t)cur.x, ry = (int64_t)cur.y, rz = (int64_t)cur.z;
    for (int k = 0; k < 27; k++) {
      int ox = (k % 3 - 1) * 2, oy = ((k / 3) % 3 - 1) * 2, oz = (k / 9 - 1) * 2;
      if (P.axis_aligned && (ox | oy | oz)) continue;                                    /* :677-685 */
      int64_t x = rx + ox, y = ry + oy, z = rz + oz;
      if (x < 0 || y < 0 || z < 0 || x >= (int64_t)v->sx || y >= (int64_t)v->sy || z >= (int64_t)v->sz) continue;  /* :665-671 */
      if (!fg(v, (uint64_t)x, (uint64_t)y, (uint64_t)z)) continue;                       /* :704 */
      uint64_t blk = (uint64_t)(x >> 1) + hx * ((uint64_t)(y >> 1) + hy * (uint64_t)(z >> 1));
      uint64_t* o = map_slot(&owner, blk, 1);
      uint64_t key = (uint64_t)r * 32 + (uint64_t)k;
      if (key < *o) *o = key;
    }
  }
  /* pass 2: each sample adds, in k order, the blocks it owns that pass the adjacency test; the
   * per-sample subtotals are added in rank order (two float32 levels: xs3d.hpp:663,713,848,952) */
  float total = 0.0f;
  for (int64_t r = 0; r < n; r++) {
    v3 cur = vround(cell_point(&P, oi[r], oj[r]));
    int64_t rx = (int64_t)cur.x, ry = (int64_t)cur.y, rz = (int64_t)cur.z;
    uint8_t centre = block_mask(v, (uint64_t)rx & ~1ull, (uint64_t)ry & ~1ull, (uint64_t)rz & ~1ull);
    float subtotal = 0.0f;
    for (int k = 0; k < 27; k++) {
      int ox = (k % 3 - 1) * 2, oy = ((k / 3) % 3 - 1) * 2, oz = (k / 9 - 1) * 2;
      if (P.axis_aligned && (ox | oy | oz)) continue;
      int64_t x = rx + ox, y = ry + oy, z = rz + oz;
      if (x < 0 || y < 0 || z < 0 || x >= (int64_t)v->sx || y >= (int64_t)v->sy || z >= (int64_t)v->sz) continue;
      uint64_t blk = (uint64_t)(x >> 1) + hx * ((uint64_t)(y >> 1) + hy * (uint64_t)(z >> 1));
      uint64_t* o = map_slot(&owner, blk, 0);
      if (!o || *o != (uint64_t)r * 32 + (uint64_t)k) continue;
      uint64_t bx = (uint64_t)x & ~1ull, by = (uint64_t)y & ~1ull, bz = (uint64_t)z & ~1ull;
      uint8_t cand = block_mask(v, bx, by, bz);
      if (blocks_adjacent(centre, cand, ox, oy, oz))
        subtotal += block_area(cand, bx, by, bz, P.pos, P.n, P.inv, P.w);
    }
    total += subtotal;
  }
  map_free(&owner); free(oi); free(oj);
  *area_out = total; *contact_out = contact;
}

/* auxiliary.hpp:193-266 cross_sectional_area_slow: every foreground voxel, double accumulation,
 * no seed-foreground test, contact = faces touched by any foreground voxel */
static void area_slow(const vol_t* v, v3 pos, v3 nraw, v3 w, float* area_out, uint8_t* contact_out) {
  *area_out = 0.0f; *contact_out = 0;
  v3 r = vround(pos);
  if (r.x < 0 || r.x >= (float)v->sx || r.y < 0 || r.y >= (float)v->sy || r.z < 0 || r.z >= (float)v->sz) return;
  v3 n = vdivs(nraw, vnorm(nraw));
  float inv[3];
  for (int i = 0; i < 3; i++) { float p = comp(n, i); inv[i] = (p == 0) ? 0.0f : (float)(1.0 / (double)p); }
  double acc = 0; uint8_t contact = 0;
  for (uint64_t z = 0; z < v->sz; z++) for (uint64_t y = 0; y < v->sy; y++) for (uint64_t x = 0; x < v->sx; x++) {
    if (!fg(v, x, y, z)) continue;
    contact |= (uint8_t)((x < 1) | ((x >= v->sx - 1) << 1) | ((y < 1) << 2) | ((y >= v->sy - 1) << 3) | ((z < 1) << 4) | ((z >= v->sz - 1) << 5));
    acc += voxel_area(x, y, z, pos, n, inv, w);
  }
  *area_out = (float)acc; *contact_out = contact;
}

void oracle_area(const uint8_t* img, uint64_t sx, uint64_t sy, uint64_t sz, const float* p, const float* n, const float* w,
                 int slow_method, int use_persistent_data, float* area, uint8_t* contact) {
  (void)use_persistent_data;   /* xs3d.hpp:963-1115: same results as the non-persistent path */
  vol_t v = { img, sx, sy, sz };
  if (slow_method) area_slow(&v, V(p[0], p[1], p[2]), V(n[0], n[1], n[2]), V(w[0], w[1], w[2]), area, contact);
  else area_fast(&v, V(p[0], p[1], p[2]), V(n[0], n[1], n[2]), V(w[0], w[1], w[2]), area, contact);
}

void oracle_area_batch(const uint8_t* img, uint64_t sx, uint64_t sy, uint64_t sz, uint64_t nq, const float* pos,
                       const float* nrm, const float* w, float* area, uint8_t* contact, int nthreads) {
  (void)nthreads;
  for (uint64_t i = 0; i < nq; i++) oracle_area(img, sx, sy, sz, pos + 3 * i, nrm + 3 * i, w, 0, 0, area + i, contact + i);
}

/* ------------------------------------------------------------------ path B: xs3d.hpp:1118-1268 + :442-535 */
static void section_fast(const vol_t* v, v3 pos, v3 nraw, v3 w, float* out, uint8_t* contact_out) {
  *contact_out = 0;
  if (v->sx == 0 || v->sy == 0 || v->sz == 0 || !seed_ok(v, pos)) return;
  plane_t P; plane_init(&P, v, pos, nraw, w);
  int64_t *oi, *oj;
  int64_t n = lattice_component(&P, v, &oi, &oj);
  map_t seen; map_init(&seen, (uint64_t)n * 27 * 2 + 64);
  uint8_t contact = 0;
  const float fsx = (float)v->sx, fsy = (float)v->sy, fsz = (float)v->sz;
  for (int64_t r = 0; r < n; r++) {
    v3 cur = cell_point(&P, oi[r], oj[r]);     /* NOT rounded on this path */
    contact |= (uint8_t)((cur.x < 0.5f) | ((cur.x >= (float)((double)v->sx - 1.5)) << 1)   /* xs3d.hpp:1215-1220 */
             | ((cur.y < 0.5f) << 2) | ((cur.y >= (float)((double)v->sy - 1.5)) << 3)
             | ((cur.z < 0.5f) << 4) | ((cur.z >= (float)((double)v->sz - 1.5)) << 5));
    /* neighbour range from the UNROUNDED sample (xs3d.hpp:458-478) */
    int lo[3] = { (cur.x - 1) >= 0 ? -1 : 0, (cur.y - 1) >= 0 ? -1 : 0, (cur.z - 1) >= 0 ? -1 : 0 };
    int hi[3] = { (cur.x + 1) < fsx ? 1 : 0, (cur.y + 1) < fsy ? 1 : 0, (cur.z + 1) < fsz ? 1 : 0 };
    if (P.axis_aligned) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0; }
    for (int dz = lo[2]; dz <= hi[2]; dz++) for (int dy = lo[1]; dy <= hi[1]; dy++) for (int dx = lo[0]; dx <= hi[0]; dx++) {
      v3 q = vround(vadd(V((float)dx, (float)dy, (float)dz), cur));    /* round(cur + d), :484-490 */
      if (q.x < 0 || q.x >= fsx || q.y < 0 || q.y >= fsy || q.z < 0 || q.z >= fsz) continue;
      uint64_t x = (uint64_t)q.x, y = (uint64_t)q.y, z = (uint64_t)q.z;
      if (!fg(v, x, y, z)) continue;
      uint64_t loc = x + v->sx * (y + v->sy * z);
      uint64_t* s = map_slot(&seen, loc, 1);
      if (*s != ~0ull) continue;           /* first touch only (:511-512) */
      *s = 1;
      float a = voxel_area(x, y, z, P.pos, P.n, P.inv, P.w);
      if (a > 0.0f) out[loc] = a;
    }
  }
  map_free(&seen); free(oi); free(oj);
  *contact_out = contact;
}

/* auxiliary.hpp:94-181 cross_section_slow (method 1) */
static void section_slow(const vol_t* v, v3 pos, v3 nraw, v3 w, float* out, uint8_t* contact_out) {
  *contact_out = 0;
  if (pos.x < 0 || pos.x >= (float)v->sx || pos.y < 0 || pos.y >= (float)v->sy || pos.z < 0 || pos.z >= (float)v->sz) return;
  v3 r = vround(pos);
  if (r.x < 0 || r.x >= (float)v->sx || r.y < 0 || r.y >= (float)v->sy || r.z < 0 || r.z >= (float)v->sz) return;
  v3 n = vdivs(nraw, vnorm(nraw));
  float inv[3];
  for (int i = 0; i < 3; i++) { float p = comp(n, i); inv[i] = (p == 0) ? 0.0f : (float)(1.0 / (double)p); }
  uint8_t contact = 0;
  for (uint64_t z = 0; z < v->sz; z++) for (uint64_t y = 0; y < v->sy; y++) for (uint64_t x = 0; x < v->sx; x++) {
    if (!fg(v, x, y, z)) continue;
    contact |= (uint8_t)((x < 1) | ((x >= v->sx - 1) << 1) | ((y < 1) << 2) | ((y >= v->sy - 1) << 3) | ((z < 1) << 4) | ((z >= v->sz - 1) << 5));
    out[x + v->sx * (y + v->sy * z)] = voxel_area(x, y, z, pos, n, inv, w);
  }
  *contact_out = contact;
}

/* auxiliary.hpp:8-90 cross_section_slow_2x2x2 (method 2): half-resolution, block polygons only */
static void section_blocks(const vol_t* v, v3 pos, v3 nraw, v3 w, float* out, uint8_t* contact_out) {
  *contact_out = 0;
  if (pos.x < 0 || pos.x >= (float)v->sx || pos.y < 0 || pos.y >= (float)v->sy || pos.z < 0 || pos.z >= (float)v->sz) return;
  v3 r = vround(pos);
  if (r.x < 0 || r.x >= (float)v->sx || r.y < 0 || r.y >= (float)v->sy || r.z < 0 || r.z >= (float)v->sz) return;
  v3 n = vdivs(nraw, vnorm(nraw));
  float inv[3];
  for (int i = 0; i < 3; i++) { float p = comp(n, i); inv[i] = (p == 0) ? 0.0f : (float)(1.0 / (double)p); }
  const uint64_t hx = (v->sx + 1) >> 1, hy = (v->sy + 1) >> 1;
  for (uint64_t z = 0; z < v->sz; z += 2) for (uint64_t y = 0; y < v->sy; y += 2) for (uint64_t x = 0; x < v->sx; x += 2) {
    if (!fg(v, x, y, z)) continue;
    v3 pts[8];
    int np = plane_cube_points(2, x, y, z, pos, n, inv, pts);
    out[(x >> 1) + hx * ((y >> 1) + hy * (z >> 1))] = polygon_area(pts, np, w, n);
  }
}

void oracle_section(const uint8_t* img, uint64_t sx, uint64_t sy, uint64_t sz, const float* p, const float* n, const float* w,
                    int method, float* out, uint8_t* contact) {
  vol_t v = { img, sx, sy, sz };
  v3 P = V(p[0], p[1], p[2]), N = V(n[0], n[1], n[2]), W = V(w[0], w[1], w[2]);
  if (method == 0) section_fast(&v, P, N, W, out, contact);
  else if (method == 1) section_slow(&v, P, N, W, out, contact);
  else section_blocks(&v, P, N, W, out, contact);
}

/* ------------------------------------------------------------------ slice: xs3d.hpp:1428-1624 */
static int argmax_abs(v3 a) {   /* vec.hpp:102 on abs() */
  float x = fabsf(a.x), y = fabsf(a.y), z = fabsf(a.z);
  if (x >= y) return (x >= z) ? 0 : 2;
  return (y >= z) ? 1 : 2;
}

void oracle_slice_basis(const float* nin, int positive, float* o1, float* o2) {   /* xs3d.hpp:1428-1471; n already unit */
  v3 n = V(nin[0], nin[1], nin[2]);
  v3 b1 = vcross(n, V(0, 1, 0));
  if (b1.x == 0 && b1.y == 0 && b1.z == 0) b1 = vcross(n, V(1, 0, 0));
  b1 = vdivs(b1, vnorm(b1));
  v3 b2 = vcross(n, b1);
  b2 = vdivs(b2, vnorm(b2));
  if (argmax_abs(b2) < argmax_abs(b1)) { v3 t = b1; b1 = b2; b2 = t; }
  if (positive) {
    v3 zone = V(1, 1, 1);
    if (vdot(n, zone) < 0) zone = V(-1, -1, -1);
    if (vdot(b1, zone) < 0) b1 = V(-b1.x, -b1.y, -b1.z);
    if (vdot(b2, zone) < 0) b2 = V(-b2.x, -b2.y, -b2.z);
  }
  o1[0] = b1.x; o1[1] = b1.y; o1[2] = b1.z; o2[0] = b2.x; o2[1] = b2.y; o2[2] = b2.z;
}

static int64_t g_clamped = 0;
int64_t oracle_projection_clamped(void) { int64_t c = g_clamped; g_clamped = 0; return c; }

static inline float fmin3(v3 a) { float m = a.x < a.y ? a.x : a.y; return m < a.z ? m : a.z; }

/* `out` holds psx*psx labels (zeroed by the caller), psx as computed by fastxs3d.cpp:137-154;
 * bbox = {x_min,x_max,y_min,y_max} before the caller's +1 (fastxs3d.cpp:176-177). */
int oracle_projection(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int itemsize, int c_order,
                      const float* p, const float* nin, const float* win, int positive_basis, float crop,
                      void* out, int64_t* bbox) {
  if (itemsize != 1 && itemsize != 2 && itemsize != 4 && itemsize != 8) return -1;
  v3 w = V(win[0], win[1], win[2]);
  const float wmin = fmin3(w);
  const float crop_sq = crop * crop / wmin / wmin;                       /* xs3d.hpp:1489 */
  w = vdivs(w, wmin);
  float wmax = fabsf(w.x) > fabsf(w.y) ? fabsf(w.x) : fabsf(w.y); if (fabsf(w.z) > wmax) wmax = fabsf(w.z);
  const uint64_t distortion = (uint64_t)ceil((double)wmax);
  v3 scale = V(1.0f / w.x, 1.0f / w.y, 1.0f / w.z);                      /* :1496 */
  uint64_t largest = sx > sy ? sx : sy; if (sz > largest) largest = sz;
  if ((float)largest > crop && crop >= 0) largest = (uint64_t)ceilf(crop);
  const int64_t psx = (int64_t)(distortion * 2 * 97 * largest / 56) + 1;
  bbox[0] = bbox[1] = bbox[2] = bbox[3] = 0;
  if (p[0] < 0 || p[0] >= (float)sx || p[1] < 0 || p[1] >= (float)sy || p[2] < 0 || p[2] >= (float)sz) return 0;
  v3 pos = V(p[0], p[1], p[2]);
  v3 n = V(nin[0], nin[1], nin[2]);
  n = vdivs(n, vnorm(n));
  float nb[3] = { n.x, n.y, n.z }, f1[3], f2[3];
  oracle_slice_basis(nb, positive_basis, f1, f2);
  v3 b1 = vmul(V(f1[0], f1[1], f1[2]), scale), b2 = vmul(V(f2[0], f2[1], f2[2]), scale);
  const int64_t c = psx / 2;
  bbox[0] = bbox[1] = bbox[2] = bbox[3] = c;
  if (!(crop > 0)) return 0;
  const float lx = (float)sx - 0.5f, ly = (float)sy - 0.5f, lz = (float)sz - 0.5f;
  uint8_t* seen = (uint8_t*)calloc((size_t)psx * (size_t)psx, 1);
  int64_t cap = 4096, top = 0;
  int64_t* st = (int64_t*)malloc(cap * sizeof(int64_t));
  st[top++] = c + psx * c;
  while (top > 0) {                                                       /* 4-connected fill, :1554-1621 */
    int64_t loc = st[--top];
    if (seen[loc]) continue;
    seen[loc] = 1;
    int64_t y = loc / psx, x = loc - y * psx;
    float dx = (float)x - (float)c, dy = (float)y - (float)c;
    v3 delta = vadd(vscale(b1, dx), vscale(b2, dy));
    v3 cur = vadd(pos, delta);
    if (cur.x < -0.5f || cur.y < -0.5f || cur.z < -0.5f) continue;
    if (cur.x >= lx || cur.y >= ly || cur.z >= lz) continue;
    if (vnorm2(delta) >= crop_sq) continue;
    v3 r = vround(cur);
    /* A coordinate of exactly -0.5 passes the reference's bounds test and rounds to -1, whose
     * uint64 cast is UB (the reference then gathers from a wrapped index).  The cell still counts
     * towards the bbox (so the output SHAPE matches the reference); its value is defined here as
     * the label of the clamped voxel.  oracle_projection_clamped() lets tests mask such cases. */
    if (r.x < 0) { r.x = 0; g_clamped++; }
    if (r.y < 0) { r.y = 0; g_clamped++; }
    if (r.z < 0) { r.z = 0; g_clamped++; }
    if (x < bbox[0]) bbox[0] = x; if (x > bbox[1]) bbox[1] = x;
    if (y < bbox[2]) bbox[2] = y; if (y > bbox[3]) bbox[3] = y;
    uint64_t vx = (uint64_t)r.x, vy = (uint64_t)r.y, vz = (uint64_t)r.z;
    uint64_t src = c_order ? (vz + sz * (vy + sy * vx)) : (vx + sx * (vy + sy * vz));
    memcpy((char*)out + (size_t)loc * itemsize, (const char*)labels + (size_t)src * itemsize, (size_t)itemsize);
    if (top + 4 > cap) { cap *= 2; st = (int64_t*)realloc(st, cap * sizeof(int64_t)); }
    if (x > 0 && !seen[loc - 1]) st[top++] = loc - 1;
    if (x < psx - 1 && !seen[loc + 1]) st[top++] = loc + 1;
    if (y > 0 && !seen[loc - psx]) st[top++] = loc - psx;
    if (y < psx - 1 && !seen[loc + psx]) st[top++] = loc + psx;
  }
  free(st); free(seen);
  return 0;
}

/* kernel-level helpers exported for unit goldens */
float oracle_voxel_area(uint64_t x, uint64_t y, uint64_t z, const float* p, const float* n, const float* w) {
  v3 N = V(n[0], n[1], n[2]);
  float inv[3];
  for (int i = 0; i < 3; i++) { float q = comp(N, i); inv[i] = (q == 0) ? 0.0f : (float)(1.0 / (double)q); }
  return voxel_area(x, y, z, V(p[0], p[1], p[2]), N, inv, V(w[0], w[1], w[2]));
}
float oracle_block_area(uint8_t cube, const float* cur, const float* p, const float* n, const float* w) {
  v3 N = V(n[0], n[1], n[2]);
  float inv[3];
  for (int i = 0; i < 3; i++) { float q = comp(N, i); inv[i] = (q == 0) ? 0.0f : (float)(1.0 / (double)q); }
  return block_area(cube, (uint64_t)cur[0] & ~1ull, (uint64_t)cur[1] & ~1ull, (uint64_t)cur[2] & ~1ull,
                    V(p[0], p[1], p[2]), N, inv, V(w[0], w[1], w[2]));
}
int oracle_is_26_connected(uint8_t centre, uint8_t cand, int x, int y, int z) { return blocks_adjacent(centre, cand, x, y, z); }
