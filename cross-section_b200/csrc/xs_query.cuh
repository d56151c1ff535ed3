// xs_query.cuh — one cross-section query, cooperatively executed by a thread group.
//
// Replaces (file:line in /root/reference/src): cross_sectional_area xs3d.hpp:1305-1365,
// cross_sectional_area_helper_2x2x2 :809-961, robust_calc_area_at_point_2x2x2 :648-726,
// cross_section :1367-1426, cross_sectional_area_helper :1118-1268, calc_area_at_point :442-535.
//
// The reference fuses everything into one sequential flood fill.  Here the query is decomposed
//  so that only a bitmap walk is sequential, and the floating-point work is split off
// into a separate, perfectly regular kernel:
//   QUERY KERNELS (one warp or one CTA per query; this file, run_area_emit / run_section_emit)
//   1. lattice tiles   : the plane lattice (basis b1,b2, unit spacing, centred on the vertex) is
//                        sampled in 32x32-cell tiles ON DEMAND into a foreground bitmap (one ballot
//                        per lattice row) — work scales with the section, not with the volume;
//   2. component + rank: one warp, in lock step, walks depth-first from the centre cell with the
//                        reference's neighbour priority; the walk both delimits the 8-connected
//                        component and yields its DFS preorder (`order[]`).  Two forms: over the tile
//                        bitmap (dfs_run: warp tier, big tier) and over one neighbour byte per cell with
//                        the tiles sampled up front (flood_component_nbr: mid and wide tiers);
//   3. claims          : every sample claims the <=27 neighbouring 2x2x2 blocks whose probe voxel
//                        is set: atomicMin of its rank into a hash table keyed by block —
//                        "first touch wins" without the sequential visit;
//   4. emit            : owners that pass the adjacency and plane-distance tests become ITEMS, in
//                        (rank,k) order; each item lists the plane∩cube POLYGONS it needs (the block
//                        polygon and/or unit-voxel polygons) as 12-byte records in HBM.
//   POLYGON KERNEL (records filtered and compacted per warp, then one thread per polygon; poly_eval)
//   5. evaluate        : intersect the 12 cube edges with the plane, order the vertices, area.
//   COMBINE KERNEL (one CTA per query; combine_query)
//   6. ordered sum     : block value from its polygons, float32 sum per sample in k order, then
//                        over samples in rank order — the reference's exact summation order.
// Path B (cross_section) is order-free and runs as grid-wide kernels of its own (xs_section.cu); it shares the
// lattice sampling (cell_probe), the window (bbox_window) and the polygon kernel.
//
// Volume layout ("cubes"): one byte per 2x2x2 block = the reference's compute_cube() mask
// (xs3d.hpp:26-48), bit = dx + 2dy + 4dz, out-of-volume voxels 0; x fastest.  A voxel test and a
// block-mask fetch are both ONE byte load, and a 512^3 volume is 16 MiB — L2 resident.
#pragma once
#include "xs_geom.cuh"

namespace xs {

struct VolView {
  const uint8_t* cubes;
  int sx, sy, sz;   // voxels
  int hx, hy, hz;   // blocks = ceil(s/2)
};

// block index arithmetic is 32-bit: volumes are limited to < 2^32 blocks (checked at xs3d_volume_create)
XS_HD uint32_t cube_index(const VolView& v, int bx, int by, int bz) {
  return (uint32_t)bx + (uint32_t)v.hx * ((uint32_t)by + (uint32_t)v.hy * (uint32_t)bz);
}
XS_HD uint32_t cube_load(const VolView& v, uint32_t idx) {
#ifdef __CUDA_ARCH__
  return __ldg(v.cubes + idx);
#else
  return v.cubes[idx];
#endif
}
XS_HD uint32_t cube_at(const VolView& v, int bx, int by, int bz) { return cube_load(v, cube_index(v, bx, by, bz)); }
// k = (ox+1) + 3(oy+1) + 9(oz+1)  ->  offsets, without integer division
XS_HD void decode_k(int k, int& ox, int& oy, int& oz) {
  const int z = (k * 57) >> 9;          // k / 9 for k < 27
  const int rem = k - 9 * z;
  const int y = (rem * 11) >> 5;        // rem / 3 for rem < 9
  ox = rem - 3 * y - 1; oy = y - 1; oz = z - 1;
}
XS_HD bool vox_fg(const VolView& v, int x, int y, int z) {
  const uint32_t c = cube_at(v, x >> 1, y >> 1, z >> 1);
  return (c >> ((x & 1) | ((y & 1) << 1) | ((z & 1) << 2))) & 1u;
}

struct PlaneCtx {
  PlaneGeom g;
  V3 b1, b2;        // lattice basis (xs3d.hpp:831-838)
  float lim[3];     // float(s) - 0.5           (xs3d.hpp:866-868)
  float hi[3];      // s - 1.5 (exact in float) (xs3d.hpp:910,912,914)
  int psx, c;       // lattice side, centre index (xs3d.hpp:823,840)
  int axis_aligned;
};

// Seed validation + per-query constants.  Returns false when the reference returns (0, 0)
// (xs3d.hpp:1315-1345): vertex outside the volume or on background.
// check_seed = false: skip the read of the seed voxel (host callers: the occupancy bytes live in device memory; the
// centre lattice cell, whose sample point IS the vertex, is then tested on the device).
XS_HD bool plane_init(PlaneCtx& c, const VolView& v, V3 pos, V3 nraw, V3 w, bool check_seed = true) {
  if (v.sx <= 0 || v.sy <= 0 || v.sz <= 0) return false;
  if (!(pos.x >= 0.0f) || pos.x >= (float)v.sx) return false;
  if (!(pos.y >= 0.0f) || pos.y >= (float)v.sy) return false;
  if (!(pos.z >= 0.0f) || pos.z >= (float)v.sz) return false;
  const V3 r = vround(pos);
  if (r.x >= (float)v.sx || r.y >= (float)v.sy || r.z >= (float)v.sz) return false;
  if (check_seed && !vox_fg(v, (int)r.x, (int)r.y, (int)r.z)) return false;
  const V3 n = vdivs(nraw, vnorm(nraw));
  geom_init(c.g, pos, n, w);
  int m = v.sx > v.sy ? v.sx : v.sy;
  if (v.sz > m) m = v.sz;
  c.psx = (int)((2ll * 97ll * (long long)m) / 56ll + 1ll);
  c.c = c.psx / 2;
  V3 b1 = vcross(n, mk(1.0f, 0.0f, 0.0f));
  if (vnull(b1)) b1 = vcross(n, mk(0.0f, 1.0f, 0.0f));
  b1 = vdivs(b1, vnorm(b1));
  V3 b2 = vcross(n, b1);
  b2 = vdivs(b2, vnorm(b2));
  c.b1 = b1; c.b2 = b2;
  c.lim[0] = fsub((float)v.sx, 0.5f); c.lim[1] = fsub((float)v.sy, 0.5f); c.lim[2] = fsub((float)v.sz, 0.5f);
  c.hi[0] = (float)v.sx - 1.5f; c.hi[1] = (float)v.sy - 1.5f; c.hi[2] = (float)v.sz - 1.5f;
  c.axis_aligned = (((n.x != 0.0f) + (n.y != 0.0f) + (n.z != 0.0f)) == 1) ? 1 : 0;
  return true;
}

// Sample point of lattice cell (i,j): (pos + b1*dx) + b2*dy with dx = float(i) - float(c) (xs3d.hpp:883-886)
XS_HD V3 cell_point(const PlaneCtx& c, int i, int j) {
  const float dx = fsub((float)i, (float)c.c), dy = fsub((float)j, (float)c.c);
  return vadd(vadd(c.g.pos, vscale(c.b1, dx)), vscale(c.b2, dy));
}

// Cell is inside the volume (xs3d.hpp:888-893) and its rounded voxel is foreground (:895-907).
// A coordinate of exactly -0.5 passes the reference's test and then rounds to -1 (UB cast);
// such samples are treated as outside (documented divergence).
// Split in two so that several cells' volume loads can be in flight at once: cell_probe says where to look (block
// index and bit within its occupancy byte), or false when the cell is outside the lattice / the volume.
XS_HD bool cell_probe(const PlaneCtx& c, const VolView& v, int i, int j, uint32_t* cube, uint32_t* bit) {
  *cube = 0u; *bit = 0u;
  if (i < 0 || j < 0 || i >= c.psx || j >= c.psx) return false;
  const V3 p = cell_point(c, i, j);
  if (p.x < -0.5f || p.y < -0.5f || p.z < -0.5f) return false;
  if (p.x >= c.lim[0] || p.y >= c.lim[1] || p.z >= c.lim[2]) return false;
  if (!(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z)) return false;
  const V3 r = vround(p);
  if (r.x < 0.0f || r.y < 0.0f || r.z < 0.0f) return false;
  const int x = (int)r.x, y = (int)r.y, z = (int)r.z;
  *cube = cube_index(v, x >> 1, y >> 1, z >> 1);
  *bit = (uint32_t)((x & 1) | ((y & 1) << 1) | ((z & 1) << 2));
  return true;
}

XS_HD bool cell_fg(const PlaneCtx& c, const VolView& v, int i, int j) {
  uint32_t cube, bit;
  if (!cell_probe(c, v, i, j, &cube, &bit)) return false;
  return (cube_load(v, cube) >> bit) & 1u;
}

// ------------------------------------------------------------------------------------------------
// Thread-group policies.  WarpGroup: one warp per query (T0).  BlockGroup: one CTA per query (T1).
// HostGroup: a single host thread, test builds only.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
struct WarpGroup {
  int lane_;
  __device__ WarpGroup() : lane_(threadIdx.x & 31) {}
  __device__ int tid() const { return lane_; }
  __device__ int size() const { return 32; }
  __device__ int lane() const { return lane_; }
  __device__ int warp() const { return 0; }
  __device__ int nwarps() const { return 1; }
  __device__ void sync() const { __syncwarp(); }
  __device__ void converge() const { __syncwarp(); }
  __device__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ int exscan(int v, int& total) const {
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane_ >= d) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
  }
  __device__ uint32_t or_reduce(uint32_t v) const { return __reduce_or_sync(0xffffffffu, v); }
  __device__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ int wsize() const { return 32; }
  __device__ int wexscan(int v, int& total) const { return exscan(v, total); }
};

struct BlockGroup {
  int* scratch;   // >= 36 ints of shared memory
  __device__ explicit BlockGroup(int* s) : scratch(s) {}
  __device__ int tid() const { return threadIdx.x; }
  __device__ int size() const { return blockDim.x; }
  __device__ int lane() const { return threadIdx.x & 31; }
  __device__ int warp() const { return threadIdx.x >> 5; }
  __device__ int nwarps() const { return blockDim.x >> 5; }
  __device__ void sync() const { __syncthreads(); }
  __device__ void converge() const { __syncwarp(); }   // re-join the lanes of each warp after divergent per-lane loops
  __device__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ int exscan(int v, int& total) const {
    const int l = lane(), w = warp(), nw = nwarps();
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (l >= d) inc += t;
    }
    if (l == 31) scratch[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int i = 0; i < nw; i++) {
      int s = scratch[i];
      if (i < w) base += s;
      tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
  }
  __device__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  __device__ int wsize() const { return 32; }
  __device__ int wexscan(int v, int& total) const {   // exclusive scan within the calling warp
    const int l = lane();
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (l >= d) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
  }
  __device__ uint32_t or_reduce(uint32_t v) const {
    v = __reduce_or_sync(0xffffffffu, v);
    if (threadIdx.x == 0) scratch[34] = 0;
    __syncthreads();
    if (lane() == 0 && v) atomicOr((unsigned int*)&scratch[34], v);
    __syncthreads();
    uint32_t r = (uint32_t)scratch[34];
    __syncthreads();
    return r;
  }
};
#endif

struct HostGroup {
  int tid() const { return 0; }
  int size() const { return 1; }
  int lane() const { return 0; }
  int warp() const { return 0; }
  int nwarps() const { return 1; }
  void sync() const {}
  void converge() const {}
  int exscan(int v, int& total) const { total = v; return 0; }
  uint32_t or_reduce(uint32_t v) const { return v; }
  int shfl(int v, int) const { return v; }
  int wsize() const { return 1; }
  int wexscan(int v, int& total) const { total = v; return 0; }
};

XS_HD uint32_t atomic_cas_u32(uint32_t* p, uint32_t cmp, uint32_t val) {
#ifdef __CUDA_ARCH__
  return atomicCAS(p, cmp, val);
#else
  uint32_t old = *p;
  if (old == cmp) *p = val;
  return old;
#endif
}
XS_HD uint32_t atomic_min_u32(uint32_t* p, uint32_t val) {   // returns the previous value
#ifdef __CUDA_ARCH__
  return atomicMin(p, val);
#else
  const uint32_t old = *p;
  if (val < old) *p = val;
  return old;
#endif
}
XS_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t val) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, val);
#else
  uint32_t old = *p; *p = old + val; return old;
#endif
}
XS_HD void atomic_or_u32(uint32_t* p, uint32_t val) {
#ifdef __CUDA_ARCH__
  atomicOr(p, val);
#else
  *p |= val;
#endif
}

// ------------------------------------------------------------------------------------------------
// Records exchanged between the query kernels, the polygon kernel and the combine kernel (HBM).
// ------------------------------------------------------------------------------------------------
struct PolyRec {            // 12 bytes: one plane∩cube polygon to evaluate
  uint32_t q;               // query index (-> GeomRec)
  uint16_t x, y, z;         // low voxel of the cube
  uint16_t kind;            // 1: unit voxel, 2: 2x2x2 block
};
struct ItemRec {            // 8 bytes: one owned block that contributes to the area
  uint32_t poly;            // absolute index of its first polygon record / value
  uint32_t meta;            // bits 0-3 polygon count, bit 4 subtract mode (>=5 voxels set), bit 5 first item of its sample
};
struct QInfo {              // 16 bytes per query
  uint32_t item_off, n_items;
  uint32_t contact;
  uint32_t status;          // QS_*
};
enum { QS_PENDING = 0, QS_READY = 1, QS_READY2 = 2 };   // READY2: the records live in the second pipeline's buffers
struct EmitOut {
  PolyRec* polys;
  ItemRec* items;
  PlaneGeom* geom;          // per query
  QInfo* qinfo;             // per query
  uint32_t* counters;       // [0] polygons used, [1] items used, [2] capacity overflow flag
  uint32_t poly_cap, item_cap;
  uint32_t ready;           // QInfo.status to publish (QS_READY / QS_READY2)
};

// ------------------------------------------------------------------------------------------------
// Per-query working set.  Pointers may address shared memory (warp tier, and the CTA tier's bitmap /
// stack ring) or a per-CTA slab in global memory (CTA tier).  CellT packs a window cell (u,v):
// uint16_t (u | v<<8) for the warp tier's 96x96 window, uint32_t (u | v<<16) otherwise.
// ------------------------------------------------------------------------------------------------
template <class CellT> struct CellPack;
template <> struct CellPack<uint16_t> {
  static XS_HD uint16_t pack(int u, int v) { return (uint16_t)(u | (v << 8)); }
  static XS_HD int u(uint16_t p) { return p & 0xff; }
  static XS_HD int v(uint16_t p) { return p >> 8; }
};
template <> struct CellPack<uint32_t> {
  static XS_HD uint32_t pack(int u, int v) { return (uint32_t)u | ((uint32_t)v << 16); }
  static XS_HD int u(uint32_t p) { return (int)(p & 0xffffu); }
  static XS_HD int v(uint32_t p) { return (int)(p >> 16); }
};

#define XS_HEMPTY 0xffffffffu

// First-touch table, packed form (warp tier): one 32-bit word per slot = tag<<9 | rank, where tag is the
// block position relative to the query's centre block (7 bits per axis) and rank < 512.  Because all
// claims on a slot that already holds the same tag share the upper bits, atomicMin on the whole word is
// atomicMin on the rank.
struct PackedHash {
  static constexpr bool kBatch = false;
  XS_HD void fit(int) {}
  uint32_t* slot;
  uint32_t cap;        // power of two
  int shift;           // 32 - log2(cap)
  int cbx, cby, cbz;   // centre block of the query
  int max_probe;
  XS_HD uint32_t tag(int bx, int by, int bz) const {
    return (uint32_t)(bx - cbx + 64) | ((uint32_t)(by - cby + 64) << 7) | ((uint32_t)(bz - cbz + 64) << 14);
  }
  template <class G> XS_D void clear(const G& g) const {
    for (uint32_t i = g.tid(); i < cap; i += g.size()) slot[i] = XS_HEMPTY;
  }
  // returns 0: table full (probe limit), 1: claimed but a smaller rank holds the block, 2: this rank holds it NOW
  // (it may still lose it to a smaller rank claiming later — "tentative owner")
  // tag of block b + o = tag of b + tag_delta(o)  (no carries between the 7-bit fields: blocks stay inside the window)
  XS_HD uint32_t tag_delta(int ox, int oy, int oz) const { return (uint32_t)(ox + (oy << 7) + (oz << 14)); }
  XS_HD int claim(int bx, int by, int bz, uint32_t rank) const { return claim_tag(tag(bx, by, bz), rank); }
  XS_HD int claim_tag(uint32_t t, uint32_t rank) const {
    const uint32_t word = (t << 9) | rank;
    uint32_t h = (t * 2654435761u) >> shift;
    for (int probes = 0; probes < max_probe; probes++) {
      uint32_t old = slot[h];
      if (old == XS_HEMPTY) old = atomic_cas_u32(&slot[h], XS_HEMPTY, word);
      if (old == XS_HEMPTY) return 2;
      if ((old >> 9) == t) {
        if (old > word) old = atomic_min_u32(&slot[h], word);
        return old > word ? 2 : 1;
      }
      h = (h + 1u) & (cap - 1u);
    }
    return 0;
  }
  XS_HD bool batched() const { return false; }     // shared-memory table: claims go offset by offset (claim_blocks)
  XS_HD uint32_t claim9(uint32_t, uint32_t, uint32_t, bool*) const { return 0u; }
  XS_HD uint32_t owner(int bx, int by, int bz) const {
    const uint32_t t = tag(bx, by, bz);
    uint32_t h = (t * 2654435761u) >> shift;
    for (;;) {
      const uint32_t e = slot[h];
      if ((e >> 9) == t) return e & 0x1ffu;
      h = (h + 1u) & (cap - 1u);
    }
  }
};

// First-touch table, wide form (CTA tier): separate tag (linear block id) and rank arrays in HBM.
// kBatch: compile the batched claim path (claim9) in — for the tiers whose table can be in HBM and large
template <bool BATCH>
struct WideHashT {
  static constexpr bool kBatch = BATCH;
  uint32_t* htag;
  uint32_t* hkey;
  uint32_t cap;        // capacity in use (power of two)
  int shift;           // 32 - log2(cap)
  int hx, hy;
  int max_probe;
  uint32_t cap_max;    // allocated capacity; with factor > 0 the capacity in use is re-chosen per query
  int factor;          //   as the power of two >= factor * samples (<= cap_max), so small queries clear little
  uint32_t* alt_tag;   // optional small table in SHARED memory (the mid tier's cell window, dead after the walk):
  uint32_t* alt_key;   //   used instead of the HBM table whenever 3 * samples fit in it
  uint32_t alt_cap;
  XS_HD void fit(int n) {
    if (alt_tag != nullptr && 3ull * (uint64_t)n <= (uint64_t)alt_cap) {
      uint32_t c2 = 1024u;
      while (c2 < 3u * (uint32_t)n) c2 <<= 1;
      if (c2 > alt_cap) c2 = alt_cap;
      int lg = 0;
      while ((1u << lg) < c2) lg++;
      htag = alt_tag; hkey = alt_key; cap = c2; shift = 32 - lg;
      if (max_probe > 256) max_probe = 256;
      return;
    }
    if (factor <= 0) return;
    uint32_t c2 = 1024u;
    const uint64_t want = (uint64_t)factor * (uint64_t)n;
    while ((uint64_t)c2 < want && c2 < cap_max) c2 <<= 1;
    int lg = 0;
    while ((1u << lg) < c2) lg++;
    cap = c2; shift = 32 - lg;
  }
  XS_HD uint32_t id(int bx, int by, int bz) const { return (uint32_t)bx + (uint32_t)hx * ((uint32_t)by + (uint32_t)hy * (uint32_t)bz); }
  template <class G> XS_D void clear(const G& g) const {
    for (uint32_t i = g.tid(); i < cap; i += g.size()) { htag[i] = XS_HEMPTY; hkey[i] = XS_HEMPTY; }
  }
  XS_HD uint32_t tag(int bx, int by, int bz) const { return id(bx, by, bz); }
  XS_HD uint32_t tag_delta(int ox, int oy, int oz) const { return (uint32_t)(ox + hx * (oy + hy * oz)); }
  XS_HD int claim(int bx, int by, int bz, uint32_t rank) const { return claim_tag(id(bx, by, bz), rank); }
  XS_HD int claim_tag(uint32_t t, uint32_t rank) const {   // 0 full / 1 not owner / 2 tentative owner
    uint32_t h = (t * 2654435761u) >> shift;
    for (int probes = 0; probes < max_probe; probes++) {
      const uint32_t old = atomic_cas_u32(&htag[h], XS_HEMPTY, t);
      if (old == XS_HEMPTY || old == t) return atomic_min_u32(&hkey[h], rank) > rank ? 2 : 1;
      h = (h + 1u) & (cap - 1u);
    }
    return 0;
  }
  XS_HD bool batched() const { return htag != alt_tag; }   // table in HBM (not the shared-memory alternative)
  // Claims for one z-layer of the 3x3x3 neighbourhood: bit j = oy * 3 + ox of m9 set -> claim block (tag tbase + delta_j);
  // returns the bits that are tentative owners.  With the table in HBM a claim is two dependent L2 round trips
  // (tag, then rank), so the nine claims of a layer are issued as batches of independent operations: nine tag loads,
  // the few compare-and-swaps that are needed, nine rank loads, the few atomicMins that can win.  Plain loads first
  // are safe even when stale: a tag never changes once set, a rank only decreases.
  XS_HD uint32_t claim9(uint32_t tbase, uint32_t m9, uint32_t rank, bool* full) const {
    uint32_t t[9], h[9], e[9];
#pragma unroll
    for (int j = 0; j < 9; j++) {
      t[j] = tbase + tag_delta(j % 3 - 1, j / 3 - 1, 0);
      h[j] = (t[j] * 2654435761u) >> shift;
      e[j] = ((m9 >> j) & 1u) ? htag[h[j]] : t[j];
    }
#pragma unroll
    for (int j = 0; j < 9; j++)
      if (((m9 >> j) & 1u) && e[j] == XS_HEMPTY) e[j] = atomic_cas_u32(&htag[h[j]], XS_HEMPTY, t[j]);
#pragma unroll
    for (int j = 0; j < 9; j++) {
      if (((m9 >> j) & 1u) && e[j] != XS_HEMPTY && e[j] != t[j]) {
        // slot taken by another block: linear probing, one at a time (rare)
        uint32_t hh = h[j];
        bool ok = false;
        for (int probes = 1; probes < max_probe; probes++) {
          hh = (hh + 1u) & (cap - 1u);
          const uint32_t old = atomic_cas_u32(&htag[hh], XS_HEMPTY, t[j]);
          if (old == XS_HEMPTY || old == t[j]) { ok = true; break; }
        }
        if (!ok) { *full = true; m9 &= ~(1u << j); }
        h[j] = hh;
      }
    }
    uint32_t k[9];
#pragma unroll
    for (int j = 0; j < 9; j++) k[j] = ((m9 >> j) & 1u) ? hkey[h[j]] : 0u;
    uint32_t won = 0u;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      if (((m9 >> j) & 1u) && k[j] > rank) {
        if (atomic_min_u32(&hkey[h[j]], rank) > rank) won |= 1u << j;
      }
    }
    return won;
  }
  XS_HD uint32_t owner(int bx, int by, int bz) const {
    const uint32_t t = id(bx, by, bz);
    uint32_t h = (t * 2654435761u) >> shift;
    while (htag[h] != t) h = (h + 1u) & (cap - 1u);
    return hkey[h];
  }
};
using WideHash = WideHashT<true>;        // wide / big tiers, path B
using WideHashSeq = WideHashT<false>;    // mid tier: its table is in shared memory except for the very largest sections


template <class CellT>
struct Arena {
  // lattice window: local cell (u,v) <-> lattice cell (ox+u, oy+v); tx x ty tiles of 32x32 cells
  int ox, oy, tx, ty, stride;   // stride = tx + 1 words per row (one zero pad word)
  uint32_t* bits;               // foreground-and-unvisited bitmap, (32*ty) rows x stride words
  uint32_t* tested;             // one bit per tile, ceil(tx*ty/32) words
  int halo;                     // tiles tested around a requested tile (0: just that tile)
  // samples
  CellT* order;                 // rank -> packed (u,v)
  int max_n;
  CellT* stack;                 // DFS stack ring (depth % stack_cap)
  int stack_cap;                // power of two when `spill` is set, else >= max_n
  CellT* spill;                 // global backing store for the ring, or null when stack_cap >= max_n
  uint32_t* mask;               // per-sample 27-bit candidate / kept-item masks (warp tier: aliases `bits`)
  // neighbour-mask variant of the window (mid / wide tiers): wside^2 cells, local cell index u + v * wside; one byte per
  // cell = which of its 8 neighbours are foreground and unvisited (see flood_component_nbr).
  uint8_t* cells;               // null: use the bitmap
  int wside;                    // sampled window: cells per side, a multiple of 32 (<= 448)
  int mside, moff, mlim;        // neighbour bytes: mside bytes per row, for window cells [moff, moff + mlim) of both axes;
                                //   byte index ci = (u - moff) + (v - moff) * mside; the outermost cells (0 and mlim - 1) are
                                //   the ring where the walk gives up.  mid / wide tiers: moff 0, mside = mlim = wside.
  uint32_t* fbits;              // foreground bitmap of the window, one zero row / word of padding all round
  int fstride;                  //   words per row = wside/32 + 2
  const int16_t* nlut;          // neighbour byte -> cell-index delta of its first set bit
  int compact_pairs;            // emit pass 1 evaluates compacted (sample, offset) pairs (needs `mask` in shared memory)
  uint16_t* pairs2;             // optional scratch for emit pass 2 on compacted items (>= 27 entries per lane of a
  int pairs2_cap;               //   one-warp group: the warp tier's first-touch table, idle by then)
  uint32_t* mask_alt;           // optional shared-memory home for the masks of sections of <= mask_alt_cap samples
  int mask_alt_cap;             //   (the neighbour-byte tiers: their foreground bitmap is dead after the walk)
};

struct Ctl {          // lives in shared memory (one per group)
  int status;
  int req_tx, req_ty;
  int n;              // samples visited so far
  int depth;          // DFS stack depth
  int lo;             // lowest stack depth resident in the ring
  int u, v;           // current cell
  int m;              // items emitted
  int npoly;          // polygons emitted
  int overflow;
  uint32_t contact;
  uint32_t counter;   // generic atomic counter (path B output length)
  uint32_t item_off, poly_off;
  int tiles;          // lattice tiles sampled (1024 cells each)
  uint32_t grow[7];   // neighbour-mask flood: tiles wanted by the closure (TileSet)
  long long tmark;    // phase profiling (thread 0): last timestamp and per-phase cycle counts
  uint32_t t[6];      //   0 flood (tiles + walk)  1 claims  2 emit  3 tile sampling (of 0)  4 emit pass 1 (of 2)  5 walk only (of 0)
};

XS_HD long long xs_clock() {
#ifdef __CUDA_ARCH__
  return clock64();
#else
  return 0;
#endif
}
// thread 0 only
XS_HD void prof_mark(Ctl& ctl, int phase) {
  const long long now = xs_clock();
  ctl.t[phase] += (uint32_t)(now - ctl.tmark);
  ctl.tmark = now;
}

enum { ST_DONE = 0, ST_NEED_TILE = 1, ST_OVERFLOW = 2 };
enum { Q_OK = 0, Q_OVERFLOW = 1, Q_CAPACITY = 2 };

template <class CellT>
XS_HD bool tile_tested(const Arena<CellT>& A, int tx, int ty) {
  const int t = ty * A.tx + tx;
  return (A.tested[t >> 5] >> (t & 31)) & 1u;
}

// Sample the lattice rows of every untested tile in the tile rectangle [tx0,tx1]x[ty0,ty1] (clipped).
template <class G, class CellT>
XS_D void test_tiles(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, Ctl& ctl, int tx0, int ty0, int tx1, int ty1) {
  tx0 = tx0 < 0 ? 0 : tx0; ty0 = ty0 < 0 ? 0 : ty0;
  tx1 = tx1 >= A.tx ? A.tx - 1 : tx1; ty1 = ty1 >= A.ty ? A.ty - 1 : ty1;
  const int nx = tx1 - tx0 + 1, ny = ty1 - ty0 + 1;
#ifdef __CUDA_ARCH__
  // a unit = four consecutive rows of one tile: their four volume loads are issued before the first ballot needs its answer
  const int units = nx * ny * 8;
  for (int unit = g.warp(); unit < units; unit += g.nwarps()) {
    const int row0 = (unit & 7) * 4, t = unit >> 3;
    const int tx = tx0 + t % nx, ty = ty0 + t / nx;
    if (tile_tested(A, tx, ty)) continue;   // warp-uniform
    const int u0 = tx * 32;
    uint32_t cube[4], bit[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) ok[j] = cell_probe(c, vol, A.ox + u0 + g.lane(), A.oy + ty * 32 + row0 + j, &cube[j], &bit[j]);
    uint32_t cb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) cb[j] = cube_load(vol, cube[j]);     // (index 0 when the cell is outside: a valid address)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t word = __ballot_sync(0xffffffffu, ok[j] && ((cb[j] >> bit[j]) & 1u));
      if (g.lane() == 0) A.bits[(size_t)(ty * 32 + row0 + j) * A.stride + tx] = word;
    }
  }
#else
  const int units = nx * ny * 32;
  for (int unit = g.warp(); unit < units; unit += g.nwarps()) {
    const int row = unit & 31, t = unit >> 5;
    const int tx = tx0 + t % nx, ty = ty0 + t / nx;
    if (tile_tested(A, tx, ty)) continue;
    const int v = ty * 32 + row, u0 = tx * 32;
    uint32_t word = 0;
    for (int l = 0; l < 32; l++) word |= (uint32_t)cell_fg(c, vol, A.ox + u0 + l, A.oy + v) << l;
    A.bits[(size_t)v * A.stride + tx] = word;
  }
#endif
  g.sync();
  if (g.tid() == 0) {
    for (int ty = ty0; ty <= ty1; ty++)
      for (int tx = tx0; tx <= tx1; tx++) {
        const int t = ty * A.tx + tx;
        if (!((A.tested[t >> 5] >> (t & 31)) & 1u)) ctl.tiles++;
        A.tested[t >> 5] |= 1u << (t & 31);
      }
  }
  g.sync();
}

XS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

XS_HD int popc8(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x & 0xffu);
#else
  return __builtin_popcount(x & 0xffu);
#endif
}
XS_HD int popcll64(unsigned long long x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
XS_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
XS_HD int ffs32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __ffs(x);
#else
  return __builtin_ffs(x);
#endif
}
// Walk direction table: index = the 3x3 availability neighbourhood (rows v-1, v, v+1, three bits each: u-1,u,u+1);
// entry = 0xff when no neighbour is available, else (du+1) | (dv+1)<<2 | more<<4 for the FIRST available neighbour
// in the reference's priority order (more = at least one other neighbour is available too) — the reverse of its push order (xs3d.hpp:926-950):
// (+1,+1) (-1,+1) (+1,-1) (-1,-1) (0,+1) (0,-1) (+1,0) (-1,0).
struct DirLut { uint8_t e[512]; };
constexpr DirLut make_dir_lut() {
  DirLut t{};
  const int du[8] = { 1, -1, 1, -1, 0, 0, 1, -1 };
  const int dv[8] = { 1, 1, -1, -1, 1, -1, 0, 0 };
  for (int idx = 0; idx < 512; idx++) {
    int found = 0xff, count = 0;
    for (int d = 7; d >= 0; d--) {
      const int row = dv[d] + 1, col = du[d] + 1;       // row 0 = v-1, row 2 = v+1
      if ((idx >> (3 * row + col)) & 1) { found = (du[d] + 1) | ((dv[d] + 1) << 2); count++; }
    }
    if (count >= 2) found |= 16;                        // bit 4: other neighbours remain after taking this one
    t.e[idx] = (uint8_t)found;
  }
  return t;
}
#ifdef __CUDACC__
__constant__ DirLut g_dir_lut = make_dir_lut();
#endif
static const DirLut h_dir_lut = make_dir_lut();
XS_HD uint32_t dir_lut(uint32_t idx) {
#ifdef __CUDA_ARCH__
  return g_dir_lut.e[idx];
#else
  return h_dir_lut.e[idx];
#endif
}

// Depth-first walk over the bitmap, run by ONE thread.  A visited cell has its bit cleared, so "set" means
// foreground-and-unvisited, which only ever decreases: re-examining a cell after a child returns picks
// exactly the neighbour the reference's explicit stack would pop next.  The loop is latency critical (one
// thread, one dependent chain per step): 32-bit index arithmetic, one table lookup for the direction, the
// visited bit cleared with a fire-and-forget atomic, spill code out of line.
// Returns ST_DONE / ST_NEED_TILE (ctl.req_*) / ST_OVERFLOW.
XS_HD void clear_bit(uint32_t* w, uint32_t bit) {
#ifdef __CUDA_ARCH__
  atomicAnd(w, ~bit);
#else
  *w &= ~bit;
#endif
}

// 3x3 availability of cell (u,v): bit 3*row + col, rows v-1..v+1, cols u-1..u+1
template <class CellT>
XS_HD uint32_t avail3x3(const Arena<CellT>& A, int u, int v) {
  const int um1 = u - 1;
  const int sh = um1 & 31;
  const uint32_t* rp = A.bits + ((v - 1) * A.stride + (um1 >> 5));
  const int stride = A.stride;
  const uint32_t a0 = rp[0], a1 = rp[1], b0 = rp[stride], b1 = rp[stride + 1], c0 = rp[2 * stride], c1 = rp[2 * stride + 1];
  return (funnel_r(a0, a1, sh) & 7u) | ((funnel_r(b0, b1, sh) & 7u) << 3) | ((funnel_r(c0, c1, sh) & 7u) << 6);
}
// A cell on a tile border may have neighbours in a tile that has not been sampled yet: which one (or -1)
template <class CellT>
XS_HD int missing_tile(const Arena<CellT>& A, int u, int v) {
  if (!((((u + 1) & 31) < 2) | (((v + 1) & 31) < 2))) return -1;
  const int ua = (u - 1) >> 5, ub = (u + 1) >> 5, va = (v - 1) >> 5, vb = (v + 1) >> 5;
  if (!tile_tested(A, ua, va)) return va * A.tx + ua;
  if (!tile_tested(A, ub, va)) return va * A.tx + ub;
  if (!tile_tested(A, ua, vb)) return vb * A.tx + ua;
  if (!tile_tested(A, ub, vb)) return vb * A.tx + ub;
  return -1;
}

// Walk, warp form.  Device: called by every lane of ONE warp in lock step (all lanes hold the same state; lane 0
// does the stores); host: lane = 0.  Forward steps are inherently one at a time.  A dead end, however, unwinds the
// stack until a frame that still has something to offer — and since availability only ever decreases, every frame
// can be judged independently: the lanes examine 32 frames at once and the walk resumes at the topmost live one
// (or at the topmost one that needs an unsampled tile, exactly where the one-at-a-time unwinding would stop).
// Unwinding is half of all iterations of a one-thread walk, most of it the final unwind of a finished component.
template <class CellT>
XS_HD int dfs_run(const Arena<CellT>& A, Ctl& ctl, int lane) {
  typedef CellPack<CellT> P;
  int n = ctl.n, depth = ctl.depth, lo = ctl.lo, u = ctl.u, v = ctl.v;
  const unsigned Wm2 = (unsigned)(A.tx * 32 - 2), Hm2 = (unsigned)(A.ty * 32 - 2);
  const int stride = A.stride, max_n = A.max_n;
  uint32_t* const bits = A.bits;
  CellT* const stack = A.stack;
  CellT* const order = A.order;
  const bool ring = A.spill != nullptr;
  const int smask = ring ? (A.stack_cap - 1) : 0x7fffffff;   // no ring wrap when the whole stack is resident
  const int cap = ring ? A.stack_cap : 0x7fffffff;
  int status;
  int req = -1;
  for (;;) {
    // all four tiles under the 3x3 neighbourhood must have been sampled (only cells on a tile border can miss)
    req = missing_tile(A, u, v);
    if (req >= 0) { status = ST_NEED_TILE; break; }
    const uint32_t idx = avail3x3(A, u, v);
    uint32_t e;
    if (sizeof(CellT) == 2) {
      e = dir_lut(idx);                       // warp tier (issue-bound): one table load
    } else {
      // CTA tiers (latency-bound): an indexed constant load costs more than the ALU it saves.
      // idx bits: rm 0..2, r0 3..5, rq 6..8  ->  priority mask [rq2 rq0 rm2 rm0 rq1 rm1 r0_2 r0_0]
      const uint32_t pm = ((idx >> 8) & 1u) | ((idx >> 5) & 2u) | (idx & 4u) | ((idx << 3) & 8u) | ((idx >> 3) & 16u) | ((idx << 4) & 32u) |
                          ((idx << 1) & 64u) | ((idx << 4) & 128u);
      if (pm == 0u) {
        e = 0xffu;
      } else {
        const int d = ffs32(pm) - 1;
        e = ((0x2522u >> (2 * d)) & 3u) | (((0x520au >> (2 * d)) & 3u) << 2) | ((pm & (pm - 1u)) ? 16u : 0u);
      }
    }
    if (e == 0xffu) {
      // dead end: drop the current frame, then find the topmost frame that is live or needs a tile
      depth--;
      bool found = false;
      while (depth > 0) {
        if (depth == lo) {
          // (ring only) refill the lower half of the ring from the global spill area
          int nlo = lo - A.stack_cap / 2;
          if (nlo < 0) nlo = 0;
#ifdef __CUDA_ARCH__
          for (int d = nlo + lane; d < lo; d += 32) stack[d & smask] = A.spill[d];
          __syncwarp();
#else
          for (int d = nlo; d < lo; d++) stack[d & smask] = A.spill[d];
#endif
          lo = nlo;
        }
        const int avail = depth - lo;
        const int cnt = avail < 32 ? avail : 32;
#ifdef __CUDA_ARCH__
        CellT f = 0;
        bool live = false;
        if (lane < cnt) {
          f = stack[(depth - 1 - lane) & smask];
          const int fu = P::u(f), fv = P::v(f);
          live = missing_tile(A, fu, fv) >= 0 || (avail3x3(A, fu, fv) & 0x1efu) != 0u;
        }
        const uint32_t ball = __ballot_sync(0xffffffffu, live);
        if (ball) {
          const int l = __ffs(ball) - 1;
          depth -= l;
          const uint32_t ff = __shfl_sync(0xffffffffu, (uint32_t)f, l);
          u = P::u((CellT)ff); v = P::v((CellT)ff);
          found = true;
          break;
        }
        depth -= cnt;
#else
        int l = 0;
        for (; l < cnt; l++) {
          const CellT f = stack[(depth - 1 - l) & smask];
          const int fu = P::u(f), fv = P::v(f);
          if (missing_tile(A, fu, fv) >= 0 || (avail3x3(A, fu, fv) & 0x1efu) != 0u) { u = fu; v = fv; found = true; break; }
        }
        depth -= l;
        if (found) break;
#endif
      }
      if (!found) { status = ST_DONE; break; }
      continue;
    }
    u += (int)(e & 3u) - 1; v += (int)((e >> 2) & 3u) - 1;
    if (lane == 0) bits[v * stride + (u >> 5)] &= ~(1u << (u & 31));
    if (((unsigned)(u - 1) >= Wm2) | ((unsigned)(v - 1) >= Hm2) | (n >= max_n)) { status = ST_OVERFLOW; break; }
    const CellT p = P::pack(u, v);
    if (lane == 0) order[n] = p;
    n++;
    // Tail call: if the cell we are leaving has no other available neighbour NOW it never will (availability only
    // decreases), so coming back to it would find nothing — reuse its frame instead of stacking on top of it.
    if (!(e & 16u)) depth--;
    if (depth - lo == cap) {
      // (ring only) ring full: spill the lower half
      const int nlo = lo + A.stack_cap / 2;
#ifdef __CUDA_ARCH__
      for (int dd = lo + lane; dd < nlo; dd += 32) A.spill[dd] = stack[dd & smask];
#else
      for (int dd = lo; dd < nlo; dd++) A.spill[dd] = stack[dd & smask];
#endif
      lo = nlo;
    }
    if (lane == 0) stack[depth & smask] = p;
    depth++;
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
  }
  if (lane == 0) {
    ctl.n = n; ctl.depth = depth; ctl.lo = lo; ctl.u = u; ctl.v = v;
    if (status == ST_NEED_TILE) { ctl.req_tx = req % A.tx; ctl.req_ty = req / A.tx; }
  }
  return status;
}

// Steps 1-2: sample tiles on demand and walk the component.  On return ctl.n samples are in
// A.order (rank order).  Returns Q_OK or Q_OVERFLOW; *empty is set when the centre cell itself is
// not foreground on the lattice (the reference then returns zeros).
template <class G, class CellT>
XS_D int flood_component(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, Ctl& ctl, bool* empty) {
  typedef CellPack<CellT> P;
  const int words = A.ty * 32 * A.stride;
  for (int i = g.tid(); i < words; i += g.size()) A.bits[i] = 0u;
  const int tw = (A.tx * A.ty + 31) >> 5;
  for (int i = g.tid(); i < tw; i += g.size()) A.tested[i] = 0u;
  const int cu = c.c - A.ox, cv = c.c - A.oy;
  if (g.tid() == 0) {
    ctl.status = ST_NEED_TILE; ctl.req_tx = cu >> 5; ctl.req_ty = cv >> 5;
    ctl.n = 0; ctl.depth = 0; ctl.lo = 0; ctl.u = cu; ctl.v = cv; ctl.m = 0; ctl.npoly = 0; ctl.overflow = 0; ctl.contact = 0u; ctl.counter = 0u; ctl.tiles = 0;
    ctl.item_off = 0u; ctl.poly_off = 0u;
    for (int i = 0; i < 6; i++) ctl.t[i] = 0u;
    ctl.tmark = xs_clock();
  }
  g.sync();
  *empty = false;
  bool first = true;
  for (;;) {
    const int rtx = ctl.req_tx, rty = ctl.req_ty;
    g.sync();
    test_tiles(g, c, vol, A, ctl, rtx - A.halo, rty - A.halo, rtx + A.halo, rty + A.halo);
    if (g.warp() == 0) {
      // one warp walks, in lock step (dfs_run)
      int st;
      const long long w0 = xs_clock();
      bool start = true;
      if (first) {
        uint32_t* w = &A.bits[(size_t)cv * A.stride + (cu >> 5)];
        const bool centre_fg = ((*w >> (cu & 31)) & 1u) != 0u;
        g.converge();
        if (!centre_fg) {
          start = false;
          if (g.lane() == 0) ctl.n = 0;
        } else if (g.lane() == 0) {
          *w &= ~(1u << (cu & 31));
          const CellT p = P::pack(cu, cv);
          A.order[0] = p; A.stack[0] = p;
          ctl.n = 1; ctl.depth = 1;
        }
        g.converge();
      }
      st = start ? dfs_run(A, ctl, g.lane()) : ST_DONE;
      if (g.lane() == 0) {
        ctl.t[5] += (uint32_t)(xs_clock() - w0);
        ctl.status = st;
      }
    }
    first = false;
    g.sync();
    const int st = ctl.status;
    if (st == ST_NEED_TILE) continue;
    if (st == ST_OVERFLOW) return Q_OVERFLOW;
    break;
  }
  *empty = (ctl.n == 0);
  return Q_OK;
}

// ------------------------------------------------------------------------------------------------
// Neighbour-mask flood (mid tier).  Same semantics as the bitmap flood above, organised so that the
// sequential part — the depth-first walk — has the shortest dependent chain per step we know how to build:
//   * tiles are sampled BEFORE the walk by a parallel closure over 32x32-cell tiles (a tile is sampled when a
//     sampled foreground cell touches it), so the walk never stops to ask for cells;
//   * every foreground cell then gets one byte M = which of its 8 neighbours are foreground-and-unvisited, bit k
//     = k-th neighbour in the reference's priority order (+1,+1) (-1,+1) (+1,-1) (-1,-1) (0,+1) (0,-1) (+1,0) (-1,0);
//   * a walk step is: load M[cell] -> look the byte up in a 256-entry table of cell-index deltas -> add.  Visiting
//     a cell clears its bit in its eight neighbours' bytes with fire-and-forget atomics, one lane per neighbour;
//   * the walk is run by one WARP in lock step (every lane holds the same state): when it hits a dead end the
//     lanes examine 32 stack frames at once, so unwinding — half of all iterations of a one-thread walk, most of
//     it the final unwind of a finished component — costs almost nothing.
// ------------------------------------------------------------------------------------------------
// delta of the first set bit of m for a window of W cells per row (entry 0 unused)
XS_HD int nbr_delta(uint32_t m, int W) {
  const int d = ffs32(m) - 1;
  const int du = (int)((0x2522u >> (2 * d)) & 3u) - 1;   // d:0..7 -> +1,-1,+1,-1,0,0,+1,-1
  const int dv = (int)((0x520au >> (2 * d)) & 3u) - 1;   // d:0..7 -> +1,+1,-1,-1,+1,-1,0,0
  return du + dv * W;
}
template <class G>
XS_D void nbr_lut_fill(const G& g, int16_t* lut, int W) {
  for (int m = g.tid(); m < 256; m += g.size()) lut[m] = (int16_t)(m ? nbr_delta((uint32_t)m, W) : 0);
}

template <class CellT>
XS_HD uint32_t* frow(const Arena<CellT>& A, int v) { return A.fbits + (size_t)(v + 1) * A.fstride + 1; }

// A set of 32x32-cell tiles of the window (bit = ty * nt + tx), up to 32 * NW tiles.
template <int NW>
struct TileSetT {
  uint32_t w[NW];
  XS_HD void clear() {
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0u;
  }
  XS_HD void set(int t) { set_word(t >> 5, 1u << (t & 31)); }
  XS_HD bool test(int t) const { return (w[t >> 5] >> (t & 31)) & 1u; }
  XS_HD bool any() const {
    uint32_t o = 0u;
#pragma unroll
    for (int i = 0; i < NW; i++) o |= w[i];
    return o != 0u;
  }
  XS_HD int count() const {
    int c = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) c += popc32(w[i]);
    return c;
  }
  // f(t) for every tile of the set, ascending; word indices are compile-time constants so the set stays in registers
  template <class F> XS_HD void for_each(F f) const {
#pragma unroll
    for (int i = 0; i < NW; i++) {
      uint32_t m = w[i];
      while (m) {
        const int b = ffs32(m) - 1;
        m &= m - 1u;
        f(32 * i + b);
      }
    }
  }
  XS_HD void set_word(int i, uint32_t bits) {
#pragma unroll
    for (int j = 0; j < NW; j++)
      if (j == i) w[j] |= bits;
  }
};
// words a group / cell type pairing needs: one warp per query -> 96-cell window (9 tiles); CTA tiers: 256-cell window
// (64 tiles) with 16-bit cells, up to 448 cells (196 tiles) with 32-bit cells; host test builds: anything
template <class G, class CellT> struct TileWords { static constexpr int v = 7; };
#ifdef __CUDACC__
template <class CellT> struct TileWords<WarpGroup, CellT> { static constexpr int v = 1; };
template <> struct TileWords<BlockGroup, uint16_t> { static constexpr int v = 2; };
#endif

// Sample tile (tx,ty) into the foreground bitmap: a warp per lattice row, a lane per cell.
template <class G, class CellT>
XS_D void nbr_sample_tile(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, int tx, int ty) {
#ifdef __CUDA_ARCH__
  // four rows at a time: their four volume loads are issued before the first ballot needs its answer
  for (int row0 = 4 * g.warp(); row0 < 32; row0 += 4 * g.nwarps()) {
    uint32_t cube[4], bit[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) ok[j] = cell_probe(c, vol, A.ox + tx * 32 + g.lane(), A.oy + ty * 32 + row0 + j, &cube[j], &bit[j]);
    uint32_t cb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) cb[j] = cube_load(vol, cube[j]);     // (index 0 when the cell is outside: a valid address)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t word = __ballot_sync(0xffffffffu, ok[j] && ((cb[j] >> bit[j]) & 1u));
      if (g.lane() == 0) frow(A, ty * 32 + row0 + j)[tx] = word;
    }
  }
#else
  for (int row = 0; row < 32; row++) {
    const int v = ty * 32 + row;
    uint32_t word = 0;
    for (int l = 0; l < 32; l++) word |= (uint32_t)cell_fg(c, vol, A.ox + tx * 32 + l, A.oy + v) << l;
    frow(A, v)[tx] = word;
  }
#endif
}

// Which neighbours of the freshly sampled tile (tx,ty) hold cells adjacent to one of its foreground cells?
// (evaluated by one warp; every lane returns the same set)
template <class G, class CellT, class TS>
XS_D void nbr_grow_tile(const G& g, const Arena<CellT>& A, int tx, int ty, TS& want) {
  const int nt = A.wside >> 5, t = ty * nt + tx;
  uint32_t left, right, top, bottom;
#ifdef __CUDA_ARCH__
  {
    const uint32_t w = frow(A, ty * 32 + g.lane())[tx];
    left = __ballot_sync(0xffffffffu, w & 1u);
    right = __ballot_sync(0xffffffffu, w >> 31);
    top = __shfl_sync(0xffffffffu, w, 0);
    bottom = __shfl_sync(0xffffffffu, w, 31);
  }
#else
  left = right = 0u;
  for (int l = 0; l < 32; l++) {
    const uint32_t w = frow(A, ty * 32 + l)[tx];
    left |= (w & 1u) << l; right |= (w >> 31) << l;
  }
  top = frow(A, ty * 32)[tx]; bottom = frow(A, ty * 32 + 31)[tx];
#endif
  const bool okl = tx > 0, okr = tx < nt - 1, okt = ty > 0, okb = ty < nt - 1;
  if (left && okl) want.set(t - 1);
  if (right && okr) want.set(t + 1);
  if (top && okt) want.set(t - nt);
  if (bottom && okb) want.set(t + nt);
  if ((top & 1u) && okl && okt) want.set(t - nt - 1);
  if ((top >> 31) && okr && okt) want.set(t - nt + 1);
  if ((bottom & 1u) && okl && okb) want.set(t + nt - 1);
  if ((bottom >> 31) && okr && okb) want.set(t + nt + 1);
}

// Neighbour bytes of every foreground cell of tile (tx,ty).  A cell on the window's outer ring gets 0: the walk
// stops when it steps on one (the section leaves the window).
template <class G, class CellT>
XS_D void nbr_masks_tile(const G& g, const Arena<CellT>& A, int tx, int ty) {
  const int MS = A.mside, ML = A.mlim, MO = A.moff;
  for (int row = g.warp(); row < 32; row += g.nwarps()) {
    const int v = ty * 32 + row;
    const uint32_t* rm = frow(A, v - 1) + tx;
    const uint32_t* r0 = frow(A, v) + tx;
    const uint32_t* rq = frow(A, v + 1) + tx;
    const uint32_t cur = r0[0];
    if (cur == 0u) continue;                     // warp-uniform
    // bit (l + 1) of x? = cell (32 tx + l) of the row; bits 0 and 33 come from the neighbouring words
    const unsigned long long xm = ((unsigned long long)rm[0] << 1) | (rm[-1] >> 31) | ((unsigned long long)(rm[1] & 1u) << 33);
    const unsigned long long x0 = ((unsigned long long)cur << 1) | (r0[-1] >> 31) | ((unsigned long long)(r0[1] & 1u) << 33);
    const unsigned long long xq = ((unsigned long long)rq[0] << 1) | (rq[-1] >> 31) | ((unsigned long long)(rq[1] & 1u) << 33);
#ifdef __CUDA_ARCH__
    const int l = g.lane();
    {
#else
    for (int l = 0; l < 32; l++) {
#endif
      const int u = tx * 32 + l;
      const uint32_t am = (uint32_t)(xm >> l) & 7u, a0 = (uint32_t)(x0 >> l) & 7u, aq = (uint32_t)(xq >> l) & 7u;   // bits: u-1, u, u+1
      const int mu = u - MO, mv = v - MO;
      if ((a0 & 2u) && mu >= 0 && mv >= 0 && mu < ML && mv < ML) {
        const bool inner = mu > 0 && mv > 0 && mu < ML - 1 && mv < ML - 1;
        const uint32_t m = ((aq >> 2) & 1u) | ((aq & 1u) << 1) | (((am >> 2) & 1u) << 2) | ((am & 1u) << 3) |
                           (((aq >> 1) & 1u) << 4) | (((am >> 1) & 1u) << 5) | (((a0 >> 2) & 1u) << 6) | ((a0 & 1u) << 7);
        A.cells[mv * MS + mu] = (uint8_t)(inner ? m : 0u);
      }
    }
  }
}

// order[] entry of window cell ci = u + v * W: the index itself for 16-bit cells (W = 256: u | v << 8), packed
// (u | v << 16) for 32-bit cells — what CellPack / sample_at decode
template <class CellT>
XS_HD CellT nbr_order_entry(int ci, int MS, int MO) {
  if (sizeof(CellT) == 2 && MS == 256) return (CellT)ci;      // (256-cell mid tier: the byte index is the packed cell already)
  const int v = MS == 64 ? (ci >> 6) : (MS == 160 ? ci / 160 : ci / MS), u = ci - v * MS;   // (constant divisors spare an integer division)
  return CellPack<CellT>::pack(u + MO, v + MO);
}

// Mark cell ci visited: clear its bit in the neighbour bytes of its eight neighbours (the bit a neighbour holds for
// us is the opposite direction, k^3 for the diagonals and k^1 for the others).  Plain form, one neighbour at a time.
XS_HD void nbr_visit_one(uint8_t* cells, int ci, int k, int W) {
  const int du = (int)((0x2522u >> (2 * k)) & 3u) - 1, dv = (int)((0x520au >> (2 * k)) & 3u) - 1;
  const int nb = ci + du + dv * W;
  const int opp = k < 4 ? (k ^ 3) : (k ^ 1);
  cells[nb] &= (uint8_t)~(1u << opp);
}
// Word form used by the warp walk: the three neighbours of a row are three consecutive bytes, i.e. one or two
// aligned 32-bit words.  Lane 2r+h (r = row 0..2, h = 0/1) read-modify-writes word h of row r — six lanes, six
// distinct words, no atomics.  `clr24` = the bits to clear in the three bytes (u-1, u, u+1) of this lane's row.
XS_HD uint32_t nbr_row_clear(int r) { return r == 0 ? 0x021001u : (r == 1 ? 0x800040u : 0x082004u); }

// The walk, plain form (host test builds; one thread).  Returns ST_DONE / ST_OVERFLOW.
template <class CellT>
XS_HD int dfs_run_nbr(const Arena<CellT>& A, int& n_io, int& depth_io, int& lo_io, int& ci_io) {
  int n = n_io, depth = depth_io, lo = lo_io, ci = ci_io;
  const int W = A.mside, ML = A.mlim;
  const int max_n = A.max_n;
  uint8_t* const cells = A.cells;
  const int16_t* const lut = A.nlut;
  CellT* const stack = A.stack;
  CellT* const order = A.order;
  const bool ring = A.spill != nullptr;
  const int smask = ring ? (A.stack_cap - 1) : 0x7fffffff;
  const int cap = ring ? A.stack_cap : 0x7fffffff;
  int status;
  for (;;) {
    const uint32_t m = cells[ci];
    if (m == 0u) {
      // dead end: drop the current frame, then find the topmost frame that still has an available neighbour
      // (examined 32 at a time, like the warp form)
      depth--;
      bool found = false;
      while (depth > 0) {
        if (depth == lo) {
          int nlo = lo - A.stack_cap / 2;
          if (nlo < 0) nlo = 0;
          for (int d = nlo; d < lo; d++) stack[d & smask] = A.spill[d];
          lo = nlo;
        }
        const int avail = depth - lo;
        const int cnt = avail < 32 ? avail : 32;
        int l = 0;
        for (; l < cnt; l++) {
          const int f = (int)stack[(depth - 1 - l) & smask];
          if (cells[f] != 0) { ci = f; found = true; break; }
        }
        depth -= l;
        if (found) break;
      }
      if (!found) { status = ST_DONE; break; }
      continue;
    }
    ci += (int)lut[m];
    const int v = ci / W, u = ci - v * W;
    if ((u == 0) | (v == 0) | (u == ML - 1) | (v == ML - 1) | (n >= max_n)) { status = ST_OVERFLOW; break; }
    for (int k = 0; k < 8; k++) nbr_visit_one(cells, ci, k, W);
    order[n++] = nbr_order_entry<CellT>(ci, W, A.moff);
    if ((m & (m - 1u)) == 0u) depth--;     // tail call: the cell we leave has nothing left, reuse its frame
    if (depth - lo == cap) {
      const int nlo = lo + A.stack_cap / 2;
      for (int dd = lo; dd < nlo; dd++) A.spill[dd] = stack[dd & smask];
      lo = nlo;
    }
    stack[depth & smask] = (CellT)ci;
    depth++;
  }
  n_io = n; depth_io = depth; lo_io = lo; ci_io = ci;
  return status;
}

#ifdef __CUDACC__
// Round trip through a per-lane shared-memory slot with volatile accesses: a value neither nvvm nor ptxas can
// re-derive, so it stays in its register.  `slots` = 32 words of scratch, all 32 lanes call.
__device__ __forceinline__ uint32_t pin_u32(uint32_t x, volatile uint32_t* slots) {
  slots[threadIdx.x & 31] = x;
  return slots[threadIdx.x & 31];
}
// Shared-memory accessors with explicit 32-bit shared addresses: no generic-address arithmetic on the walk's
// dependent chain, and (asm volatile + memory clobber) exactly the program order written below.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// stack-ring entries are CellT wide
template <class CellT> __device__ __forceinline__ uint32_t lds_cell(uint32_t base, int i) {
  return sizeof(CellT) == 2 ? lds_u16(base + 2u * (uint32_t)i) : lds_u32(base + 4u * (uint32_t)i);
}
template <class CellT> __device__ __forceinline__ void sts_cell(uint32_t base, int i, uint32_t v) {
  if (sizeof(CellT) == 2) sts_u16(base + 2u * (uint32_t)i, v); else sts_u32(base + 4u * (uint32_t)i, v);
}

// The walk, warp form: called by every lane of ONE warp, all lanes in lock step (they all hold the same state).
// cells / lut / stack ring must be in SHARED memory; order and spill in global memory.  The stack holds cell indices
// u + v * W; `order` gets what sample_at decodes (nbr_order_entry).
// Per forward step the dependent chain is: neighbour byte -> delta table -> add -> neighbour byte of the new cell,
// i.e. two shared-memory loads and one add.  The byte of the cell we arrive at is loaded BEFORE that cell is marked
// visited (marking only rewrites its neighbours' bytes): three rows of three consecutive bytes = at most two
// aligned words per row, read-modify-written by lanes 2r+h (the other lanes duplicate them: same value, same word).
// A dead end examines 32 stack frames at once.
template <class CellT>
__device__ __forceinline__ int dfs_run_nbr_warp(const Arena<CellT>& A, int& n_io, int& depth_io, int& lo_io, int& ci_io, int lane, uint32_t* scratch32) {
  int n = n_io, depth = depth_io, lo = lo_io, ci = ci_io;
  const int W = A.mside;
  const int max_n = A.max_n;
  // loop invariants are pinned in registers (opaque moves): ptxas otherwise re-derives them — special-register reads
  // and the lane arithmetic included — inside the loop, on the dependent chain
  const uint32_t s_cells = pin_u32((uint32_t)__cvta_generic_to_shared(A.cells), scratch32);
  const uint32_t s_lut = pin_u32((uint32_t)__cvta_generic_to_shared(A.nlut), scratch32);
  const uint32_t s_stack = pin_u32((uint32_t)__cvta_generic_to_shared(A.stack), scratch32);
  CellT* const order = A.order;
  CellT* const spill = A.spill;
  const bool ring = spill != nullptr;
  const int smask = ring ? (A.stack_cap - 1) : 0x7fffffff;
  const int cap = ring ? A.stack_cap : 0x7fffffff;
  const int half = A.stack_cap / 2;
  const int l6 = lane % 6;
  const int vr = l6 >> 1;
  const bool hiword = (l6 & 1) != 0;
  const uint32_t vh4 = pin_u32(hiword ? 4u : 0u, scratch32);
  const uint32_t vsel = pin_u32(s_cells + (uint32_t)((vr - 1) * W - 1), scratch32);   // + ci = address of the first of this lane's three cells
  // the clear mask of this lane's word = upper half of (fhi:flo) << 8 * (address & 3)
  const uint32_t fhi = pin_u32(hiword ? 0u : nbr_row_clear(vr), scratch32), flo = pin_u32(hiword ? nbr_row_clear(vr) : 0u, scratch32);
  const uint32_t lim = (uint32_t)(A.mlim - 2);
  const bool lane0 = pin_u32((uint32_t)lane, scratch32) == 0u;
  int status;
  int budget = 0;   // steps that can be taken before `order` or the stack ring may be full (re-evaluated when it runs out)
  uint32_t m = lds_u8(s_cells + (uint32_t)ci);
  for (;;) {
    while (m != 0u) {
      ci += lds_s16(s_lut + 2u * m);
      const uint32_t m_next = lds_u8(s_cells + (uint32_t)ci);     // loads first (in range: the arrays are padded by a row)
      const uint32_t vb = vsel + (uint32_t)ci;
      const uint32_t va = (vb & ~3u) + vh4;
      const uint32_t vold = lds_u32(va);
      if ((m & (m - 1u)) == 0u) depth--;                           // tail call: the cell we leave has nothing left, reuse its frame
      if (--budget < 0) {
        if (n >= max_n) { status = ST_OVERFLOW; goto out; }
        if (depth - lo == cap) {
          // ring full: spill the lower half
          const int nlo = lo + half;
          for (int dd = lo + lane; dd < nlo; dd += 32) spill[dd] = (CellT)lds_cell<CellT>(s_stack, dd & smask);
          __syncwarp();
          lo = nlo;
        }
        const int room_n = max_n - n, room_s = cap - (depth - lo);
        budget = (room_n < room_s ? room_n : room_s) - 1;
      }
      sts_u32(va, vold & ~__funnelshift_l(flo, fhi, vb << 3));
      if (lane0) {
        order[n] = nbr_order_entry<CellT>(ci, W, A.moff);
        sts_cell<CellT>(s_stack, depth & smask, (uint32_t)ci);
      }
      n++;
      depth++;
      m = m_next;
    }
    // dead end.  A cell on the window's outer ring is one by construction (its byte is 0): the section leaves the window.
    {
      const uint32_t v = (uint32_t)ci / (uint32_t)W, u = (uint32_t)ci - v * (uint32_t)W;
      if (((u - 1u) >= lim) | ((v - 1u) >= lim)) { status = ST_OVERFLOW; goto out; }
    }
    // dead end: drop the current frame, then find the topmost frame that still has an available neighbour
    __syncwarp();
    depth--;
    bool found = false;
    while (depth > 0) {
      if (depth == lo) {
        // (ring only) refill the lower half of the ring from the global spill area
        int nlo = lo - half;
        if (nlo < 0) nlo = 0;
        for (int d = nlo + lane; d < lo; d += 32) sts_cell<CellT>(s_stack, d & smask, (uint32_t)spill[d]);
        __syncwarp();
        lo = nlo;
      }
      const int avail = depth - lo;
      const int cnt = avail < 32 ? avail : 32;
      uint32_t f = 0u, fm = 0u;
      if (lane < cnt) { f = lds_cell<CellT>(s_stack, (depth - 1 - lane) & smask); fm = lds_u8(s_cells + f); }
      const uint32_t ball = __ballot_sync(0xffffffffu, fm != 0u);
      if (ball) {
        const int l = __ffs(ball) - 1;
        depth -= l;
        ci = (int)__shfl_sync(0xffffffffu, f, l);
        m = __shfl_sync(0xffffffffu, fm, l);
        found = true;
        break;
      }
      depth -= cnt;
    }
    if (!found) { status = ST_DONE; break; }
  }
out:
  n_io = n; depth_io = depth; lo_io = lo; ci_io = ci;
  return status;
}
#endif

template <class G, class CellT>
XS_D int flood_component_nbr(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, Ctl& ctl, bool* empty) {
  const int W = A.wside, nt = W >> 5, MS = A.mside;
  {
    // rows -1 .. W, one pad word either side of each row — or, with fstride = W/32 + 1, a single pad word per row that
    // serves as the right pad of its row and the left pad of the next (then one word more at the very end)
    const int words = (W + 2) * A.fstride + (A.fstride == nt + 1 ? 1 : 0);
    for (int i = g.tid(); i < words; i += g.size()) A.fbits[i] = 0u;
  }
  const int cu = c.c - A.ox, cv = c.c - A.oy;
  if (g.tid() == 0) {
    ctl.status = ST_DONE; ctl.req_tx = 0; ctl.req_ty = 0;
    ctl.n = 0; ctl.depth = 0; ctl.lo = 0; ctl.u = (cu - A.moff) + (cv - A.moff) * MS; ctl.v = 0; ctl.m = 0; ctl.npoly = 0; ctl.overflow = 0; ctl.contact = 0u; ctl.counter = 0u; ctl.tiles = 0;
    ctl.item_off = 0u; ctl.poly_off = 0u;
    for (int i = 0; i < 7; i++) ctl.grow[i] = 0u;
    for (int i = 0; i < 6; i++) ctl.t[i] = 0u;
    ctl.tmark = xs_clock();
  }
  g.sync();
  *empty = false;
  // closure over tiles, starting from the centre cell's tile (and `halo` tiles around it): a tile is sampled as soon
  // as a sampled foreground cell touches it, so the walk never has to stop and ask for cells
  constexpr int NW = TileWords<G, CellT>::v;
  typedef TileSetT<NW> TileSet;
  if (nt * nt > 32 * NW) return Q_OVERFLOW;        // (cannot happen with the windows the tiers use)
  TileSet done, todo;
  done.clear(); todo.clear();
  {
    const int ctx = cu >> 5, cty = cv >> 5;
    for (int ty = cty - A.halo; ty <= cty + A.halo; ty++)
      for (int tx = ctx - A.halo; tx <= ctx + A.halo; tx++)
        if (tx >= 0 && ty >= 0 && tx < nt && ty < nt) todo.set(ty * nt + tx);
  }
  while (todo.any()) {
    todo.for_each([&](int t) { nbr_sample_tile(g, c, vol, A, t % nt, t / nt); });
#pragma unroll
    for (int i = 0; i < NW; i++) done.w[i] |= todo.w[i];
    g.sync();
    TileSet want;
    want.clear();
    int k = 0;
    todo.for_each([&](int t) {
      if ((k++ % g.nwarps()) == g.warp()) nbr_grow_tile(g, A, t % nt, t / nt, want);
    });
    if (g.lane() == 0) {
#pragma unroll
      for (int i = 0; i < NW; i++)
        if (want.w[i] & ~done.w[i]) atomic_or_u32(&ctl.grow[i], want.w[i] & ~done.w[i]);
    }
    g.sync();
#pragma unroll
    for (int i = 0; i < NW; i++) todo.w[i] = ctl.grow[i] & ~done.w[i];   // (next write to ctl.grow: after the next sync)
  }
  if (g.tid() == 0) ctl.t[3] += (uint32_t)(xs_clock() - ctl.tmark);      // tile sampling + closure (part of phase 0)
  done.for_each([&](int t) { nbr_masks_tile(g, A, t % nt, t / nt); });
  g.sync();
  if (g.warp() == 0) {
    const long long w0 = xs_clock();
    int st;
    const int ci0 = ctl.u;
    if (!((frow(A, cv)[cu >> 5] >> (cu & 31)) & 1u)) {
      st = ST_DONE;
      if (g.lane() == 0) ctl.n = 0;
    } else {
      g.converge();
      if (g.lane() == 0) {
        A.order[0] = nbr_order_entry<CellT>(ci0, MS, A.moff); A.stack[0] = (CellT)ci0;
        for (int k = 0; k < 8; k++) nbr_visit_one(A.cells, ci0, k, MS);
      }
      g.converge();
      int n = 1, depth = 1, lo = 0, ci = ci0;
#ifdef __CUDA_ARCH__
      st = dfs_run_nbr_warp(A, n, depth, lo, ci, g.lane(), A.fbits);   // (the foreground bitmap is dead: 32 words of scratch)
#else
      st = dfs_run_nbr(A, n, depth, lo, ci);
#endif
      if (g.lane() == 0) { ctl.n = n; ctl.depth = depth; ctl.lo = lo; ctl.u = ci; }
    }
    if (g.lane() == 0) {
      ctl.status = st;
      ctl.tiles = done.count();
      ctl.t[5] += (uint32_t)(xs_clock() - w0);
    }
  }
  g.sync();
  if (ctl.status == ST_OVERFLOW) return Q_OVERFLOW;
  *empty = (ctl.n == 0);
  return Q_OK;
}

struct Sample {
  int rx, ry, rz;   // rounded voxel of the lattice sample
};
template <class CellT>
XS_HD Sample sample_at(const PlaneCtx& c, const Arena<CellT>& A, int r, V3* rounded) {
  typedef CellPack<CellT> P;
  const CellT p = A.order[r];
  const V3 cur = vround(cell_point(c, A.ox + P::u(p), A.oy + P::v(p)));
  Sample s;
  s.rx = (int)cur.x; s.ry = (int)cur.y; s.rz = (int)cur.z;
  if (rounded) *rounded = cur;
  return s;
}

// Contact bits of a sample from its ROUNDED point (path A, xs3d.hpp:909-914).
XS_HD uint32_t contact_bits_rounded(const PlaneCtx& c, V3 cur) {
  return (uint32_t)(cur.x < 1.0f) | ((uint32_t)(cur.x >= c.hi[0]) << 1) | ((uint32_t)(cur.y < 1.0f) << 2) |
         ((uint32_t)(cur.y >= c.hi[1]) << 3) | ((uint32_t)(cur.z < 1.0f) << 4) | ((uint32_t)(cur.z >= c.hi[2]) << 5);
}

// Which of the 27 neighbouring 2x2x2 blocks a sample probes with a SET probe voxel (bit k = (ox+1) + 3(oy+1) + 9(oz+1)):
// probe voxel r + 2o must be inside the volume (xs3d.hpp:665-671) and set (:704); axis-aligned planes only look at the
// sample's own block (:677-685).  Branch-free: an out-of-volume probe reads the sample's own block and is masked off.
XS_HD uint32_t probe_candidates(const PlaneCtx& c, const VolView& vol, const Sample& s) {
  const int bx = s.rx >> 1, by = s.ry >> 1, bz = s.rz >> 1;
  const int pb = (s.rx & 1) | ((s.ry & 1) << 1) | ((s.rz & 1) << 2);
  const uint32_t base = cube_index(vol, bx, by, bz);
  if (c.axis_aligned) return ((cube_load(vol, base) >> pb) & 1u) << 13;
  const int hx = vol.hx, hxy = vol.hx * vol.hy;
  const uint32_t okx = (s.rx - 2 >= 0 ? 1u : 0u) | 2u | (s.rx + 2 < vol.sx ? 4u : 0u);
  const uint32_t oky = (s.ry - 2 >= 0 ? 1u : 0u) | 2u | (s.ry + 2 < vol.sy ? 4u : 0u);
  const uint32_t okz = (s.rz - 2 >= 0 ? 1u : 0u) | 2u | (s.rz + 2 < vol.sz ? 4u : 0u);
  uint32_t cmask = 0u;
#pragma unroll
  for (int oz = 0; oz < 3; oz++) {
#pragma unroll
    for (int oy = 0; oy < 3; oy++) {
#pragma unroll
      for (int ox = 0; ox < 3; ox++) {
        const uint32_t ok = (okx >> ox) & (oky >> oy) & (okz >> oz) & 1u;
        const uint32_t off = (uint32_t)((ox - 1) + (oy - 1) * hx + (oz - 1) * hxy);
        const uint32_t cb = cube_load(vol, base + (ok ? off : 0u));
        cmask |= ((cb >> pb) & ok) << (ox + 3 * oy + 9 * oz);
      }
    }
  }
  return cmask;
}

// Step 3: contact bits (xs3d.hpp:909-914) and first-touch claims (xs3d.hpp:689-708).
// The owner of a block is the claimer with the smallest DFS rank; since a sample probes each block at
// most once, the rank alone identifies the (rank,k) pair the reference would have visited first.
template <class G, class CellT, class H>
XS_D void claim_blocks(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, const H& hash, Ctl& ctl) {
  const int n = ctl.n;
  hash.clear(g);
  g.sync();
  uint32_t ct = 0u;
  bool ovf = false;
  // uniform trip count + an explicit reconvergence point per chunk: the per-lane claim loop below has a
  // data-dependent length, and without the re-join the lanes of a warp drift apart for the rest of the phase
  for (int r0 = 0; r0 < n; r0 += g.size()) {
    const int r = r0 + g.tid();
    uint32_t my_cmask = 0u, my_tag = 0u;
    if (r < n) {
      V3 cur;
      const Sample s = sample_at(c, A, r, &cur);
      ct |= contact_bits_rounded(c, cur);
      // claim every probed block (below); remember only the claims that WON at claim time: a sample can end up owning
      // a block only if it held it at some point, so the ownership pass re-checks those few instead of all ~20
      my_cmask = probe_candidates(c, vol, s);
      my_tag = hash.tag(s.rx >> 1, s.ry >> 1, s.rz >> 1);
    }
    {
      uint32_t tentative = 0u;
#ifdef __CUDA_ARCH__
      const uint32_t any = __reduce_or_sync(0xffffffffu, my_cmask);
#else
      const uint32_t any = my_cmask;
#endif
      bool use_batch = false;
      if constexpr (H::kBatch) use_batch = hash.batched();
      if (use_batch) {
        // table in HBM: one z-layer (nine offsets) at a time, the nine claims overlapping their L2 round trips
#pragma unroll 1
        for (int oz = 0; oz < 3; oz++) {
          if (!((any >> (9 * oz)) & 0x1ffu)) continue;            // warp-uniform
          const uint32_t m9 = (my_cmask >> (9 * oz)) & 0x1ffu;
          if (m9) tentative |= hash.claim9(my_tag + hash.tag_delta(0, 0, oz - 1), m9, (uint32_t)r, &ovf) << (9 * oz);
        }
      } else {
        // table in shared memory: offset by offset in lock step — k is the same for every lane; the tag of block
        // b + o is linear in o, so the tag delta is carried along with two additions per step instead of decoding k
        const uint32_t dx = hash.tag_delta(1, 0, 0), dy = hash.tag_delta(0, 1, 0), dz = hash.tag_delta(0, 0, 1);
        uint32_t dt = 0u - dx - dy - dz;              // o = (-1,-1,-1)
        int k = 0;
#pragma unroll 1
        for (int oz = 0; oz < 3; oz++) {
#pragma unroll 1
          for (int oy = 0; oy < 3; oy++) {
#pragma unroll 1
            for (int ox = 0; ox < 3; ox++, k++, dt += dx) {
              if (!((any >> k) & 1u)) continue;       // warp-uniform
              if ((my_cmask >> k) & 1u) {
                const int won = hash.claim_tag(my_tag + dt, (uint32_t)r);
                if (won == 0) ovf = true;
                if (won == 2) tentative |= 1u << k;
              }
            }
            dt += dy - 3u * dx;
          }
          dt += dz - 3u * dy;
        }
      }
      if (r < n) A.mask[r] = tentative;
    }
    g.converge();
  }
  ct = g.or_reduce(ct);
  if (g.tid() == 0) ctl.contact = ct;
  if (ovf) ctl.overflow = 1;
  g.sync();
}

// polygons an item needs (xs3d.hpp:420-437): >=5 voxels set: the block polygon + one per ABSENT voxel; else one per present voxel
XS_HD int item_polys(uint32_t cube) {
  const int pop = popc8(cube);
  return pop >= 5 ? 1 + (8 - pop) : pop;
}

// Step 4: ownership -> adjacency test (xs3d.hpp:710-712) -> plane-distance test (:384-389) -> item and polygon
// records, written in (rank,k) order into ranges reserved with one atomicAdd per query.
template <class G, class CellT, class H>
XS_D int emit_items(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, const H& hash, Ctl& ctl,
                    uint32_t q, const EmitOut& out) {
  const int n = ctl.n;
  // pass 1: which candidates become items; totals.  A sample has 0-5 tentative blocks, most have one: looping per
  // sample leaves most lanes idle, so each warp first lists its (sample, offset) pairs in shared memory (the dead DFS
  // stack) and then evaluates the pairs 32 at a time; the owning sample's block and centre mask arrive by shuffle.
  int my_items = 0, my_polys = 0;
  const int pcap = A.stack_cap / g.nwarps();
  CellT* const pairs = A.stack + (size_t)g.warp() * pcap;
  for (int r0 = 0; r0 < n; r0 += g.size()) {
    const int r = r0 + g.tid();
    const uint32_t tent = (r < n) ? A.mask[r] : 0u;
    int bx = 0, by = 0, bz = 0;
    uint32_t centre = 0u;
    if (tent) {
      const Sample s = sample_at(c, A, r, (V3*)nullptr);
      bx = s.rx >> 1; by = s.ry >> 1; bz = s.rz >> 1;
      centre = cube_at(vol, bx, by, bz);
    }
    int T;
    int off = g.wexscan(popc32(tent), T);
    if (A.compact_pairs && T <= pcap) {
      if (r < n) A.mask[r] = 0u;                 // kept bits are OR-ed in below
      uint32_t t = tent;
      while (t) {
        const int k = ffs32(t) - 1;
        t &= t - 1u;
        pairs[off++] = (CellT)((g.lane() << 5) | k);
      }
      g.converge();
      for (int j0 = 0; j0 < T; j0 += g.wsize()) {
        const int j = j0 + g.lane();
        const bool act = j < T;
        const int pr = act ? (int)pairs[j] : 0;
        const int sl = pr >> 5, k = pr & 31;
        const int sbx = g.shfl(bx, sl), sby = g.shfl(by, sl), sbz = g.shfl(bz, sl);
        const uint32_t scentre = (uint32_t)g.shfl((int)centre, sl);
        if (act) {
          const int rr = r - g.lane() + sl;
          int ox, oy, oz;
          decode_k(k, ox, oy, oz);
          bool keep = hash.owner(sbx + ox, sby + oy, sbz + oz) == (uint32_t)rr;       // else: someone with a smaller rank touched it first
          uint32_t cand = 0u;
          if (keep) {
            cand = cube_at(vol, sbx + ox, sby + oy, sbz + oz);
            keep = blocks_adjacent(scentre, cand, ox, oy, oz);                         // else: visited, but contributes nothing
          }
          if (keep) {
            const V3 ctr = vadd(mk((float)(2 * (sbx + ox)), (float)(2 * (sby + oy)), (float)(2 * (sbz + oz))), mk(0.5f, 0.5f, 0.5f));
            keep = !(fabsf(vdot(vsub(ctr, c.g.pos), c.g.n)) > XS_FAR2);               // else: block value is exactly 0
          }
          if (keep) {
            atomic_or_u32(&A.mask[rr], 1u << k);
            my_items++;
            my_polys += item_polys(cand);
          }
        }
      }
    } else if (r < n) {
      // more pairs than the scratch holds: one sample per lane
      uint32_t cm = tent, keep = 0u;
      while (cm) {
        const int k = ffs32(cm) - 1;
        cm &= cm - 1u;
        int ox, oy, oz;
        decode_k(k, ox, oy, oz);
        if (hash.owner(bx + ox, by + oy, bz + oz) != (uint32_t)r) continue;
        const uint32_t cand = cube_at(vol, bx + ox, by + oy, bz + oz);
        if (!blocks_adjacent(centre, cand, ox, oy, oz)) continue;
        const V3 ctr = vadd(mk((float)(2 * (bx + ox)), (float)(2 * (by + oy)), (float)(2 * (bz + oz))), mk(0.5f, 0.5f, 0.5f));
        if (fabsf(vdot(vsub(ctr, c.g.pos), c.g.n)) > XS_FAR2) continue;
        keep |= 1u << k;
        my_items++;
        my_polys += item_polys(cand);
      }
      A.mask[r] = keep;
    }
    g.converge();
  }
  g.sync();
  if (g.tid() == 0) ctl.t[4] += (uint32_t)(xs_clock() - ctl.tmark);      // emit pass 1 (part of phase 2)
  // reserve the output ranges
  int m, npoly;
  {
    int t1, t2;
    g.exscan(my_items, t1);
    g.exscan(my_polys, t2);
    m = t1; npoly = t2;
  }
  if (g.tid() == 0) {
    ctl.m = m; ctl.npoly = npoly;
    ctl.item_off = m ? atomic_add_u32(&out.counters[1], (uint32_t)m) : 0u;
    ctl.poly_off = npoly ? atomic_add_u32(&out.counters[0], (uint32_t)npoly) : 0u;
    if ((uint64_t)ctl.item_off + (uint64_t)m > out.item_cap || (uint64_t)ctl.poly_off + (uint64_t)npoly > out.poly_cap) {
      ctl.overflow = 2;
      out.counters[2] = 1u;
    }
  }
  g.sync();
  if (ctl.overflow) return Q_CAPACITY;
  const uint32_t item_off = ctl.item_off, poly_off = ctl.poly_off;
  // pass 2, one-warp groups with scratch: the kept (sample, offset) pairs of 32 consecutive ranks, listed in (rank, k)
  // order, are the items; lanes take items 32 at a time (a sample keeps 0-3 items: one loop per sample leaves most
  // lanes idle), polygon offsets by a warp scan
  if (g.nwarps() == 1 && A.pairs2 != nullptr && A.pairs2_cap >= 27 * g.wsize()) {
    int ibase = 0, pbase = 0;
    const int ws = g.wsize();
    for (int r0 = 0; r0 < n; r0 += ws) {
      const int r = r0 + g.lane();
      const uint32_t keep = (r < n) ? A.mask[r] : 0u;
      int bx = 0, by = 0, bz = 0;
      if (keep) {
        const Sample s = sample_at(c, A, r, (V3*)nullptr);
        bx = s.rx >> 1; by = s.ry >> 1; bz = s.rz >> 1;
      }
      int T;
      int off = g.wexscan(popc32(keep), T);
      uint32_t t = keep;
      while (t) {
        const int k = ffs32(t) - 1;
        t &= t - 1u;
        A.pairs2[off++] = (uint16_t)((g.lane() << 5) | k);
      }
      g.converge();
      for (int j0 = 0; j0 < T; j0 += ws) {
        const int j = j0 + g.lane();
        const bool act = j < T;
        const int pr2 = act ? (int)A.pairs2[j] : 0;
        const int sl = pr2 >> 5, k = pr2 & 31;
        const bool first = act && (j == 0 || ((int)A.pairs2[j - 1] >> 5) != sl);
        const int sbx = g.shfl(bx, sl), sby = g.shfl(by, sl), sbz = g.shfl(bz, sl);
        int ox, oy, oz;
        decode_k(k, ox, oy, oz);
        const int cx = sbx + ox, cy = sby + oy, cz = sbz + oz;
        const uint32_t cand = act ? cube_at(vol, cx, cy, cz) : 0u;
        const int pop = popc8(cand);
        const int cnt = act ? item_polys(cand) : 0;
        int tp;
        uint32_t po = poly_off + (uint32_t)(pbase + g.wexscan(cnt, tp));
        if (act) {
          ItemRec it;
          it.poly = po;
          it.meta = (uint32_t)cnt | (pop >= 5 ? 16u : 0u) | (first ? 32u : 0u);
          out.items[item_off + (uint32_t)(ibase + j)] = it;
          PolyRec pr;
          pr.q = q;
          if (pop >= 5) {
            pr.x = (uint16_t)(2 * cx); pr.y = (uint16_t)(2 * cy); pr.z = (uint16_t)(2 * cz); pr.kind = 2;
            out.polys[po++] = pr;
          }
          const uint32_t want = (pop >= 5) ? (~cand & 0xffu) : cand;   // absent voxels are subtracted, present ones added
          for (int b = 0; b < 8; b++) {
            if (!((want >> b) & 1u)) continue;
            pr.x = (uint16_t)(2 * cx + (b & 1)); pr.y = (uint16_t)(2 * cy + ((b >> 1) & 1)); pr.z = (uint16_t)(2 * cz + (b >> 2)); pr.kind = 1;
            out.polys[po++] = pr;
          }
        }
        pbase += tp;
      }
      ibase += T;
      g.converge();
    }
    return Q_OK;
  }
  // pass 2: write, one chunk of g.size() consecutive ranks at a time
  int ibase = 0, pbase = 0;
  const int gs = g.size();
  for (int r0 = 0; r0 < n; r0 += gs) {
    const int r = r0 + g.tid();
    const uint32_t keep = (r < n) ? A.mask[r] : 0u;
    int bx = 0, by = 0, bz = 0, ni = 0, np = 0;
    if (keep) {
      const Sample s = sample_at(c, A, r, (V3*)nullptr);
      bx = s.rx >> 1; by = s.ry >> 1; bz = s.rz >> 1;
      uint32_t km = keep;
      while (km) {
        const int k = ffs32(km) - 1;
        km &= km - 1u;
        ni++;
        int ox, oy, oz;
        decode_k(k, ox, oy, oz);
        np += item_polys(cube_at(vol, bx + ox, by + oy, bz + oz));
      }
    }
    int ti, tp;
    uint32_t io = item_off + (uint32_t)(ibase + g.exscan(ni, ti));
    uint32_t po = poly_off + (uint32_t)(pbase + g.exscan(np, tp));
    uint32_t km = keep;
    bool first = true;
    while (km) {
      const int k = ffs32(km) - 1;
      km &= km - 1u;
      int ox, oy, oz;
      decode_k(k, ox, oy, oz);
      const int cx = bx + ox, cy = by + oy, cz = bz + oz;
      const uint32_t cand = cube_at(vol, cx, cy, cz);
      const int pop = popc8(cand);
      const int cnt = item_polys(cand);
      ItemRec it;
      it.poly = po;
      it.meta = (uint32_t)cnt | (pop >= 5 ? 16u : 0u) | (first ? 32u : 0u);
      out.items[io++] = it;
      first = false;
      PolyRec pr;
      pr.q = q;
      if (pop >= 5) {
        pr.x = (uint16_t)(2 * cx); pr.y = (uint16_t)(2 * cy); pr.z = (uint16_t)(2 * cz); pr.kind = 2;
        out.polys[po++] = pr;
      }
      const uint32_t want = (pop >= 5) ? (~cand & 0xffu) : cand;   // absent voxels are subtracted, present ones added
      for (int b = 0; b < 8; b++) {
        if (!((want >> b) & 1u)) continue;
        pr.x = (uint16_t)(2 * cx + (b & 1)); pr.y = (uint16_t)(2 * cy + ((b >> 1) & 1)); pr.z = (uint16_t)(2 * cz + (b >> 2)); pr.kind = 1;
        out.polys[po++] = pr;
      }
    }
    ibase += ti; pbase += tp;
    g.converge();
  }
  return Q_OK;
}

// Path A, query side: everything up to the polygon records.  Every thread returns the same status.
// On Q_OK thread 0 has filled qinfo[q] (QS_READY) and geom[q].
template <class G, class CellT, class H>
XS_D int run_area_emit(const G& g, const PlaneCtx& c, const VolView& vol, const Arena<CellT>& A, H& hash, Ctl& ctl,
                       uint32_t q, const EmitOut& out) {
  bool empty;
  const int fst = A.cells ? flood_component_nbr(g, c, vol, A, ctl, &empty) : flood_component(g, c, vol, A, ctl, &empty);
  if (fst != Q_OK) return Q_OVERFLOW;
  if (g.tid() == 0) prof_mark(ctl, 0);
  if (empty) {
    if (g.tid() == 0) {
      QInfo qi; qi.item_off = 0; qi.n_items = 0; qi.contact = 0; qi.status = out.ready;
      out.qinfo[q] = qi;
    }
    return Q_OK;
  }
  hash.fit(ctl.n);
  Arena<CellT> B = A;
  if (A.mask_alt != nullptr && ctl.n <= A.mask_alt_cap) { B.mask = A.mask_alt; B.compact_pairs = 1; }
  claim_blocks(g, c, vol, B, hash, ctl);
  if (ctl.overflow) return Q_OVERFLOW;
  if (g.tid() == 0) prof_mark(ctl, 1);
  const int st = emit_items(g, c, vol, B, hash, ctl, q, out);
  if (st != Q_OK) return st;
  if (g.tid() == 0) {
    prof_mark(ctl, 2);
    out.geom[q] = c.g;
    QInfo qi; qi.item_off = ctl.item_off; qi.n_items = (uint32_t)ctl.m; qi.contact = ctl.contact; qi.status = out.ready;
    out.qinfo[q] = qi;
  }
  return Q_OK;
}

// Step 5: one polygon record -> its anisotropy-scaled area.
XS_HD float poly_eval(const PolyRec& r, const PlaneGeom& g) {
  return cube_polygon_area((int)r.kind, (float)r.x, (float)r.y, (float)r.z, g);
}

// Step 6a: value of one item from the values of its polygons (xs3d.hpp:420-437), in record order.
XS_HD float item_value(const ItemRec& it, const float* vals) {
  const int cnt = (int)(it.meta & 15u);
  float a = 0.0f;
  if (it.meta & 16u) {
    a = vals[it.poly];
    if (a == 0.0f) return a;
    for (int j = 1; j < cnt; j++) a = fsub(a, vals[it.poly + j]);
  } else {
    for (int j = 0; j < cnt; j++) a = fadd(a, vals[it.poly + j]);
  }
  return a;
}

// Step 6b: two-level float32 sum in the reference's order (xs3d.hpp:663,713 then :848,952) over item
// values `iv[0..n)` whose `first` flag marks the first item of each lattice sample.
struct OrderedSum {
  float total, sub;
  XS_HD void init() { total = 0.0f; sub = 0.0f; }
  XS_HD void add(float v, bool first) {
    if (first) { total = fadd(total, sub); sub = 0.0f; }
    sub = fadd(sub, v);
  }
  XS_HD float finish() { return fadd(total, sub); }
};

// Lattice window that provably contains every in-volume lattice cell of the plane: the projection
// of the volume box onto (b1,b2), padded, clipped to the lattice plus one ring of outside cells.
XS_HD void bbox_window(const PlaneCtx& c, const VolView& v, int* ox, int* oy, int* tx, int* ty) {
  float mn[2] = { 0.0f, 0.0f }, mx[2] = { 0.0f, 0.0f };
  for (int k = 0; k < 8; k++) {
    const float qx = ((k & 1) ? (float)v.sx - 0.5f : -0.5f) - c.g.pos.x;
    const float qy = ((k & 2) ? (float)v.sy - 0.5f : -0.5f) - c.g.pos.y;
    const float qz = ((k & 4) ? (float)v.sz - 0.5f : -0.5f) - c.g.pos.z;
    const float a = qx * c.b1.x + qy * c.b1.y + qz * c.b1.z;
    const float b = qx * c.b2.x + qy * c.b2.y + qz * c.b2.z;
    mn[0] = a < mn[0] ? a : mn[0]; mx[0] = a > mx[0] ? a : mx[0];
    mn[1] = b < mn[1] ? b : mn[1]; mx[1] = b > mx[1] ? b : mx[1];
  }
  // b1,b2 are unit to ~1e-7, so |true - projected| << 1 cell; pad by 3 cells + 0.1%
  int lo[2], hi[2];
  for (int a = 0; a < 2; a++) {
    const float pad = 3.0f + 0.001f * (mx[a] - mn[a]);
    lo[a] = c.c + (int)floorf(mn[a] - pad);
    hi[a] = c.c + (int)ceilf(mx[a] + pad);
    if (lo[a] < -1) lo[a] = -1;
    if (hi[a] > c.psx) hi[a] = c.psx;
  }
  *ox = lo[0]; *oy = lo[1];
  *tx = (hi[0] - lo[0] + 1 + 31) / 32;
  *ty = (hi[1] - lo[1] + 1 + 31) / 32;
}

}  // namespace xs
