// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE: compiles the kernel headers (csrc/xs_*.cuh) as plain
// C++ with a one-thread "group", so the exact query logic that runs on the GPU can be checked against
// the oracle on a machine without a GPU (pytest -m "not gpu").  Never linked into libxs3d_b200.so and
// never used by the product path.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "xs_query.cuh"
#include "xs_slice.cuh"

using namespace xs;

static void host_ingest(const uint8_t* img, int sx, int sy, int sz, std::vector<uint8_t>& cubes) {
  const int hx = (sx + 1) >> 1, hy = (sy + 1) >> 1, hz = (sz + 1) >> 1;
  cubes.assign((size_t)hx * hy * hz, 0);
  for (int z = 0; z < sz; z++)
    for (int y = 0; y < sy; y++)
      for (int x = 0; x < sx; x++)
        if (img[(size_t)x + (size_t)sx * ((size_t)y + (size_t)sy * z)])
          cubes[(size_t)(x >> 1) + (size_t)hx * ((size_t)(y >> 1) + (size_t)hy * (z >> 1))] |=
              (uint8_t)(1u << ((x & 1) | ((y & 1) << 1) | ((z & 1) << 2)));
}

static uint32_t pow2_at_least(uint64_t x) { uint32_t p = 64; while (p < x) p <<= 1; return p; }

// Output buffers of the query kernels (what lives in HBM on the device)
struct HostOut {
  std::vector<PolyRec> polys;
  std::vector<ItemRec> items;
  std::vector<PlaneGeom> geom;
  std::vector<QInfo> qinfo;
  uint32_t counters[4];
  EmitOut out;
  void setup(size_t nq, size_t poly_cap, size_t item_cap) {
    polys.assign(poly_cap, PolyRec()); items.assign(item_cap, ItemRec()); geom.assign(nq, PlaneGeom()); qinfo.assign(nq, QInfo());
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    out.polys = polys.data(); out.items = items.data(); out.geom = geom.data(); out.qinfo = qinfo.data();
    out.counters = counters; out.poly_cap = (uint32_t)poly_cap; out.item_cap = (uint32_t)item_cap; out.ready = QS_READY;
  }
};

// warp-tier-like arena: 96x96 window, uint16 cells, packed hash, fixed capacities
struct SmallArena {
  std::vector<uint32_t> bits, tested, slots;
  std::vector<uint16_t> order, stack;
  Arena<uint16_t> A;
  PackedHash H;
  void setup(const PlaneCtx& c, const Sample& centre) {
    const int tiles = 3, max_n = 384, hcap = 1024;
    A.tx = A.ty = tiles; A.ox = A.oy = c.c - 16 - 32 * (tiles / 2); A.halo = 0; A.stride = tiles + 1;
    bits.assign((size_t)tiles * 32 * A.stride, 0xdeadbeefu); tested.assign(4, 0xdeadbeefu);
    order.assign(max_n, 0); stack.assign(max_n, 0); slots.assign(hcap, 0);
    A.bits = bits.data(); A.tested = tested.data(); A.order = order.data(); A.stack = stack.data();
    A.max_n = max_n; A.stack_cap = max_n; A.spill = nullptr; A.mask = A.bits;   // masks alias the (dead) bitmap
    A.cells = nullptr; A.wside = 0; A.mside = 0; A.moff = 0; A.mlim = 0; A.fbits = nullptr; A.fstride = 0; A.nlut = nullptr; A.compact_pairs = 1; A.mask_alt = nullptr; A.mask_alt_cap = 0;
    A.pairs2 = reinterpret_cast<uint16_t*>(slots.data()); A.pairs2_cap = 2 * hcap;   // like the warp tier: the table is idle during emit pass 2
    H.slot = slots.data(); H.cap = hcap; H.shift = 32 - __builtin_ctz(hcap); H.max_probe = 64;
    H.cbx = centre.rx >> 1; H.cby = centre.ry >> 1; H.cbz = centre.rz >> 1;
  }
};

// warp-tier arena, neighbour-byte form (what the device's T0 runs): 96x96-cell bitmap with one shared pad word per
// row, 64-byte-wide neighbour rows for the 63x63 cells around the vertex; the first-touch table and the pair list of
// emit pass 2 alias the neighbour bytes, the masks alias the bitmap — exactly the device's aliasing
struct SmallNbrArena {
  std::vector<uint32_t> fbits, cellwords;
  std::vector<uint16_t> order, stack;
  std::vector<int16_t> lut;
  Arena<uint16_t> A;
  PackedHash H;
  void setup(const PlaneCtx& c, const Sample& centre) {
    const int max_n = 384, hcap = 1024, WS = 96, MS = 64, ML = 63, MO = 17, fs = WS / 32 + 1;
    fbits.assign((size_t)(WS + 2) * fs + 2, 0xdeadbeefu);
    cellwords.assign(64 + MS * MS / 4 + 64, 0xeeeeeeeeu);          // (a row of slack either side, like the device's neighbours)
    order.assign(max_n, 0); stack.assign(max_n, 0); lut.assign(256, 0);
    for (int m = 1; m < 256; m++) lut[m] = (int16_t)nbr_delta((uint32_t)m, MS);
    A.tx = A.ty = WS / 32; A.ox = A.oy = c.c - WS / 2; A.halo = 0; A.stride = 0;
    A.bits = nullptr; A.tested = nullptr;
    A.order = order.data(); A.stack = stack.data(); A.max_n = max_n; A.stack_cap = max_n; A.spill = nullptr;
    A.fbits = fbits.data(); A.fstride = fs; A.mask = A.fbits;
    A.cells = reinterpret_cast<uint8_t*>(cellwords.data() + 64); A.wside = WS; A.mside = MS; A.moff = MO; A.mlim = ML;
    A.nlut = lut.data(); A.compact_pairs = 1; A.mask_alt = nullptr; A.mask_alt_cap = 0;
    A.pairs2 = reinterpret_cast<uint16_t*>(cellwords.data() + 64); A.pairs2_cap = 2 * hcap;
    H.slot = cellwords.data() + 64; H.cap = hcap; H.shift = 32 - __builtin_ctz(hcap); H.max_probe = 64;
    H.cbx = centre.rx >> 1; H.cby = centre.ry >> 1; H.cbz = centre.rz >> 1;
  }
};

// CTA-tier-like arena: bbox window, uint32 cells, wide hash, optional tiny stack ring
struct BigArena {
  std::vector<uint32_t> bits, tested, order, stack, spill, mask, htag, hkey;
  Arena<uint32_t> A;
  WideHash H;
  void setup(const PlaneCtx& c, const VolView& v, int stack_cap, int factor) {
    bbox_window(c, v, &A.ox, &A.oy, &A.tx, &A.ty);
    A.halo = 1; A.stride = A.tx + 1;
    const int max_n = A.tx * A.ty * 1024;
    bits.assign((size_t)A.ty * 32 * A.stride, 0xdeadbeefu); tested.assign((A.tx * A.ty + 31) / 32, 0xdeadbeefu);
    order.assign(max_n, 0); mask.assign(max_n, 0);
    A.max_n = max_n;
    if (stack_cap > 0) { stack.assign(stack_cap, 0); spill.assign(max_n, 0); A.stack_cap = stack_cap; A.spill = spill.data(); }
    else { stack.assign(max_n, 0); A.stack_cap = max_n; A.spill = nullptr; }
    const uint32_t hcap = pow2_at_least((uint64_t)max_n * (uint64_t)factor + 64);
    htag.assign(hcap, 0); hkey.assign(hcap, 0);
    A.bits = bits.data(); A.tested = tested.data(); A.order = order.data(); A.stack = stack.data(); A.mask = mask.data();
    A.cells = nullptr; A.wside = 0; A.mside = 0; A.moff = 0; A.mlim = 0; A.fbits = nullptr; A.fstride = 0; A.nlut = nullptr; A.compact_pairs = 1; A.mask_alt = nullptr; A.mask_alt_cap = 0;
    A.pairs2 = nullptr; A.pairs2_cap = 0;
    H.htag = htag.data(); H.hkey = hkey.data(); H.cap = hcap; H.cap_max = hcap; H.shift = 32 - __builtin_ctz(hcap);
    H.hx = v.hx; H.hy = v.hy; H.max_probe = 1 << 30; H.factor = factor;
    H.alt_tag = nullptr; H.alt_key = nullptr; H.alt_cap = 0;
  }
};

// mid-tier-like arena: neighbour-byte 256x256 window, uint16 cells, wide hash, small stack ring
struct ByteArena {
  std::vector<uint8_t> cells;
  std::vector<uint32_t> fbits;
  std::vector<int16_t> lut;
  std::vector<uint16_t> order, stack, spill;
  std::vector<uint32_t> mask, htag, hkey;
  Arena<uint16_t> A;
  WideHash H;
  void setup(const PlaneCtx& c, const VolView& v, int max_n, int ring) {
    cells.assign(65536, 0xee); order.assign(max_n, 0); stack.assign(ring, 0); spill.assign(max_n, 0); mask.assign(max_n, 0);
    const uint32_t hcap = pow2_at_least((uint64_t)max_n * 16 + 64);
    htag.assign(hcap, 0); hkey.assign(hcap, 0);
    fbits.assign((size_t)(256 + 2) * 10, 0xdeadbeefu); lut.assign(256, 0);
    for (int m = 1; m < 256; m++) lut[m] = (int16_t)nbr_delta((uint32_t)m, 256);
    A.fbits = fbits.data(); A.fstride = 10; A.nlut = lut.data(); A.compact_pairs = 0;
    A.pairs2 = nullptr; A.pairs2_cap = 0;
    A.mask_alt = fbits.data() + 32; A.mask_alt_cap = (int)fbits.size() - 32;   // like the device: small sections keep their masks in the dead bitmap
    A.cells = cells.data(); A.wside = 256; A.mside = 256; A.moff = 0; A.mlim = 256; A.ox = A.oy = c.c - 128 - 16; A.halo = 1;
    A.bits = nullptr; A.tested = nullptr; A.tx = A.ty = 8; A.stride = 0;
    A.order = order.data(); A.max_n = max_n; A.stack = stack.data(); A.stack_cap = ring; A.spill = spill.data(); A.mask = mask.data();
    H.htag = htag.data(); H.hkey = hkey.data(); H.cap = hcap; H.cap_max = hcap; H.shift = 32 - __builtin_ctz(hcap);
    H.hx = v.hx; H.hy = v.hy; H.max_probe = 1 << 30; H.factor = 16;
    // like the device: a small table aliased onto the (dead) cell window
    H.alt_tag = reinterpret_cast<uint32_t*>(cells.data()); H.alt_key = H.alt_tag + 8192; H.alt_cap = 8192;
  }
};

// wide-tier-like arena: neighbour-byte window of WS x WS cells (WS a multiple of 32, not a power of two), uint32 cells
struct WideArena {
  static constexpr int WS = 160;
  std::vector<uint8_t> cells;
  std::vector<uint32_t> fbits, order, stack, spill, mask, htag, hkey;
  std::vector<int16_t> lut;
  Arena<uint32_t> A;
  WideHash H;
  void setup(const PlaneCtx& c, const VolView& v, int max_n, int ring) {
    cells.assign((size_t)WS * WS + 1024, 0xee); order.assign(max_n, 0); stack.assign(ring, 0); spill.assign(max_n, 0); mask.assign(max_n, 0);
    const uint32_t hcap = pow2_at_least((uint64_t)max_n * 16 + 64);
    htag.assign(hcap, 0); hkey.assign(hcap, 0);
    const int fs = WS / 32 + 2;
    fbits.assign((size_t)(WS + 2) * fs, 0xdeadbeefu); lut.assign(256, 0);
    for (int m = 1; m < 256; m++) lut[m] = (int16_t)nbr_delta((uint32_t)m, WS);
    A.fbits = fbits.data(); A.fstride = fs; A.nlut = lut.data(); A.compact_pairs = 0; A.mask_alt = nullptr; A.mask_alt_cap = 0; A.pairs2 = nullptr; A.pairs2_cap = 0;
    A.cells = cells.data() + 512; A.wside = WS; A.mside = WS; A.moff = 0; A.mlim = WS; A.ox = A.oy = c.c - (WS / 64) * 32 - 16; A.halo = 1;
    A.bits = nullptr; A.tested = nullptr; A.tx = A.ty = WS / 32; A.stride = 0;
    A.order = order.data(); A.max_n = max_n; A.stack = stack.data(); A.stack_cap = ring; A.spill = spill.data(); A.mask = mask.data();
    H.htag = htag.data(); H.hkey = hkey.data(); H.cap = hcap; H.cap_max = hcap; H.shift = 32 - __builtin_ctz(hcap);
    H.hx = v.hx; H.hy = v.hy; H.max_probe = 1 << 30; H.factor = 16;
    H.alt_tag = nullptr; H.alt_key = nullptr; H.alt_cap = 0;
  }
};

extern "C" {

// mode 0: big arena only.  mode 1: small (warp-tier-like) arena first, big arena on overflow.
// mode 2: like 0 but with a tiny DFS stack ring (exercises spill/refill).
// mode 3: neighbour-byte (mid-tier-like) arena with a 64-entry ring first, big arena on overflow.
// mode 4: neighbour-byte arena with a 160-cell window and 32-bit cells (wide-tier-like) first, big arena on overflow.
// mode 5: the warp tier's neighbour-byte arena first, big arena on overflow.
// stats[0] = queries that overflowed the small arena, stats[1] = total samples, stats[2] = total items, stats[3] = polygons
void hostsim_area_batch(const uint8_t* img, uint64_t sx, uint64_t sy, uint64_t sz, uint64_t nq, const float* pos,
                        const float* nrm, const float* w, float* area, uint8_t* contact, int mode, int64_t* stats) {
  std::vector<uint8_t> cubes;
  host_ingest(img, (int)sx, (int)sy, (int)sz, cubes);
  VolView v = { cubes.data(), (int)sx, (int)sy, (int)sz, ((int)sx + 1) >> 1, ((int)sy + 1) >> 1, ((int)sz + 1) >> 1 };
  HostGroup g;
  int64_t n_over = 0, n_samples = 0;
  HostOut O;
  O.setup(nq, 1 << 22, 1 << 21);
  // ---- query kernels
  for (uint64_t q = 0; q < nq; q++) {
    PlaneCtx c;
    QInfo qi; qi.item_off = 0; qi.n_items = 0; qi.contact = 0; qi.status = QS_READY;
    O.qinfo[q] = qi;
    if (!plane_init(c, v, mk(pos[3*q], pos[3*q+1], pos[3*q+2]), mk(nrm[3*q], nrm[3*q+1], nrm[3*q+2]), mk(w[0], w[1], w[2]))) continue;
    O.qinfo[q].status = QS_PENDING;
    Ctl ctl;
    int st = Q_OVERFLOW;
    if (mode == 1) {
      SmallArena sa;
      const V3 rp = vround(c.g.pos);
      Sample ctr; ctr.rx = (int)rp.x; ctr.ry = (int)rp.y; ctr.rz = (int)rp.z;
      sa.setup(c, ctr);
      st = run_area_emit(g, c, v, sa.A, sa.H, ctl, (uint32_t)q, O.out);
      if (st != Q_OK) n_over++;
    }
    if (mode == 3) {
      ByteArena ya;
      ya.setup(c, v, 4096, 64);
      st = run_area_emit(g, c, v, ya.A, ya.H, ctl, (uint32_t)q, O.out);
      if (st != Q_OK) n_over++;
    }
    if (mode == 5) {
      SmallNbrArena sn;
      const V3 rp = vround(c.g.pos);
      Sample ctr; ctr.rx = (int)rp.x; ctr.ry = (int)rp.y; ctr.rz = (int)rp.z;
      sn.setup(c, ctr);
      st = run_area_emit(g, c, v, sn.A, sn.H, ctl, (uint32_t)q, O.out);
      if (st != Q_OK) n_over++;
    }
    if (mode == 4) {
      WideArena wa;
      wa.setup(c, v, 8192, 128);
      st = run_area_emit(g, c, v, wa.A, wa.H, ctl, (uint32_t)q, O.out);
      if (st != Q_OK) n_over++;
    }
    if (st != Q_OK) {
      BigArena ba;
      ba.setup(c, v, mode == 2 ? 16 : 0, 16);
      st = run_area_emit(g, c, v, ba.A, ba.H, ctl, (uint32_t)q, O.out);
    }
    n_samples += ctl.n;
  }
  // ---- polygon kernel
  const uint32_t np = O.counters[0];
  std::vector<float> vals(np);
  for (uint32_t i = 0; i < np; i++) vals[i] = poly_eval(O.polys[i], O.geom[O.polys[i].q]);
  // ---- combine kernel
  for (uint64_t q = 0; q < nq; q++) {
    const QInfo qi = O.qinfo[q];
    if (qi.status != QS_READY) { area[q] = -1.f; contact[q] = 0xff; continue; }
    OrderedSum s; s.init();
    for (uint32_t i = 0; i < qi.n_items; i++) {
      const ItemRec it = O.items[qi.item_off + i];
      s.add(item_value(it, vals.data()), (it.meta & 32u) != 0);
    }
    area[q] = s.finish(); contact[q] = (uint8_t)qi.contact;
  }
  if (stats) { stats[0] = n_over; stats[1] = n_samples; stats[2] = O.counters[1]; }
}

float hostsim_voxel_area(uint64_t x, uint64_t y, uint64_t z, const float* p, const float* n, const float* w) {
  PlaneGeom g;
  geom_init(g, mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), mk(w[0], w[1], w[2]));
  return voxel_area((float)x, (float)y, (float)z, g);
}

float hostsim_block_area(uint8_t cube, const float* cur, const float* p, const float* n, const float* w) {
  PlaneGeom g;
  geom_init(g, mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), mk(w[0], w[1], w[2]));
  return block_value(cube, (float)((uint64_t)cur[0] & ~1ull), (float)((uint64_t)cur[1] & ~1ull), (float)((uint64_t)cur[2] & ~1ull), g);
}

// plane_cube_miss must never drop a cube for which plane_cube_points finds a point.  Random and adversarial planes
// (through / next to corners, edges and faces, tiny and zero normal components, coordinates up to `span`).
// out[0] cases, out[1] violations, out[2] dropped by plane_cube_miss, out[3] dropped by plane_cube_far, out[4] polygons (>= 3 points)
static uint64_t hs_rng(uint64_t& s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float hs_unit(uint64_t& s) { return (float)((hs_rng(s) >> 11) * (1.0 / 9007199254740992.0)); }
void hostsim_miss_check(uint64_t n_cases, uint64_t seed, float span, int64_t* out) {
  uint64_t s = seed * 2654435761ull + 88172645463325252ull;
  for (int i = 0; i < 5; i++) out[i] = 0;
  for (uint64_t it = 0; it < n_cases; it++) {
    const int S = 1 + (int)(hs_rng(s) & 1);
    V3 n = mk(hs_unit(s) * 2 - 1, hs_unit(s) * 2 - 1, hs_unit(s) * 2 - 1);
    const int style = (int)(hs_rng(s) % 8);
    if (style == 1) n.x = 0.0f;
    if (style == 2) { n.x = 0.0f; n.y = 0.0f; }
    if (style == 3) n.z *= 1e-6f;
    if (style == 4) { n.x *= 1e-4f; n.y *= 1e-7f; }
    if (style == 5) { n.x = n.y; n.z = -n.y; }
    if (vnull(n)) n.z = 1.0f;
    n = vdivs(n, vnorm(n));
    const float x = floorf(hs_unit(s) * span), y = floorf(hs_unit(s) * span), z = floorf(hs_unit(s) * span);
    // a point of (or right next to) the cube: on a corner, an edge, a face, inside, or up to 2.5 cells away
    const float fS = (float)S;
    V3 q = mk(x - 0.5f, y - 0.5f, z - 0.5f);
    const int where = (int)(hs_rng(s) % 6);
    float fx = hs_unit(s), fy = hs_unit(s), fz = hs_unit(s);
    if (where <= 2) fx = (hs_rng(s) & 1) ? 1.0f : 0.0f;
    if (where <= 1) fy = (hs_rng(s) & 1) ? 1.0f : 0.0f;
    if (where == 0) fz = (hs_rng(s) & 1) ? 1.0f : 0.0f;
    q = vadd(q, mk(fx * fS, fy * fS, fz * fS));
    float push = 0.0f;
    const int pk = (int)(hs_rng(s) % 6);
    if (pk == 1) push = (hs_unit(s) * 2 - 1) * 1e-4f;
    if (pk == 2) push = (hs_unit(s) * 2 - 1) * 3e-2f;
    if (pk >= 4 || where == 5) push = (hs_unit(s) * 2 - 1) * 2.5f;
    // the plane point: q moved along the normal by `push` and anywhere within the plane (up to 40 cells, like a query vertex; a quarter of the cases up to 3000)
    V3 t1 = vcross(n, mk(1.0f, 0.0f, 0.0f));
    if (vnorm(t1) < 1e-3f) t1 = vcross(n, mk(0.0f, 1.0f, 0.0f));
    t1 = vdivs(t1, vnorm(t1));
    const V3 t2 = vcross(n, t1);
    const float reachab = (hs_rng(s) % 4 == 0) ? 3000.0f : 40.0f;
    const float a = (hs_unit(s) * 2 - 1) * reachab, b = (hs_unit(s) * 2 - 1) * reachab;
    const V3 pos = vadd(vadd(vadd(q, vscale(n, push)), vscale(t1, a)), vscale(t2, b));
    PlaneGeom g;
    geom_init(g, pos, n, mk(1.0f, 1.0f, 1.0f));
    Poly poly;
    plane_cube_points(S, x, y, z, g, poly);
    const bool miss = plane_cube_miss(S, x, y, z, pos, n);
    out[0]++;
    if (miss && poly.n > 0) out[1]++;
    if (miss) out[2]++;
    if (plane_cube_far(S, x, y, z, pos, n)) out[3]++;
    if (poly.n >= 3) out[4]++;
  }
}

int hostsim_is_26_connected(uint8_t centre, uint8_t cand, int x, int y, int z) { return blocks_adjacent(centre, cand, x, y, z) ? 1 : 0; }

float hostsim_round(float x) { return fround(x); }

// slice with the reference's raw layout: `out` holds psx*psx labels (zeroed by the caller), bbox as
// cross_section_projection returns it.  mode 1 forces the iterative fallback fill.
int hostsim_projection(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int itemsize, int c_order,
                       const float* p, const float* n, const float* w, int positive_basis, float crop,
                       void* out, int64_t* bbox, int mode) {
  SliceCtx s;
  slice_init(s, (int)sx, (int)sy, (int)sz, mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), mk(w[0], w[1], w[2]), positive_basis != 0, crop);
  bbox[0] = bbox[1] = bbox[2] = bbox[3] = 0;
  if (!s.pos_inside) return 0;
  bbox[0] = bbox[1] = bbox[2] = bbox[3] = s.c;
  if (!s.active) return 0;
  const int words = (s.ww + 31) / 32;
  std::vector<uint32_t> V((size_t)words * s.wh, 0), R;
  for (int y = 0; y < s.wh; y++)
    for (int x = 0; x < s.ww; x++)
      if (slice_cell(s, s.wx0 + x, s.wy0 + y, false, nullptr)) V[(size_t)y * words + (x >> 5)] |= 1u << (x & 31);
  std::vector<RowRuns> rows(s.wh);
  for (int y = 0; y < s.wh; y++) rows[y] = row_runs(&V[(size_t)y * words], words);
  const int cx = (int)(s.c - s.wx0), cy = (int)(s.c - s.wy0);
  SliceResult r = scan_rows(rows.data(), s.wh, cx, cy);
  const uint32_t* mask = V.data();
  int used_fallback = 0;
  if (r.fallback || (mode == 1 && r.y0 <= r.y1)) {
    used_fallback = 1;
    R.assign(V.size(), 0);
    R[(size_t)cy * words + (cx >> 5)] = V[(size_t)cy * words + (cx >> 5)] & (1u << (cx & 31));
    bool changed = true;
    while (changed) {
      changed = false;
      for (int y = 0; y < s.wh; y++)
        for (int x = 0; x < s.ww; x++) {
          const size_t wi = (size_t)y * words + (x >> 5);
          const uint32_t bit = 1u << (x & 31);
          if (!(V[wi] & bit) || (R[wi] & bit)) continue;
          auto get = [&](int xx, int yy) { return xx >= 0 && yy >= 0 && xx < s.ww && yy < s.wh && ((R[(size_t)yy * words + (xx >> 5)] >> (xx & 31)) & 1u); };
          if (get(x - 1, y) || get(x + 1, y) || get(x, y - 1) || get(x, y + 1)) { R[wi] |= bit; changed = true; }
        }
    }
    for (int y = 0; y < s.wh; y++) rows[y] = row_runs(&R[(size_t)y * words], words);
    r.y0 = 0x7fffffff; r.y1 = -1; r.x0 = 0x7fffffff; r.x1 = -1;
    for (int y = 0; y < s.wh; y++) {
      if (rows[y].runs == 0) continue;
      if (y < r.y0) r.y0 = y;
      r.y1 = y;
      if (rows[y].a < r.x0) r.x0 = rows[y].a;
      if (rows[y].b > r.x1) r.x1 = rows[y].b;
    }
    mask = R.data();
  }
  if (r.y0 > r.y1) return used_fallback;
  bbox[0] = s.wx0 + r.x0; bbox[1] = s.wx0 + r.x1; bbox[2] = s.wy0 + r.y0; bbox[3] = s.wy0 + r.y1;
  for (int y = r.y0; y <= r.y1; y++)
    for (int x = r.x0; x <= r.x1; x++) {
      if (!((mask[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u)) continue;
      size_t loc;
      slice_cell(s, s.wx0 + x, s.wy0 + y, c_order != 0, &loc);
      memcpy((char*)out + ((size_t)(s.wx0 + x) + (size_t)s.psx * (size_t)(s.wy0 + y)) * itemsize, (const char*)labels + loc * itemsize, itemsize);
    }
  return used_fallback;
}

long long hostsim_plane_size(uint64_t sx, uint64_t sy, uint64_t sz, const float* w, float crop) {
  return slice_plane_size((int)sx, (int)sy, (int)sz, mk(w[0], w[1], w[2]), crop);
}

}  // extern "C"
