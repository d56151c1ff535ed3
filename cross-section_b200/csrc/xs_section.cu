// xs_section.cu — path B (xs3d.cross_section, method 0) on the whole GPU.
//
// Replaces (file:line in /root/reference/src): cross_section xs3d.hpp:1367-1426, cross_sectional_area_helper
// :1118-1268 (the lattice flood fill), calc_area_at_point :442-535 (27-neighbour voxel visit with first-touch marking).
//
// Path B is order-free: a voxel's value is its own plane∩cube polygon, whichever lattice sample touches it first, and
// the contact bits are an OR.  So nothing of the reference's sequential flood fill has to survive — only its RESULT, the
// 8-connected lattice component S of the centre cell.  The section is therefore computed by grid-wide kernels, whatever
// its size (a 500^3 sphere cut through its centre has 1.3e5 samples; a neurite section a few hundred):
//   1. lattice_sample_kernel : the whole in-volume window of the plane lattice (the projection of the volume box,
//      bbox_window) sampled into a foreground bitmap F — one warp per 32 cells, one ballot per word;
//   2. run labelling         : the maximal runs of set bits of every lattice row are the nodes of a UNION-FIND;
//      runs_count_kernel + runs_scan_kernel number them (run id = number of run starts before it, row-major);
//      runs_union_kernel links every run with the runs of the row above that touch it 8-connectedly (column ranges
//      overlapping after widening by one cell) — lock-free union by smaller root (atomicCAS), finds with path halving;
//   3. component_kernel      : the runs whose root is the root of the centre cell's run are painted into the component
//      bitmap S (and counted);
//   4. section_voxels_kernel : one warp per word of S, one lane per sample: contact bits of the unrounded sample
//      (xs3d.hpp:1215-1220), its <= 27 neighbouring voxels (:458-510), first touch decided by an atomicOr on a
//      one-bit-per-voxel visited set in HBM (the GPU counterpart of the reference's `ccl` bit vector); fresh
//      foreground voxels become polygon records;
//   5. poly_eval_kernel (xs_device.cu) values them; section_compact_kernel keeps the non-zero ones (xs3d.hpp:526-528)
//      as (linear index, value) pairs and clears the visited bits again (the set is all-zero between calls, so it is
//      never memset: the reference's PersistentShapeManager idea, xs3d.hpp:50-96).
// Work: the window (<= 1.5 * psx^2 / ... cells, 3 M for 512^3) costs ~50 instructions per cell once; everything after
// scales with the section.  HBM traffic: F and S bitmaps (window / 8 bytes each, L2 resident), one visited word + one
// occupancy byte per probed voxel, 24 B per emitted voxel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>

#include "xs_device.h"
#include "../../include/xs3d_b200.h"

using namespace xs;

struct Lattice {
  PlaneCtx c;
  VolView vol;
  int ox, oy;          // lattice cell of window cell (0, 0)
  int rows, words;     // window: rows x (32 * words) cells
  int stride;          // words per bitmap row = words + 2 (a zero pad word either side); one zero pad row above and below
  uint32_t* F;         // foreground bitmap, (rows + 2) * stride words; row r starts at F + (r + 1) * stride + 1
  uint32_t* S;         // component bitmap, same layout
};
__device__ __forceinline__ uint32_t* lat_row(uint32_t* base, const Lattice& L, int r) { return base + (size_t)(r + 1) * L.stride + 1; }

__global__ void __launch_bounds__(256) lattice_sample_kernel(Lattice L) {
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)L.rows * L.words;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += nwarps) {
    const int r = (int)(wi / L.words), w = (int)(wi - (size_t)r * L.words);
    uint32_t cube, bit;
    const bool ok = cell_probe(L.c, L.vol, L.ox + 32 * w + lane, L.oy + r, &cube, &bit);
    const uint32_t cb = cube_load(L.vol, cube);
    const uint32_t word = __ballot_sync(0xffffffffu, ok && ((cb >> bit) & 1u));
    if (lane == 0) lat_row(L.F, L, r)[w] = word;
  }
}

// ---- run numbering ------------------------------------------------------------------------------------
// run starts of padded word index i (the pad words are zero, so a row never continues the previous one)
__device__ __forceinline__ uint32_t run_starts(const uint32_t* __restrict__ F, size_t i) {
  const uint32_t f = F[i];
  return f & ~((f << 1) | (F[i - 1] >> 31));
}
constexpr int SCAN_BLOCK = 2048;      // words per CTA of runs_count_kernel (256 threads x 8)

// per word: exclusive count of run starts inside its block of SCAN_BLOCK words; per block: its total
__global__ void __launch_bounds__(256) runs_count_kernel(const uint32_t* __restrict__ F, size_t first, size_t nwords, uint32_t* __restrict__ local,
                                                         uint32_t* __restrict__ block_total) {
  __shared__ uint32_t warp_sum[8];
  const size_t base = first + (size_t)blockIdx.x * SCAN_BLOCK + (size_t)threadIdx.x * 8;
  uint32_t cnt[8], mine = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const size_t i = base + j;
    cnt[j] = (i < first + nwords) ? (uint32_t)__popc(run_starts(F, i)) : 0u;
    mine += cnt[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) { const uint32_t s = warp_sum[k]; if (k < warp) before += s; total += s; }
  uint32_t run = before + inc - mine;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const size_t i = base + j;
    if (i < first + nwords) local[i] = run;
    run += cnt[j];
  }
  if (threadIdx.x == 0) block_total[blockIdx.x] = total;
}
// exclusive scan of the block totals in place (one CTA); total[nblocks] = number of runs
__global__ void __launch_bounds__(1024) runs_scan_kernel(uint32_t* __restrict__ block_total, uint32_t nblocks) {
  __shared__ uint32_t part[1024];
  const uint32_t per = (nblocks + 1023u) / 1024u;
  const uint32_t lo = min(nblocks, threadIdx.x * per), hi = min(nblocks, lo + per);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; i++) s += block_total[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 1024; i++) { const uint32_t t = part[i]; part[i] = run; run += t; }
    block_total[nblocks] = run;
  }
  __syncthreads();
  s = part[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) { const uint32_t t = block_total[i]; block_total[i] = s; s += t; }
}

struct Runs {
  const uint32_t* F;
  const uint32_t* local;        // per padded word
  const uint32_t* block_base;   // per SCAN_BLOCK words (counted from `first`)
  size_t first;                 // padded index of the first word that was counted
  uint32_t* parent;
};
// id of the run that contains bit `b` of padded word i (the bit must be set)
__device__ __forceinline__ uint32_t run_of(const Runs& R, size_t i, int b) {
  const uint32_t st = run_starts(R.F, i) & (0xffffffffu >> (31 - b));      // run starts at or below bit b
  return R.block_base[(i - R.first) / SCAN_BLOCK] + R.local[i] + (uint32_t)__popc(st) - 1u;   // (-1: may have started in an earlier word)
}
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  for (;;) {
    const uint32_t p = *reinterpret_cast<volatile uint32_t*>(&parent[x]);
    if (p == x) return x;
    const uint32_t g = *reinterpret_cast<volatile uint32_t*>(&parent[p]);
    if (g != p) parent[x] = g;                 // path halving (a plain store: any ancestor is a valid parent)
    x = g;
  }
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a); b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }       // the larger root goes under the smaller one
    const uint32_t old = atomicCAS(&parent[a], a, b);
    if (old == a) return;
  }
}

__global__ void __launch_bounds__(256) runs_init_kernel(uint32_t* __restrict__ parent, const uint32_t* __restrict__ nruns) {
  const uint32_t n = *nruns;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) parent[i] = i;
}

// one thread per word of the window: for every run that STARTS in it, the runs of the row above that touch it
__global__ void __launch_bounds__(256) runs_union_kernel(Lattice L, Runs R) {
  const size_t total = (size_t)L.rows * L.words;
  for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < total; wi += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(wi / L.words), w = (int)(wi - (size_t)r * L.words);
    if (r == 0) continue;
    const size_t i = (size_t)(r + 1) * L.stride + 1 + w;
    uint32_t st = run_starts(R.F, i);
    while (st) {
      const int b0 = __ffs(st) - 1;
      st &= st - 1u;
      const uint32_t me = run_of(R, i, b0);
      // column range of the run: [a, e], in cells of the window row
      const int a = 32 * w + b0;
      int e;
      {
        size_t j = i;
        uint32_t rest = ~R.F[j] & (0xfffffffeu << b0);       // first clear bit above b0 (b0 = 31: none in this word)
        if (b0 == 31) rest = 0u;
        int wj = w;
        while (rest == 0u && R.F[j + 1] & 1u) { j++; wj++; rest = ~R.F[j]; }    // the pad word ends the last run of a row
        e = rest ? 32 * wj + __ffs(rest) - 2 : 32 * wj + 31;
      }
      // the row above, cells [a - 1, e + 1]: every maximal group of set bits in that range is (part of) one run
      const int lo = a - 1, hi = e + 1;
      const size_t up = i - (size_t)L.stride - w;           // padded index of word 0 of the row above
      uint32_t carry = 0u;
      for (int ww = (lo < 0 ? 0 : lo >> 5); ww <= (hi >> 5) && ww < L.words; ww++) {
        uint32_t m = R.F[up + ww];
        const int c0 = 32 * ww;
        if (lo > c0) m &= 0xffffffffu << (lo - c0);
        if (hi < c0 + 31) m &= 0xffffffffu >> (c0 + 31 - hi);
        uint32_t gs = m & ~((m << 1) | carry);               // group starts inside the range
        carry = m >> 31;
        while (gs) {
          const int b = __ffs(gs) - 1;
          gs &= gs - 1u;
          uf_union(R.parent, me, run_of(R, up + ww, b));
        }
      }
    }
  }
}

// paint the runs of the centre cell's component into S; result[0] = samples, result[1] = 1 when the centre cell is foreground
__global__ void __launch_bounds__(256) component_kernel(Lattice L, Runs R, int cu, int cv, uint32_t* __restrict__ result) {
  const size_t ci = (size_t)(cv + 1) * L.stride + 1 + (cu >> 5);
  if (!((R.F[ci] >> (cu & 31)) & 1u)) return;              // the reference returns zeros (xs3d.hpp:1395-1412)
  const uint32_t want = uf_find(R.parent, run_of(R, ci, cu & 31));
  if (blockIdx.x == 0 && threadIdx.x == 0) result[1] = 1u;
  const size_t total = (size_t)L.rows * L.words;
  uint32_t count = 0;
  for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < total; wi += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(wi / L.words), w = (int)(wi - (size_t)r * L.words);
    const size_t i = (size_t)(r + 1) * L.stride + 1 + w;
    uint32_t st = run_starts(R.F, i);
    while (st) {
      const int b0 = __ffs(st) - 1;
      st &= st - 1u;
      if (uf_find(R.parent, run_of(R, i, b0)) != want) continue;
      // paint the run, word by word (several runs may share a word: atomicOr)
      size_t j = i;
      uint32_t f = R.F[j];
      uint32_t rest = (b0 == 31) ? 0u : (~f & (0xfffffffeu << b0));
      uint32_t bits = rest ? (((1u << (__ffs(rest) - 1)) - 1u) & (0xffffffffu << b0)) : (0xffffffffu << b0);
      for (;;) {
        atomicOr(&L.S[j], bits);
        count += (uint32_t)__popc(bits);
        if (rest != 0u || !(R.F[j + 1] & 1u)) break;
        j++;
        f = R.F[j];
        rest = ~f;
        bits = rest ? ((1u << (__ffs(rest) - 1)) - 1u) : 0xffffffffu;
      }
    }
  }
  count = __reduce_add_sync(0xffffffffu, count);
  if ((threadIdx.x & 31) == 0 && count) atomicAdd(&result[0], count);
}

// ---- samples -> voxels --------------------------------------------------------------------------------
struct VoxelOut {
  PolyRec* rec;
  uint32_t* counter;     // records emitted
  uint32_t cap;
  uint32_t* contact;
  uint32_t* overflow;
  uint32_t* visited;     // one bit per voxel (linear F-order index), all zero between calls
};

__global__ void __launch_bounds__(256) section_voxels_kernel(Lattice L, VoxelOut o) {
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)L.rows * L.words;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const PlaneCtx& c = L.c;
  const VolView& vol = L.vol;
  const float fsx = (float)vol.sx, fsy = (float)vol.sy, fsz = (float)vol.sz;
  uint32_t ct = 0u;
  for (size_t wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < total; wi += nwarps) {
    const int r = (int)(wi / L.words), w = (int)(wi - (size_t)r * L.words);
    const uint32_t sw = lat_row(L.S, L, r)[w];
    if (sw == 0u) continue;                                   // warp-uniform
    const bool on = (sw >> lane) & 1u;                        // every lane stays in step (the emission below is a warp scan)
    const V3 cur = cell_point(c, L.ox + 32 * w + lane, L.oy + r);
    if (on)
      ct |= (uint32_t)(cur.x < 0.5f) | ((uint32_t)(cur.x >= c.hi[0]) << 1) | ((uint32_t)(cur.y < 0.5f) << 2) |
            ((uint32_t)(cur.y >= c.hi[1]) << 3) | ((uint32_t)(cur.z < 0.5f) << 4) | ((uint32_t)(cur.z >= c.hi[2]) << 5);
    int lo[3], hi[3];
    lo[0] = (fsub(cur.x, 1.0f) >= 0.0f) ? -1 : 0; lo[1] = (fsub(cur.y, 1.0f) >= 0.0f) ? -1 : 0; lo[2] = (fsub(cur.z, 1.0f) >= 0.0f) ? -1 : 0;
    hi[0] = (fadd(cur.x, 1.0f) < fsx) ? 1 : 0; hi[1] = (fadd(cur.y, 1.0f) < fsy) ? 1 : 0; hi[2] = (fadd(cur.z, 1.0f) < fsz) ? 1 : 0;
    if (c.axis_aligned) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0; }
    // one z-layer (up to nine voxels) at a time, each stage as a batch of independent memory operations
#pragma unroll 1
    for (int dz = -1; dz <= 1; dz++) {
      uint32_t id[9], bit[9], cube[9];
      uint32_t live = 0u;
      const bool layer = on && dz >= lo[2] && dz <= hi[2];
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const int dx = j % 3 - 1, dy = j / 3 - 1;
        id[j] = 0u; bit[j] = 0u; cube[j] = 0u;
        if (!layer || dx < lo[0] || dx > hi[0] || dy < lo[1] || dy > hi[1]) continue;
        const V3 qv = vround(vadd(mk((float)dx, (float)dy, (float)dz), cur));
        if (qv.x < 0.0f || qv.x >= fsx || qv.y < 0.0f || qv.y >= fsy || qv.z < 0.0f || qv.z >= fsz) continue;
        const int x = (int)qv.x, y = (int)qv.y, z = (int)qv.z;
        id[j] = (uint32_t)x + (uint32_t)vol.sx * ((uint32_t)y + (uint32_t)vol.sy * (uint32_t)z);
        bit[j] = (uint32_t)((x & 1) | ((y & 1) << 1) | ((z & 1) << 2));
        cube[j] = cube_index(vol, x >> 1, y >> 1, z >> 1);
        live |= 1u << j;
      }
      uint32_t cb[9], seen[9];
#pragma unroll
      for (int j = 0; j < 9; j++) cb[j] = ((live >> j) & 1u) ? cube_load(vol, cube[j]) : 0u;
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (!((cb[j] >> bit[j]) & 1u)) live &= ~(1u << j);            // background voxel
        seen[j] = ((live >> j) & 1u) ? o.visited[id[j] >> 5] : 0xffffffffu;   // plain load first: most voxels have been seen
      }
      uint32_t fresh = 0u;
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (!((live >> j) & 1u)) continue;
        const uint32_t m = 1u << (id[j] & 31u);
        if (seen[j] & m) continue;
        if (!(atomicOr(&o.visited[id[j] >> 5], m) & m)) fresh |= 1u << j;   // else someone else was first (xs3d.hpp:511-512)
      }
      // one counter update per warp and layer
      __syncwarp();
      const int mine = __popc(fresh);
      int inc = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      const int tot = __shfl_sync(0xffffffffu, inc, 31);
      if (tot == 0) continue;                                  // warp-uniform
      uint32_t base = 0u;
      if (lane == 0) base = atomicAdd(o.counter, (uint32_t)tot);
      base = __shfl_sync(0xffffffffu, base, 0);
      uint32_t slot = base + (uint32_t)(inc - mine);
#pragma unroll
      for (int j = 0; j < 9; j++) {
        if (!((fresh >> j) & 1u)) continue;
        if (slot < o.cap) {
          const uint32_t x = id[j] % (uint32_t)vol.sx, yz = id[j] / (uint32_t)vol.sx;
          PolyRec pr; pr.q = 0; pr.x = (uint16_t)x; pr.y = (uint16_t)(yz % (uint32_t)vol.sy); pr.z = (uint16_t)(yz / (uint32_t)vol.sy); pr.kind = 1;
          o.rec[slot] = pr;
        } else {
          *o.overflow = 1u;
        }
        slot++;
      }
    }
  }
  __syncwarp();
  ct = __reduce_or_sync(0xffffffffu, ct);
  if (lane == 0 && ct) atomicOr(o.contact, ct);
}

// keep the voxels with a non-zero contribution (xs3d.hpp:526-528) as (linear index, value); clear every visited bit
__global__ void __launch_bounds__(256) section_compact_kernel(const PolyRec* __restrict__ rec, const float* __restrict__ vals, const uint32_t* __restrict__ n_ptr,
                                                              uint32_t cap, int sx, int sy, uint32_t* __restrict__ visited, uint64_t* __restrict__ idx,
                                                              float* __restrict__ val, uint32_t* __restrict__ nnz) {
  uint32_t n = *n_ptr;
  if (n > cap) n = cap;
  const int lane = threadIdx.x & 31;
  for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    float v = 0.0f;
    uint32_t id = 0;
    if (i < n) {
      const PolyRec r = rec[i];
      id = (uint32_t)r.x + (uint32_t)sx * ((uint32_t)r.y + (uint32_t)sy * (uint32_t)r.z);
      atomicAnd(&visited[id >> 5], ~(1u << (id & 31u)));
      v = vals[i];
    }
    const bool keep = v > 0.0f;
    const uint32_t ball = __ballot_sync(0xffffffffu, keep);
    uint32_t base = 0;
    if (lane == 0 && ball) base = atomicAdd(nnz, (uint32_t)__popc(ball));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const uint32_t k = base + (uint32_t)__popc(ball & ((1u << lane) - 1u));
      idx[k] = (uint64_t)id; val[k] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Sparse result of xs3d.cross_section(method 0) for one plane, left on the device: *d_idx / *d_val hold *nnz entries.
int xs_section_fast_sparse(xs3d_volume* v, const float p[3], const float n[3], const float w[3],
                           uint64_t** d_idx, float** d_val, uint64_t* nnz, uint8_t* contact) {
  *nnz = 0; *contact = 0;
  XS_RET(xs_set_device(v));
  const VolView vol = xs_view(v);
  const unsigned long long voxels = (unsigned long long)vol.sx * vol.sy * vol.sz;
  if (voxels >= 0xffffffffull) return xs_fail(XS3D_ERR_UNSUPPORTED, "cross_section on volumes with >= 2^32 voxels");
  Lattice L;
  L.vol = vol;
  // seed outside the volume: zeros (xs3d.hpp:1382-1412); seed on background: decided on the device — the centre lattice
  // cell's sample point is the vertex itself, so component_kernel's test of that cell is the reference's seed test
  if (!plane_init(L.c, vol, mk(p[0], p[1], p[2]), mk(n[0], n[1], n[2]), mk(w[0], w[1], w[2]), false)) return XS3D_OK;
  int tx, ty;
  bbox_window(L.c, vol, &L.ox, &L.oy, &tx, &ty);
  L.words = tx; L.rows = 32 * ty; L.stride = tx + 2;
  const int cu = L.c.c - L.ox, cv = L.c.c - L.oy;
  if (cu < 0 || cv < 0 || cu >= 32 * tx || cv >= L.rows) return xs_fail(XS3D_ERR_CAPACITY, "cross_section: centre cell outside the lattice window");
  cudaStream_t st = xs_stream(v);
  const int sms = xs_sm_count(v);
  // workspace: F, S, local prefix, block totals, parent, small result block
  const size_t padded = (size_t)(L.rows + 2) * L.stride;
  const size_t first = (size_t)L.stride;                              // counted words: rows 0 .. rows-1 with their pads
  const size_t counted = (size_t)L.rows * L.stride;
  const uint32_t nblocks = (uint32_t)((counted + SCAN_BLOCK - 1) / SCAN_BLOCK);
  const size_t max_runs = (size_t)L.rows * (16 * (size_t)L.words) + 1;
  const size_t s_bits = al256(padded * 4), s_blk = al256(((size_t)nblocks + 1) * 4), s_par = al256(max_runs * 4);
  char* ws;
  XS_RET(xs_buf(v, 8, 3 * s_bits + s_blk + s_par + 256, (void**)&ws));
  L.F = (uint32_t*)ws; L.S = (uint32_t*)(ws + s_bits);
  uint32_t* local = (uint32_t*)(ws + 2 * s_bits);
  uint32_t* block_total = (uint32_t*)(ws + 3 * s_bits);
  uint32_t* parent = (uint32_t*)(ws + 3 * s_bits + s_blk);
  uint32_t* res = (uint32_t*)(ws + 3 * s_bits + s_blk + s_par);      // [0] samples [1] centre fg [2] records [3] contact [4] overflow [5] nnz
  XS_CK(cudaMemsetAsync(ws, 0, 2 * s_bits, st));                     // F (pads must be zero) and S
  XS_CK(cudaMemsetAsync(res, 0, 64, st));
  const size_t nw = (size_t)L.rows * L.words;
  lattice_sample_kernel<<<(unsigned)std::min<size_t>((nw + 7) / 8, (size_t)sms * 16), 256, 0, st>>>(L);
  runs_count_kernel<<<nblocks, 256, 0, st>>>(L.F, first, counted, local, block_total);
  runs_scan_kernel<<<1, 1024, 0, st>>>(block_total, nblocks);
  Runs R;
  R.F = L.F; R.local = local; R.block_base = block_total; R.first = first; R.parent = parent;
  runs_init_kernel<<<sms * 4, 256, 0, st>>>(parent, block_total + nblocks);
  const unsigned wgrid = (unsigned)std::min<size_t>((nw + 255) / 256, (size_t)sms * 8);
  runs_union_kernel<<<wgrid, 256, 0, st>>>(L, R);
  component_kernel<<<wgrid, 256, 0, st>>>(L, R, cu, cv, res);
  XS_CK(cudaGetLastError());
  uint32_t* h = xs_pinned(v);
  XS_CK(cudaMemcpyAsync(h, res, 8, cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  const uint64_t samples = h[0];
  if (!h[1] || samples == 0) return XS3D_OK;
  // every sample introduces at most 27 voxels, bounded by the volume
  const uint64_t cap = std::min<uint64_t>(std::min<uint64_t>(27ull * samples, voxels), 0x7fffffffull);
  char* recs;
  XS_RET(xs_buf(v, 2, al256(cap * sizeof(PolyRec)) + sizeof(PlaneGeom) + 256, (void**)&recs));
  PlaneGeom* d_geom = (PlaneGeom*)(recs + al256(cap * sizeof(PolyRec)));
  float* d_vals; uint64_t* out_idx; char* outv;
  XS_RET(xs_buf(v, 0, cap * 8 + 64, (void**)&out_idx));
  XS_RET(xs_buf(v, 1, 2 * al256(cap * 4) + 64, (void**)&outv));
  d_vals = (float*)outv;
  float* out_val = (float*)(outv + al256(cap * 4));
  uint32_t* visited;
  XS_RET(xs_visited_bits(v, &visited));
  XS_CK(cudaMemcpyAsync(d_geom, &L.c.g, sizeof(PlaneGeom), cudaMemcpyHostToDevice, st));
  VoxelOut o;
  o.rec = (PolyRec*)recs; o.counter = res + 2; o.cap = (uint32_t)cap; o.contact = res + 3; o.overflow = res + 4; o.visited = visited;
  section_voxels_kernel<<<(unsigned)std::min<size_t>((nw + 7) / 8, (size_t)sms * 16), 256, 0, st>>>(L, o);
  XS_CK(cudaGetLastError());
  XS_RET(xs_launch_poly_eval(v, (const PolyRec*)recs, d_geom, d_vals, res + 2, (uint32_t)cap));
  section_compact_kernel<<<sms * 8, 256, 0, st>>>((const PolyRec*)recs, d_vals, res + 2, (uint32_t)cap, vol.sx, vol.sy, visited, out_idx, out_val, res + 5);
  XS_CK(cudaGetLastError());
  XS_CK(cudaMemcpyAsync(h, res, 32, cudaMemcpyDeviceToHost, st));
  XS_CK(cudaStreamSynchronize(st));
  if (h[4]) return xs_fail(XS3D_ERR_CAPACITY, "cross_section: voxel record buffer overflow");
  *nnz = h[5];
  *contact = (uint8_t)h[3];
  *d_idx = out_idx; *d_val = out_val;
  return XS3D_OK;
}
