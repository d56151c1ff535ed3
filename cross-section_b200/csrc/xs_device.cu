// xs_device.cu — sm_100a kernels + the C ABI of libxs3d_b200.so (include/xs3d_b200.h).
//
// Kernels
//   ingest_f16_kernel / ingest_generic_kernel : uint8 volume -> 2x2x2 occupancy bytes ("cubes"),
//        the only HBM-streaming kernel of the area path (reads V bytes, writes V/8).  Replaces the
//        host-side np.asfortranarray + per-query compute_cube() loads (xs3d.hpp:26-48).
//   classify_kernel / order_kernel : size estimate per query -> queries ordered largest first.
//   area_warpn_kernel (T0): persistent, one WARP per query, whole working set in shared memory (96x96-cell lattice
//        bitmap + one neighbour byte per cell around the vertex).
//   area_mid_kernel   (T1): persistent, one CTA per query, 256x256-cell window as one neighbour byte per cell
//        in shared memory; runs on a second stream beside T0 and takes what T0 hands over.
//   area_wide_kernel  (T2): the same walk with a 416x416-cell window, one CTA per SM.
//   area_block_kernel (T3/T4): whole-plane bitmap (shared memory when it fits, HBM otherwise) + DFS stack ring;
//        T4 is the same kernel with one CTA and a worst-case slab.
//   poly_eval_kernel / combine_kernel : polygon areas of the emitted records; ordered float32 sums per query.
//   (path B, xs3d.cross_section, lives in xs_section.cu: grid-wide kernels, union-find over lattice row runs.)
// The batched launch sequence (two pipelines on three streams) is in area_batch_device.
// No CPU implementation exists behind the ABI: without a device every call fails loudly.
#include <cuda.h>            // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <string>
#include <vector>
#include <mutex>

#include "xs_query.cuh"
#include "xs_device.h"
#include "../../include/xs3d_b200.h"

using namespace xs;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      char b__[512];                                                                                 \
      snprintf(b__, sizeof(b__), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return fail(XS3D_ERR_CUDA, b__);                                                               \
    }                                                                                                \
  } while (0)
#define RET(call)                   \
  do {                              \
    int r__ = (call);               \
    if (r__ != XS3D_OK) return r__; \
  } while (0)

extern "C" const char* xs3d_last_error(void) { return g_err.c_str(); }
extern "C" int xs3d_version(void) { return 100; }
extern "C" int xs3d_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(XS3D_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
  *count = n;
  return XS3D_OK;
}

// ------------------------------------------------------------------------------------------------
// ingest: uint8 voxels -> one occupancy byte per 2x2x2 block
// ------------------------------------------------------------------------------------------------
// Fast path: Fortran order, sx % 16 == 0, 16-byte aligned base.  Each thread turns 16 voxels x 4 rows
// (four 16-byte streaming loads) into 8 output bytes (one 8-byte store): fully coalesced both ways.
__device__ __forceinline__ uint32_t pair_bits(uint32_t w) {
  // bytes (v0,v1,v2,v3) -> byte0 = [v0!=0] | [v1!=0]<<1, byte2 = [v2!=0] | [v3!=0]<<1
  const uint32_t m = __vcmpne4(w, 0u) & 0x01010101u;
  return (m | (m >> 7)) & 0x00030003u;
}

__global__ void __launch_bounds__(256) ingest_f16_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ cubes,
                                                         int sx, int sy, int sz, int hx, int hy, int hz) {
  const int gx = hx >> 3;   // groups of 8 blocks along x
  const size_t total = (size_t)gx * hy * hz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xg = (int)(idx % gx);
    const size_t t = idx / gx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    const int y0 = 2 * by, z0 = 2 * bz;
    const bool y1ok = (y0 + 1) < sy, z1ok = (z0 + 1) < sz;
    const size_t row = (size_t)sx;
    const uint8_t* p00 = img + (size_t)16 * xg + row * ((size_t)y0 + (size_t)sy * z0);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p00));
    const uint4 b = y1ok ? __ldcs(reinterpret_cast<const uint4*>(p00 + row)) : zero;
    const uint4 c = z1ok ? __ldcs(reinterpret_cast<const uint4*>(p00 + row * sy)) : zero;
    const uint4 d = (y1ok && z1ok) ? __ldcs(reinterpret_cast<const uint4*>(p00 + row * sy + row)) : zero;
    uint32_t o[4];
    const uint32_t aw[4] = { a.x, a.y, a.z, a.w }, bw[4] = { b.x, b.y, b.z, b.w };
    const uint32_t cw[4] = { c.x, c.y, c.z, c.w }, dw[4] = { d.x, d.y, d.z, d.w };
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t q = pair_bits(aw[i]) | (pair_bits(bw[i]) << 2) | (pair_bits(cw[i]) << 4) | (pair_bits(dw[i]) << 6);
      o[i] = (q & 0xffu) | ((q >> 8) & 0xff00u);   // two cube bytes
    }
    uint2 out;
    out.x = o[0] | (o[1] << 16);
    out.y = o[2] | (o[3] << 16);
    *reinterpret_cast<uint2*>(cubes + (size_t)8 * xg + (size_t)hx * ((size_t)by + (size_t)hy * bz)) = out;
  }
}

// Generic path: any size / alignment / memory order, one thread per output byte.
// ZFAST selects the thread->block mapping that keeps reads contiguous for C-ordered input.
template <bool ZFAST>
__global__ void __launch_bounds__(256) ingest_generic_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ cubes,
                                                             int sx, int sy, int sz, int hx, int hy, int hz,
                                                             long long stx, long long sty, long long stz) {
  const size_t total = (size_t)hx * hy * hz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int bx, by, bz;
    if (ZFAST) { bz = (int)(idx % hz); const size_t t = idx / hz; by = (int)(t % hy); bx = (int)(t / hy); }
    else { bx = (int)(idx % hx); const size_t t = idx / hx; by = (int)(t % hy); bz = (int)(t / hy); }
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int x = 2 * bx + (b & 1), y = 2 * by + ((b >> 1) & 1), z = 2 * bz + (b >> 2);
      if (x < sx && y < sy && z < sz && __ldg(img + (long long)x * stx + (long long)y * sty + (long long)z * stz)) m |= 1u << b;
    }
    cubes[(size_t)bx + (size_t)hx * ((size_t)by + (size_t)hy * bz)] = (uint8_t)m;
  }
}

// Bit-packed image (bit i <-> voxel with linear index i != 0, packed on the host by xs_pack_bits) -> occupancy bytes.
// Fast path: Fortran order, sx % 16 == 0 — each thread turns 16 voxels x 4 rows (four 2-byte loads) into 8 output bytes.
__global__ void __launch_bounds__(256) ingest_bits_f16_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ cubes,
                                                              int sx, int sy, int sz, int hx, int hy, int hz) {
  const int gx = hx >> 3;
  const size_t total = (size_t)gx * hy * hz;
  const size_t rowb = (size_t)sx >> 3;          // bytes per x-row of the bit stream
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xg = (int)(idx % gx);
    const size_t t = idx / gx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    const int y0 = 2 * by, z0 = 2 * bz;
    const bool y1ok = (y0 + 1) < sy, z1ok = (z0 + 1) < sz;
    const uint8_t* p00 = bits + (size_t)2 * xg + rowb * ((size_t)y0 + (size_t)sy * z0);
    const uint32_t a = *reinterpret_cast<const uint16_t*>(p00);
    const uint32_t b = y1ok ? *reinterpret_cast<const uint16_t*>(p00 + rowb) : 0u;
    const uint32_t c = z1ok ? *reinterpret_cast<const uint16_t*>(p00 + rowb * sy) : 0u;
    const uint32_t d = (y1ok && z1ok) ? *reinterpret_cast<const uint16_t*>(p00 + rowb * sy + rowb) : 0u;
    uint32_t o[2] = { 0u, 0u };
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t q = ((a >> (2 * j)) & 3u) | (((b >> (2 * j)) & 3u) << 2) | (((c >> (2 * j)) & 3u) << 4) | (((d >> (2 * j)) & 3u) << 6);
      o[j >> 2] |= q << (8 * (j & 3));
    }
    *reinterpret_cast<uint2*>(cubes + (size_t)8 * xg + (size_t)hx * ((size_t)by + (size_t)hy * bz)) = make_uint2(o[0], o[1]);
  }
}
// Fortran order, sx % 32 == 0: 32 voxels x 4 rows (four 4-byte loads) -> 16 output bytes (one 16-byte store) per thread.
__device__ __forceinline__ uint32_t spread_pairs(uint32_t x) {
  // bits (p0 p1 p2 p3), two bits each, of the low byte -> pair j in bits [8j, 8j+2)
  x &= 0xffu;
  x = (x | (x << 12)) & 0x000f000fu;
  return (x | (x << 6)) & 0x03030303u;
}
__global__ void __launch_bounds__(256) ingest_bits_f32_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ cubes,
                                                              int sx, int sy, int sz, int hx, int hy, int hz) {
  const int gx = hx >> 4;
  const size_t total = (size_t)gx * hy * hz;
  const size_t rowb = (size_t)sx >> 3;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xg = (int)(idx % gx);
    const size_t t = idx / gx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    const int y0 = 2 * by, z0 = 2 * bz;
    const bool y1ok = (y0 + 1) < sy, z1ok = (z0 + 1) < sz;
    const uint8_t* p00 = bits + (size_t)4 * xg + rowb * ((size_t)y0 + (size_t)sy * z0);
    const uint32_t a = __ldcs(reinterpret_cast<const uint32_t*>(p00));
    const uint32_t b = y1ok ? __ldcs(reinterpret_cast<const uint32_t*>(p00 + rowb)) : 0u;
    const uint32_t c = z1ok ? __ldcs(reinterpret_cast<const uint32_t*>(p00 + rowb * sy)) : 0u;
    const uint32_t d = (y1ok && z1ok) ? __ldcs(reinterpret_cast<const uint32_t*>(p00 + rowb * sy + rowb)) : 0u;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
      o[k] = spread_pairs(a >> (8 * k)) | (spread_pairs(b >> (8 * k)) << 2) | (spread_pairs(c >> (8 * k)) << 4) | (spread_pairs(d >> (8 * k)) << 6);
    *reinterpret_cast<uint4*>(cubes + (size_t)16 * xg + (size_t)hx * ((size_t)by + (size_t)hy * bz)) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// Fortran order, any sx: rows of the bit stream start at any bit, so each of the four rows is cut out of three bytes.
// 16 voxels x 4 rows -> up to 8 occupancy bytes per thread.
__global__ void __launch_bounds__(256) ingest_bits_fx_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ cubes,
                                                             int sx, int sy, int sz, int hx, int hy, int hz) {
  const int gx = (hx + 7) >> 3;
  const size_t total = (size_t)gx * hy * hz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xg = (int)(idx % gx);
    const size_t t = idx / gx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    const int nv = min(16, sx - 16 * xg);            // voxels of this group that exist along x (>= 1)
    uint32_t rw[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {                    // c = dy + 2 dz
      const int y = 2 * by + (c & 1), z = 2 * bz + (c >> 1);
      uint32_t a = 0u;
      if (y < sy && z < sz) {
        const size_t first = ((size_t)z * sy + y) * sx + (size_t)16 * xg;
        const size_t byte0 = first >> 3;
        const int sh = (int)(first & 7);
        uint32_t w = __ldg(bits + byte0);
        if (sh + nv > 8) w |= (uint32_t)__ldg(bits + byte0 + 1) << 8;
        if (sh + nv > 16) w |= (uint32_t)__ldg(bits + byte0 + 2) << 16;
        a = (w >> sh) & ((1u << nv) - 1u);
      }
      rw[c] = a;
    }
    uint8_t* dst = cubes + (size_t)8 * xg + (size_t)hx * ((size_t)by + (size_t)hy * bz);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t q = ((rw[0] >> (2 * j)) & 3u) | (((rw[1] >> (2 * j)) & 3u) << 2) | (((rw[2] >> (2 * j)) & 3u) << 4) | (((rw[3] >> (2 * j)) & 3u) << 6);
      if (8 * xg + j < hx) dst[j] = (uint8_t)q;
    }
  }
}

// C order (z fastest; fastest when sz % 8 == 0 and every (x, y) row of the bit stream starts on a byte): reads want z fastest, the
// occupancy bytes x fastest, so a CTA transposes a tile through shared memory — 32 blocks in x (64 voxel rows) x one block
// in y (2 voxel rows) x 128 blocks in z (256 voxels = 32 bytes per row): 128 rows of one 32-byte sector in, 128 x 32 bytes out.
constexpr int BC_BX = 32, BC_BZ = 128, BC_PITCH = 33;   // odd pitch: the emit pass reads rows 4 apart per lane, conflict free
// second half of the transposing tiles: bit rows [2 * (local voxel x) + dy][byte along z] in shared memory -> the tile's
// 128 x 32 occupancy bytes, x fastest (32-byte segments).  One work item = one block column bxl and one byte along z
// (four blocks in z): four shared-memory bytes in, four occupancy bytes out.
__device__ __forceinline__ uint32_t gather_low_bits(uint32_t t) {   // bit 0 of byte c -> bit c
  return (((t & 0x01010101u) * 0x01020408u) >> 24) & 0xfu;
}
__device__ __forceinline__ void emit_tile_from_bit_rows(const uint8_t (*rows)[BC_PITCH], uint8_t* __restrict__ cubes,
                                                        int bx0, int by, int bz0, int hx, int hy, int hz) {
#pragma unroll
  for (int k = 0; k < (BC_BX * BC_BZ / 4) / 256; k++) {
    const int o = threadIdx.x + 256 * k;
    const int bxl = o % BC_BX, zb = o / BC_BX;
    const int bx = bx0 + bxl;
    // byte c = dx + 2 dy of R: the row of voxel column (2 bxl + dx, dy)
    const uint32_t R = (uint32_t)rows[4 * bxl][zb] | ((uint32_t)rows[4 * bxl + 2][zb] << 8) |
                       ((uint32_t)rows[4 * bxl + 1][zb] << 16) | ((uint32_t)rows[4 * bxl + 3][zb] << 24);
    if (bx < hx) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int bz = bz0 + 4 * zb + i;
        const uint32_t t = R >> (2 * i);             // per byte: bit 0 = voxel dz 0, bit 1 = voxel dz 1
        const uint32_t m = gather_low_bits(t) | (gather_low_bits(t >> 1) << 4);
        if (bz < hz) cubes[(size_t)bx + (size_t)hx * ((size_t)by + (size_t)hy * bz)] = (uint8_t)m;
      }
    }
  }
}
__global__ void __launch_bounds__(256) ingest_bits_c_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ cubes,
                                                            int sx, int sy, int sz, int hx, int hy, int hz) {
  __shared__ uint8_t rows[2 * BC_BX * 2][BC_PITCH];   // [2 * (local voxel x) + dy][byte along z]
  const int tx = (hx + BC_BX - 1) / BC_BX, tz = (hz + BC_BZ - 1) / BC_BZ;
  const size_t tiles = (size_t)tx * tz * hy;
  const size_t rowb = (size_t)sz >> 3;               // bytes per (x, y) row of the bit stream
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tzi = (int)(tile % tz);
    const size_t r = tile / tz;
    const int txi = (int)(r % tx), by = (int)(r / tx);
    const int bx0 = txi * BC_BX, bz0 = tzi * BC_BZ;
    // load: row index = 2 * (local voxel x) + dy, 32 bytes each; two threads per row, 16 bytes per thread (four 4-byte
    // loads when the rows are word aligned)
    {
      const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
      const int x = 2 * bx0 + (row >> 1), y = 2 * by + (row & 1);
      const size_t zb0 = (size_t)(2 * bz0) >> 3;     // first byte of the tile along z
      const bool ok = x < sx && y < sy;
      const uint8_t* src = bits + rowb * ((size_t)x * sy + y) + zb0 + 16 * half;
      if ((sz & 31) == 0) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = (ok && zb0 + 16 * half + 4 * k < rowb) ? __ldg(reinterpret_cast<const uint32_t*>(src) + k) : 0u;
#pragma unroll
        for (int k = 0; k < 16; k++) rows[row][16 * half + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
      } else if ((sz & 7) == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) rows[row][16 * half + k] = (ok && zb0 + 16 * half + k < rowb) ? __ldg(src + k) : (uint8_t)0;
      } else {
        // sz not a multiple of 8: a row starts at any bit of the stream — shift 17 bytes into 16 and cut the row off at sz
        const size_t first = ((size_t)x * sy + y) * sz + 2 * bz0 + 128 * half;   // first bit this thread wants
        const int nvalid = ok ? max(0, min(128, sz - (2 * bz0 + 128 * half))) : 0;
        const size_t byte0 = first >> 3;
        const int sh = (int)(first & 7);
        uint32_t lo = (nvalid > 0) ? __ldg(bits + byte0) : 0u;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          uint32_t hi = 0u, val = 0u;
          if (8 * k < nvalid) {
            if (sh + nvalid > 8 * (k + 1)) hi = __ldg(bits + byte0 + k + 1);   // a bit of that byte is wanted, so it exists
            val = ((lo >> sh) | (hi << (8 - sh))) & 0xffu;
            if (nvalid < 8 * (k + 1)) val &= (1u << (nvalid - 8 * k)) - 1u;
          }
          rows[row][16 * half + k] = (uint8_t)val;
          lo = hi;
        }
      }
    }
    __syncthreads();
    emit_tile_from_bit_rows(rows, cubes, bx0, by, bz0, hx, hy, hz);
    __syncthreads();
  }
}

// The same tile for a C-ordered uint8 image in HBM (device tensors are C-ordered), sz % 16 == 0 and a 16-byte aligned
// base: 128 rows x 256 voxels, sixteen 16-byte streaming loads per row, turned into bit rows on the way into shared memory.
// eq != 0: a voxel counts when its byte equals the one replicated in want4 (uint8 label volumes) instead of when it is non-zero.
__global__ void __launch_bounds__(256) ingest_c16_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ cubes,
                                                         int sx, int sy, int sz, int hx, int hy, int hz, uint32_t want4, int eq) {
  __shared__ uint8_t rows[2 * BC_BX * 2][BC_PITCH];
  const int tx = (hx + BC_BX - 1) / BC_BX, tz = (hz + BC_BZ - 1) / BC_BZ;
  const size_t tiles = (size_t)tx * tz * hy;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tzi = (int)(tile % tz);
    const size_t r = tile / tz;
    const int txi = (int)(r % tx), by = (int)(r / tx);
    const int bx0 = txi * BC_BX, bz0 = tzi * BC_BZ;
    uint4 q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {                    // all eight loads in flight before the first use
      const int idx = threadIdx.x + 256 * k, row = idx >> 4, seg = idx & 15;
      const int x = 2 * bx0 + (row >> 1), y = 2 * by + (row & 1), z = 2 * bz0 + 16 * seg;
      q[k] = (x < sx && y < sy && z < sz) ? __ldcs(reinterpret_cast<const uint4*>(img + ((size_t)x * sy + y) * sz + z)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int idx = threadIdx.x + 256 * k, row = idx >> 4, seg = idx & 15;
      const uint32_t w[4] = { q[k].x, q[k].y, q[k].z, q[k].w };
      uint32_t m16 = 0;
#pragma unroll
      for (int j = 0; j < 4; j++)                    // four voxel bytes -> four bits (byte j != 0 -> bit j)
        m16 |= (((((eq ? __vcmpeq4(w[j], want4) : __vcmpne4(w[j], 0u)) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * j);
      rows[row][2 * seg] = (uint8_t)m16;
      rows[row][2 * seg + 1] = (uint8_t)(m16 >> 8);
    }
    __syncthreads();
    emit_tile_from_bit_rows(rows, cubes, bx0, by, bz0, hx, hy, hz);
    __syncthreads();
  }
}

// ---- the same tile, fetched by TMA -----------------------------------------------------------------------
// ingest_c16_kernel spends most of its issue slots on moving the tile: eight 16-byte loads per thread with their
// address arithmetic and bounds tests (31.7 M warp-instructions per 512^3 volume, IPC 2.5: instruction-bound at 51 % of the
// HBM peak).  Here one elected thread asks the TMA unit for the whole 64 x 2 x 256-voxel box (cp.async.bulk.tensor.3d:
// 32 KiB, out-of-volume parts zero-filled by the hardware) into one of two shared-memory stages while the CTA turns the
// previous box into bit rows (one 16-byte shared-memory load per 16 voxels) and occupancy bytes (four table look-ups per
// four bytes: spread2[] moves the bit pair of each 2x2x2 block of a bit-row byte to bits 0 and 4 of its own byte).
// Persistent CTAs, mbarrier per stage (complete_tx), tiles taken in launch order so that neighbouring CTAs stream
// neighbouring z-tiles of the same rows.
constexpr int TMA_TILE_BYTES = 2 * BC_BX * 2 * 2 * BC_BZ;       // 64 x 2 x 256 voxels
constexpr int TMA_SMEM_BYTES = 2 * TMA_TILE_BYTES;              // dynamic part (the bit rows, the table and the barriers are static)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "XS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra XS_DONE;\n"
      "bra XS_WAIT;\n"
      "XS_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <bool EQ>
__global__ void __launch_bounds__(256) ingest_c_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* __restrict__ cubes,
                                                           int hx, int hy, int hz, uint32_t want4) {
  extern __shared__ __align__(128) uint8_t tma_smem[];          // two stages of TMA_TILE_BYTES
  __shared__ uint8_t rows[2 * BC_BX * 2][BC_PITCH];
  __shared__ uint32_t spread2[256];
  __shared__ __align__(8) unsigned long long bars[2];
  const int tx = (hx + BC_BX - 1) / BC_BX, tz = (hz + BC_BZ - 1) / BC_BZ;
  const uint32_t tiles = (uint32_t)tx * (uint32_t)tz * (uint32_t)hy;
  // spread2[b]: the bit pair (2i, 2i+1) of b to bits (0, 4) of byte i
  {
    const uint32_t b = threadIdx.x;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) v |= (((b >> (2 * i)) & 1u) | (((b >> (2 * i + 1)) & 1u) << 4)) << (8 * i);
    spread2[b] = v;
  }
  const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
  const uint32_t tile0 = smem_u32(tma_smem);
  auto coords = [&](uint32_t t, int* bx0, int* by, int* bz0) {
    const uint32_t tzi = t % (uint32_t)tz, r = t / (uint32_t)tz;
    *bx0 = (int)(r % (uint32_t)tx) * BC_BX; *by = (int)(r / (uint32_t)tx); *bz0 = (int)tzi * BC_BZ;
  };
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t t = blockIdx.x;
  if (t < tiles && threadIdx.x == 0) {
    int bx0, by, bz0;
    coords(t, &bx0, &by, &bz0);
    mbar_expect_tx(bar0, TMA_TILE_BYTES);
    tma_load_3d(tile0, &tmap, 2 * bz0, 2 * by, 2 * bx0, bar0);
  }
  const int bxl = threadIdx.x & (BC_BX - 1), zb_lo = threadIdx.x / BC_BX;         // emit: block column and first bit-row byte of this thread
  const size_t zs = (size_t)hx * hy;
  for (uint32_t it = 0; t < tiles; it++, t += gridDim.x) {
    const uint32_t st = it & 1u;
    const uint32_t tn = t + gridDim.x;
    if (tn < tiles && threadIdx.x == 0) {            // the other stage was released by the barrier that ended the previous iteration
      int bx0, by, bz0;
      coords(tn, &bx0, &by, &bz0);
      mbar_expect_tx(st ? bar0 : bar1, TMA_TILE_BYTES);
      tma_load_3d(tile0 + (st ^ 1u) * TMA_TILE_BYTES, &tmap, 2 * bz0, 2 * by, 2 * bx0, st ? bar0 : bar1);
    }
    int bx0, by, bz0;
    coords(t, &bx0, &by, &bz0);
    mbar_wait(st ? bar1 : bar0, (it >> 1) & 1u);
    const uint4* src = reinterpret_cast<const uint4*>(tma_smem + st * TMA_TILE_BYTES);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int idx = threadIdx.x + 256 * k, row = idx >> 4, seg = idx & 15;      // box layout: [x 64][y 2][z 256] = 128 rows of 256 bytes
      const uint4 q = src[idx];
      // bit 7 of every byte <- byte != 0 (== want in EQ mode): three logic / add operations per four voxels; the flags are
      // then gathered by the byte dot product (IDP.4A) with the weights 1 2 4 8 | 16 32 64 128: eight voxels -> 128 x their
      // bit row, two instructions instead of a shift, a multiply, a shift and an OR per four voxels
      auto flag4 = [&](uint32_t x) -> uint32_t {
        if (EQ) x ^= want4;
        const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
        return EQ ? (~(t | x) & 0x80808080u) : ((t | x) & 0x80808080u);
      };
      const uint32_t lo = __dp4a(flag4(q.y), 0x80402010u, __dp4a(flag4(q.x), 0x08040201u, 0u));
      const uint32_t hi = __dp4a(flag4(q.w), 0x80402010u, __dp4a(flag4(q.z), 0x08040201u, 0u));
      rows[row][2 * seg] = (uint8_t)(lo >> 7);
      rows[row][2 * seg + 1] = (uint8_t)(hi >> 7);
    }
    __syncthreads();
    if (bx0 + bxl < hx) {
      uint8_t* dst0 = cubes + (size_t)(bx0 + bxl) + (size_t)hx * (size_t)by + zs * (size_t)bz0;
#pragma unroll
      for (int k = 0; k < (BC_BX * BC_BZ / 4) / 256; k++) {
        const int zb = zb_lo + 8 * k;
        const uint32_t m = spread2[rows[4 * bxl][zb]] | (spread2[rows[4 * bxl + 2][zb]] << 1) | (spread2[rows[4 * bxl + 1][zb]] << 2) |
                           (spread2[rows[4 * bxl + 3][zb]] << 3);                  // byte i: the occupancy byte of block (bxl, 4 zb + i)
        uint8_t* dst = dst0 + zs * (size_t)(4 * zb);
        if (bz0 + 4 * zb + 3 < hz) {
          dst[0] = (uint8_t)m; dst[zs] = (uint8_t)(m >> 8); dst[2 * zs] = (uint8_t)(m >> 16); dst[3 * zs] = (uint8_t)(m >> 24);
        } else {
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (bz0 + 4 * zb + i < hz) dst[zs * i] = (uint8_t)(m >> (8 * i));
        }
      }
    }
    __syncthreads();
  }
}

// Generic path: any size / memory order, one thread per output byte.
__global__ void __launch_bounds__(256) ingest_bits_generic_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ cubes,
                                                                  int sx, int sy, int sz, int hx, int hy, int hz,
                                                                  long long stx, long long sty, long long stz) {
  const size_t total = (size_t)hx * hy * hz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int bx = (int)(idx % hx);
    const size_t t = idx / hx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int x = 2 * bx + (b & 1), y = 2 * by + ((b >> 1) & 1), z = 2 * bz + (b >> 2);
      if (x < sx && y < sy && z < sz) {
        const unsigned long long i = (unsigned long long)((long long)x * stx + (long long)y * sty + (long long)z * stz);
        if ((__ldg(bits + (i >> 3)) >> (i & 7)) & 1u) m |= 1u << b;
      }
    }
    cubes[idx] = (uint8_t)m;
  }
}

// Label volume -> occupancy bytes of ONE label: one resident segmentation serves every object in
// it without materialising (or uploading) a boolean image per object.  One thread per output byte, any item size,
// either memory order.  HBM-bound: reads V * itemsize bytes, writes V/8.
template <typename T>
__global__ void __launch_bounds__(256) ingest_label_kernel(const T* __restrict__ labels, T want, uint8_t* __restrict__ cubes,
                                                           int sx, int sy, int sz, int hx, int hy, int hz,
                                                           long long stx, long long sty, long long stz) {
  const size_t total = (size_t)hx * hy * hz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int bx = (int)(idx % hx);
    const size_t t = idx / hx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int x = 2 * bx + (b & 1), y = 2 * by + ((b >> 1) & 1), z = 2 * bz + (b >> 2);
      if (x < sx && y < sy && z < sz && __ldg(labels + (long long)x * stx + (long long)y * sty + (long long)z * stz) == want) m |= 1u << b;
    }
    cubes[idx] = (uint8_t)m;
  }
}

// Fast path of the same: Fortran order, rows of whole 16-byte groups (sx * itemsize % 16 == 0), 16-byte aligned base.
// One 16-byte streaming load per voxel row and thread (2 / 4 / 8 / 16 labels), four rows -> 1 / 2 / 4 / 8 occupancy bytes,
// stored as one item: fully coalesced both ways, the label volume is read exactly once.
template <typename T> __device__ __forceinline__ uint32_t label_eq_mask(uint4 q, T want);   // bit i <-> label i of the group == want
template <> __device__ __forceinline__ uint32_t label_eq_mask<unsigned long long>(uint4 q, unsigned long long want) {
  const uint32_t lo = (uint32_t)want, hi = (uint32_t)(want >> 32);
  return (uint32_t)(q.x == lo && q.y == hi) | ((uint32_t)(q.z == lo && q.w == hi) << 1);
}
template <> __device__ __forceinline__ uint32_t label_eq_mask<uint32_t>(uint4 q, uint32_t want) {
  return (uint32_t)(q.x == want) | ((uint32_t)(q.y == want) << 1) | ((uint32_t)(q.z == want) << 2) | ((uint32_t)(q.w == want) << 3);
}
template <> __device__ __forceinline__ uint32_t label_eq_mask<uint16_t>(uint4 q, uint16_t want) {
  const uint32_t w2 = (uint32_t)want * 0x00010001u;
  const uint32_t c[4] = { __vcmpeq2(q.x, w2), __vcmpeq2(q.y, w2), __vcmpeq2(q.z, w2), __vcmpeq2(q.w, w2) };
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) m |= ((c[j] & 1u) | ((c[j] >> 15) & 2u)) << (2 * j);
  return m;
}
template <> __device__ __forceinline__ uint32_t label_eq_mask<uint8_t>(uint4 q, uint8_t want) {
  const uint32_t w4 = (uint32_t)want * 0x01010101u;
  const uint32_t c[4] = { __vcmpeq4(q.x, w4), __vcmpeq4(q.y, w4), __vcmpeq4(q.z, w4), __vcmpeq4(q.w, w4) };
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) m |= ((((c[j] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * j);
  return m;
}
template <typename T>
__global__ void __launch_bounds__(256) ingest_label_f16_kernel(const T* __restrict__ labels, T want, uint8_t* __restrict__ cubes,
                                                               int sx, int sy, int sz, int hx, int hy, int hz) {
  constexpr int VPT = 16 / (int)sizeof(T);           // labels per 16-byte load
  constexpr int OUT = VPT / 2;                       // occupancy bytes per thread
  const int gx = hx / OUT;                           // sx % VPT == 0
  const size_t total = (size_t)gx * hy * hz;
  const size_t row = (size_t)sx;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xg = (int)(idx % gx);
    const size_t t = idx / gx;
    const int by = (int)(t % hy), bz = (int)(t / hy);
    const int y0 = 2 * by, z0 = 2 * bz;
    const bool y1ok = (y0 + 1) < sy, z1ok = (z0 + 1) < sz;
    const T* p00 = labels + (size_t)VPT * xg + row * ((size_t)y0 + (size_t)sy * z0);
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p00));
    uint4 b = a, c = a, d = a;
    if (y1ok) b = __ldcs(reinterpret_cast<const uint4*>(p00 + row));
    if (z1ok) c = __ldcs(reinterpret_cast<const uint4*>(p00 + row * sy));
    if (y1ok && z1ok) d = __ldcs(reinterpret_cast<const uint4*>(p00 + row * sy + row));
    const uint32_t ra = label_eq_mask<T>(a, want);
    const uint32_t rb = y1ok ? label_eq_mask<T>(b, want) : 0u;
    const uint32_t rc = z1ok ? label_eq_mask<T>(c, want) : 0u;
    const uint32_t rd = (y1ok && z1ok) ? label_eq_mask<T>(d, want) : 0u;
    unsigned long long o = 0ull;
#pragma unroll
    for (int j = 0; j < OUT; j++) {
      const uint32_t q = ((ra >> (2 * j)) & 3u) | (((rb >> (2 * j)) & 3u) << 2) | (((rc >> (2 * j)) & 3u) << 4) | (((rd >> (2 * j)) & 3u) << 6);
      o |= (unsigned long long)q << (8 * j);
    }
    uint8_t* dst = cubes + (size_t)OUT * xg + (size_t)hx * ((size_t)by + (size_t)hy * bz);
    if (OUT == 1) *dst = (uint8_t)o;
    else if (OUT == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)o;
    else if (OUT == 4) *reinterpret_cast<uint32_t*>(dst) = (uint32_t)o;
    else *reinterpret_cast<uint2*>(dst) = make_uint2((uint32_t)o, (uint32_t)(o >> 32));
  }
}

// C-ordered label volumes of 2 / 4 / 8-byte items (sz % 8 == 0, 16-byte aligned base): the transposing tile of
// ingest_c16_kernel.  8 labels along z make one byte of a bit row = 1 / 2 / 4 sixteen-byte groups.
template <typename T>
__global__ void __launch_bounds__(256) ingest_label_c_kernel(const T* __restrict__ labels, T want, uint8_t* __restrict__ cubes,
                                                             int sx, int sy, int sz, int hx, int hy, int hz) {
  constexpr int S = (int)sizeof(T), LPB = S / 2, VPT = 16 / S;
  __shared__ uint8_t rows[2 * BC_BX * 2][BC_PITCH];
  const int tx = (hx + BC_BX - 1) / BC_BX, tz = (hz + BC_BZ - 1) / BC_BZ;
  const size_t tiles = (size_t)tx * tz * hy;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tzi = (int)(tile % tz);
    const size_t r = tile / tz;
    const int txi = (int)(r % tx), by = (int)(r / tx);
    const int bx0 = txi * BC_BX, bz0 = tzi * BC_BZ;
    // one 16-byte load per thread and step, consecutive lanes on consecutive groups of a row (fully coalesced whatever the
    // item size); the LPB lanes that share a bit-row byte merge their bits by shuffle
    for (int k0 = 0; k0 < 16 * LPB; k0 += 8) {
      uint4 q[8];
      bool ok[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {                  // eight loads in flight before the first compare
        const int gidx = threadIdx.x + 256 * (k0 + u), unit = gidx / LPB, j = gidx % LPB, row = unit >> 5, ub = unit & 31;
        const int x = 2 * bx0 + (row >> 1), y = 2 * by + (row & 1), z = 2 * bz0 + 8 * ub;
        ok[u] = x < sx && y < sy && z < sz;
        q[u] = ok[u] ? __ldcs(reinterpret_cast<const uint4*>(labels + ((size_t)x * sy + y) * sz + z) + j) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int gidx = threadIdx.x + 256 * (k0 + u), unit = gidx / LPB, j = gidx % LPB, row = unit >> 5, ub = unit & 31;
        uint32_t m = ok[u] ? (label_eq_mask<T>(q[u], want) << (VPT * j)) : 0u;
#pragma unroll
        for (int o = 1; o < LPB; o <<= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
        if (j == 0) rows[row][ub] = (uint8_t)m;
      }
    }
    __syncthreads();
    emit_tile_from_bit_rows(rows, cubes, bx0, by, bz0, hx, hy, hz);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Second half of the CTA tiers ("P2").  A CTA-tier query (mid / wide / big) keeps an SM's shared memory only for what
// needs it — lattice tiles, neighbour bytes, the sequential walk.  When the walk is done the CTA parks the query in
// HBM pools (its rank -> cell list, a cleared first-touch table of its own, one chunk entry per 256 ranks) and takes
// the next query; claims, ownership and record emission then run as grid-wide kernels over ALL parked chunks
// (p2_claims / p2_emit1 / p2_scan / p2_emit2), hundreds of CTAs at full occupancy whatever the sizes of the individual
// sections — one 1.3e5-sample section (BASELINE config 0) is 490 chunks spread over the whole GPU instead of 20 ms on
// one SM, and in a sweep the mid tier's CTAs spend their time on walks only.
// ------------------------------------------------------------------------------------------------
struct P2Rec {               // one parked query
  uint32_t q, n;
  int32_t ox, oy;            // lattice cell of window cell (0, 0)
  uint32_t order_off;        // first entry in the order / mask pools
  uint32_t hash_off, hash_cap;   // first-touch table: tags at hash_off, ranks at hash_off + hash_cap (power of two)
  uint32_t chunk0;           // first of its ceil(n / 256) chunks
  uint32_t contact;
  uint32_t pad[3];
  PlaneCtx c;                // the query's plane, as the flood CTA set it up
};
enum { P2_CHUNK = 256,
       P2C_ORDER = 0, P2C_HASH = 1, P2C_CHUNK = 2, P2C_REC = 3, P2C_FULL = 4 /* a pool overflowed */, P2C_HASHFULL = 5 /* a table filled up */,
       P2C_DONE = 6 /* chunks the P2 kernels have finished with */, P2C_EXITED = 7 /* CTAs of the running p2_emit2 that are through */,
       P2C_WORDS = 8 };
struct P2Pools {
  uint32_t* cursors;         // P2C_*
  uint32_t* order;  uint32_t* mask;  uint32_t order_cap;
  uint32_t* hash;   uint32_t hash_cap;
  uint32_t* chunk_rec; uint32_t* chunk_items; uint32_t* chunk_polys; uint32_t* item_off; uint32_t* poly_off; uint32_t chunk_cap;
  P2Rec* recs; uint32_t rec_cap;
  int hash_factor;
};

struct BatchArgs {
  P2Pools p2;
  VolView vol;
  const float* pos;
  const float* nrm;
  float wx, wy, wz;
  uint32_t nq;
  EmitOut out;
  uint32_t* counters;        // see CNT_* below
  unsigned long long* stats; // [0] samples [1] items [2] tiles, [4..9] warp-tier phase cycles, [10..15] CTA-tier phase cycles
  uint32_t* list_in;         // queries to process (null: 0..nq-1)
  uint32_t* list_out;        // queries this tier could not fit
  uint32_t* late;            // mid tier: queries the warp tier hands over while both run (tail = counters[CNT_OVER0])
  uint32_t* done_list;       // second pipeline: queries it has finished (count = counters[CNT_DONE2]), for its combine kernel
};
enum { CNT_NEXT0 = 0, CNT_OVER0 = 1, CNT_NEXT1 = 2, CNT_OVER1 = 3, CNT_NEXT2 = 4, CNT_OVER2 = 5, CNT_NEXT3 = 6, CNT_OVER3 = 7, CNT_WORDS = 8,
       CNT_NBIG = 48,                  // queries routed straight to the mid tier = prefix of the size-ordered list
       CNT_BADN = 49,                  // a query with an all-zero normal was seen
       CNT_NEXTW = 50, CNT_OVERW = 51, // wide tier
       CNT_DONE2 = 52,                 // queries finished by the second pipeline
       CNT_HIST = 128, CNT_CURSOR = 384, CNT_TOTAL_WORDS = 640 };
#define XS_LATE_EMPTY 0xffffffffu
#define XS_LATE_CLAIMED 0xfffffffeu

__device__ __forceinline__ void finish_query(const BatchArgs& a, const Ctl& ctl, uint32_t q, int st, int cnt_over, int stats_base) {
  // thread 0 of the group
  if (st == Q_OK) {
    // CTA tiers: a non-empty query is parked (p2_handoff) and joins the done list when p2_emit2 has written its records
    if (a.done_list && (stats_base == 4 || ctl.n == 0)) a.done_list[atomicAdd(&a.counters[CNT_DONE2], 1u)] = q;
    atomicAdd(&a.stats[0], (unsigned long long)ctl.n);
    atomicAdd(&a.stats[1], (unsigned long long)ctl.m);
    atomicAdd(&a.stats[2], (unsigned long long)ctl.tiles);
    for (int i = 0; i < 6; i++) atomicAdd(&a.stats[stats_base + i], (unsigned long long)ctl.t[i]);
    if (stats_base != 4) atomicAdd(&a.stats[16], (unsigned long long)ctl.n);   // CTA tiers' share of the samples (items: p2_emit2)
  } else {
    QInfo qi; qi.item_off = 0; qi.n_items = 0; qi.contact = 0; qi.status = QS_PENDING;
    a.out.qinfo[q] = qi;
    __threadfence();   // the next tier may pick the query up (and write qinfo[q]) while this kernel is still running
    if (st == Q_OVERFLOW) a.list_out[atomicAdd(&a.counters[cnt_over], 1u)] = q;   // Q_CAPACITY: the host re-runs the batch
  }
}
__device__ __forceinline__ void zero_query(const BatchArgs& a, uint32_t q) {
  QInfo qi; qi.item_off = 0; qi.n_items = 0; qi.contact = 0; qi.status = a.out.ready;
  a.out.qinfo[q] = qi;
  if (a.done_list) a.done_list[atomicAdd(&a.counters[CNT_DONE2], 1u)] = q;
}

// ------------------------------------------------------------------------------------------------
// T0: one warp per query.  The lattice is sampled into a 96x96-cell bitmap (3x3 tiles, closure
// from the centre tile) as before, but the walk runs on one neighbour byte per cell for the 63x63 cells around the
// vertex (sections of <= 384 samples reach at most 31 cells from it) — two shared-memory loads per step, no
// unwinding iterations, no tile requests.  Per warp: control block, DFS stack, order list, bitmap, neighbour bytes;
// after the walk the neighbour bytes become the first-touch table (and the pair list of emit pass 2), the bitmap
// the per-sample masks, the stack the pair list of emit pass 1: 7.35 KB in all.
// ------------------------------------------------------------------------------------------------
template <int MAXN, int HCAP>
struct WarpNLayout {
  static constexpr int WSIDE = 96, MSIDE = 64, MLIM = 63, MOFF = 17;
  static constexpr int FSTRIDE = WSIDE / 32 + 1;                 // one shared pad word per bitmap row
  static constexpr int O_CTL = 0;
  static constexpr int O_STACK = O_CTL + 36;                     // uint16 cells
  static constexpr int O_ORDER = O_STACK + MAXN / 2;
  static constexpr int O_FBITS = O_ORDER + MAXN / 2;             // bitmap; after the walk: per-sample masks
  static constexpr int F_WORDS = ((WSIDE + 2) * FSTRIDE + 1 + 1) & ~1;
  static constexpr int O_CELLS = O_FBITS + F_WORDS;              // neighbour bytes; after the walk: first-touch table
  static constexpr int WORDS = O_CELLS + MSIDE * MSIDE / 4;
  static_assert(sizeof(Ctl) <= 36 * 4, "ctl slot");
  static_assert(MAXN <= F_WORDS, "masks alias the bitmap");
  static_assert(HCAP * 4 == MSIDE * MSIDE, "the first-touch table aliases the neighbour bytes");
  static_assert(MAXN <= 512, "packed hash rank range");
  static_assert(MOFF + MLIM <= WSIDE && WSIDE / 2 - MOFF == MLIM / 2, "byte window centred on the vertex, inside the bitmap");
};

template <int WARPS, int MAXN, int HCAP>
__global__ void __launch_bounds__(WARPS * 32, 3) area_warpn_kernel(BatchArgs a) {
  using L = WarpNLayout<MAXN, HCAP>;
  extern __shared__ __align__(16) uint32_t smem[];
  int16_t* lut = reinterpret_cast<int16_t*>(smem);               // 256 x int16, shared by the CTA's warps
  for (int m = threadIdx.x; m < 256; m += blockDim.x) lut[m] = (int16_t)(m ? nbr_delta((uint32_t)m, L::MSIDE) : 0);
  __syncthreads();
  uint32_t* base = smem + 128 + (threadIdx.x >> 5) * L::WORDS;
  Arena<uint16_t> A;
  A.tx = A.ty = L::WSIDE / 32; A.stride = 0; A.halo = 0;
  A.bits = nullptr; A.tested = nullptr;
  A.order = reinterpret_cast<uint16_t*>(base + L::O_ORDER); A.max_n = MAXN;
  A.stack = reinterpret_cast<uint16_t*>(base + L::O_STACK); A.stack_cap = MAXN; A.spill = nullptr;
  A.fbits = base + L::O_FBITS; A.fstride = L::FSTRIDE;
  A.mask = A.fbits;
  A.cells = reinterpret_cast<uint8_t*>(base + L::O_CELLS);
  A.wside = L::WSIDE; A.mside = L::MSIDE; A.moff = L::MOFF; A.mlim = L::MLIM;
  A.nlut = lut;
  A.compact_pairs = 1; A.mask_alt = nullptr; A.mask_alt_cap = 0;
  A.pairs2 = reinterpret_cast<uint16_t*>(base + L::O_CELLS); A.pairs2_cap = 2 * HCAP;
  PackedHash H;
  H.slot = base + L::O_CELLS; H.cap = HCAP; H.max_probe = 48;
  int lg = 0;
  while ((1 << lg) < HCAP) lg++;
  H.shift = 32 - lg;
  Ctl& ctl = *reinterpret_cast<Ctl*>(base + L::O_CTL);
  WarpGroup g;
  const V3 w = mk(a.wx, a.wy, a.wz);
  const uint32_t first = a.list_in ? a.counters[CNT_NBIG] : 0u;
  for (;;) {
    uint32_t q = 0;
    if (g.lane() == 0) q = first + atomicAdd(&a.counters[CNT_NEXT0], 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= a.nq) break;
    if (a.list_in) q = a.list_in[q];
    PlaneCtx c;
    const V3 p = mk(__ldg(a.pos + 3 * q), __ldg(a.pos + 3 * q + 1), __ldg(a.pos + 3 * q + 2));
    const V3 n = mk(__ldg(a.nrm + 3 * q), __ldg(a.nrm + 3 * q + 1), __ldg(a.nrm + 3 * q + 2));
    if (!plane_init(c, a.vol, p, n, w)) {
      if (g.lane() == 0) zero_query(a, q);
      continue;
    }
    A.ox = A.oy = c.c - L::WSIDE / 2;       // the vertex's cell sits in the middle of the middle tile
    const V3 rp = vround(p);
    H.cbx = (int)rp.x >> 1; H.cby = (int)rp.y >> 1; H.cbz = (int)rp.z >> 1;
    const int st = run_area_emit(g, c, a.vol, A, H, ctl, q, a.out);
    if (g.lane() == 0) finish_query(a, ctl, q, st, CNT_OVER0, 4);
    __syncwarp();
  }
}

// CTA tiers, first half: flood (tiles + walk); a non-empty section is parked in the P2 pools.  Every thread returns the
// same status: Q_OK (parked, or answered with zeros), Q_OVERFLOW (this tier's window / slab cannot hold the section),
// Q_CAPACITY (a pool is full: the host grows it and runs the batch again).  `slot`: 8 words of shared memory.
template <class CellT>
__device__ __forceinline__ int run_area_flood(const BlockGroup& g, const BatchArgs& a, const PlaneCtx& c, const Arena<CellT>& A, Ctl& ctl,
                                              uint32_t q, uint32_t* slot) {
  typedef CellPack<CellT> P;
  bool empty;
  const int fst = A.cells ? flood_component_nbr(g, c, a.vol, A, ctl, &empty) : flood_component(g, c, a.vol, A, ctl, &empty);
  if (fst != Q_OK) return Q_OVERFLOW;
  if (g.tid() == 0) prof_mark(ctl, 0);
  if (empty) {
    if (g.tid() == 0) {
      QInfo qi; qi.item_off = 0; qi.n_items = 0; qi.contact = 0; qi.status = a.out.ready;
      a.out.qinfo[q] = qi;
    }
    return Q_OK;
  }
  const P2Pools& p = a.p2;
  const uint32_t n = (uint32_t)ctl.n;
  if (g.tid() == 0) {
    const uint32_t nch = (n + P2_CHUNK - 1) / P2_CHUNK;
    uint32_t cap = 1024u;
    while (cap < (uint32_t)p.hash_factor * n && cap < (1u << 30)) cap <<= 1;
    slot[0] = atomicAdd(&p.cursors[P2C_ORDER], n);
    slot[1] = atomicAdd(&p.cursors[P2C_HASH], 2u * cap);
    slot[2] = atomicAdd(&p.cursors[P2C_CHUNK], nch);
    slot[3] = atomicAdd(&p.cursors[P2C_REC], 1u);
    slot[4] = cap;
    const bool full = (uint64_t)slot[0] + n > p.order_cap || (uint64_t)slot[1] + 2ull * cap > p.hash_cap ||
                      (uint64_t)slot[2] + nch > p.chunk_cap || slot[3] >= p.rec_cap;
    slot[5] = full ? 1u : 0u;
    if (full) p.cursors[P2C_FULL] = 1u;
  }
  g.sync();
  if (slot[5]) return Q_CAPACITY;
  const uint32_t order_off = slot[0], hash_off = slot[1], chunk0 = slot[2], rec = slot[3], cap = slot[4];
  for (uint32_t i = g.tid(); i < n; i += g.size()) {
    const CellT e = A.order[i];
    p.order[order_off + i] = CellPack<uint32_t>::pack(P::u(e), P::v(e));
  }
  for (uint32_t i = g.tid(); i < 2u * cap; i += g.size()) p.hash[hash_off + i] = XS_HEMPTY;
  const uint32_t nch = (n + P2_CHUNK - 1) / P2_CHUNK;
  for (uint32_t j = g.tid(); j < nch; j += g.size()) p.chunk_rec[chunk0 + j] = rec;
  if (g.tid() == 0) {
    P2Rec r;
    r.q = q; r.n = n; r.ox = A.ox; r.oy = A.oy; r.order_off = order_off; r.hash_off = hash_off; r.hash_cap = cap; r.chunk0 = chunk0; r.contact = 0u;
    r.pad[0] = r.pad[1] = r.pad[2] = 0u;
    r.c = c;
    p.recs[rec] = r;
    prof_mark(ctl, 1);
  }
  g.sync();
  return Q_OK;
}

// ------------------------------------------------------------------------------------------------
// T1 / T2: one CTA per query
// ------------------------------------------------------------------------------------------------
struct SlabCfg {
  char* base;            // slab of CTA b starts at base + b * stride
  size_t stride;
  int max_n;             // order / mask / spill capacity (words each)
  uint32_t hcap_max;     // htag / hkey capacity
  size_t bits_words;     // global fallback for the lattice bitmap
  size_t tested_words;   // global fallback for the tile-tested mask
  int hash_factor;
  int max_probe;
};

__host__ __device__ inline size_t slab_bytes(const SlabCfg& s) {
  size_t words = (size_t)s.max_n * 3 + (size_t)s.hcap_max * 2 + s.bits_words + s.tested_words;
  return ((words * 4 + 255) / 256) * 256;
}

// CTA-tier shapes.  MID: many small CTAs with a centred window (the few-percent of sweep queries that do not
// fit a warp).  BIG: one full-SM CTA with the whole-plane window in shared memory (large single sections).
template <int THREADS_, int RING_, int WIN_TILES_, int TESTED_, int MIN_BLOCKS_>
struct BlockTier {
  static constexpr int THREADS = THREADS_;
  static constexpr int RING = RING_;               // DFS stack ring (words) in shared memory
  static constexpr int WIN_TILES = WIN_TILES_;     // > 0: centred window of WIN_TILES^2 tiles; 0: whole-plane (bbox) window
  static constexpr int TESTED = TESTED_;           // tile-mask words in shared memory
  static constexpr int MIN_BLOCKS = MIN_BLOCKS_;
  static constexpr int CTL_WORDS = 56;            // Ctl (<= 40 words), next-query word (42), hand-off slot (43..48)
  static constexpr int O_CTL = 64;
  static constexpr int O_RING = O_CTL + CTL_WORDS;
  static constexpr int O_TESTED = O_RING + RING;
  static constexpr int O_BITS = O_TESTED + TESTED;
  static constexpr int WIN_WORDS = WIN_TILES * 32 * (WIN_TILES + 1);
  static constexpr int FIXED_WORDS = O_BITS;       // + bitmap words
};
using BigTier = BlockTier<512, 16384, 0, 1024, 1>;
constexpr int BLK_THREADS = BigTier::THREADS;
static_assert(sizeof(Ctl) <= 40 * 4, "ctl slot");

// returns false when the window does not fit any buffer this CTA owns
template <class T>
__device__ __forceinline__ bool block_arena(Arena<uint32_t>& A, WideHash& H, const SlabCfg& s, uint32_t* smem, int smem_bits_words,
                                            const PlaneCtx& c, const VolView& vol) {
  char* slab = s.base + (size_t)blockIdx.x * s.stride;
  uint32_t* p = reinterpret_cast<uint32_t*>(slab);
  A.order = p; p += s.max_n;
  A.mask = p; p += s.max_n;
  A.spill = p; p += s.max_n;
  H.htag = p; p += s.hcap_max;
  H.hkey = p; p += s.hcap_max;
  uint32_t* g_bits = p; p += s.bits_words;
  uint32_t* g_tested = p;
  A.max_n = s.max_n;
  A.cells = nullptr; A.wside = 0; A.mside = 0; A.moff = 0; A.mlim = 0; A.fbits = nullptr; A.fstride = 0; A.nlut = nullptr;
  A.compact_pairs = 0; A.mask_alt = nullptr; A.mask_alt_cap = 0;   // masks live in the HBM slab
  A.pairs2 = nullptr; A.pairs2_cap = 0;
  H.cap = s.hcap_max; H.cap_max = s.hcap_max; H.factor = s.hash_factor; H.max_probe = s.max_probe;
  H.hx = vol.hx; H.hy = vol.hy;
  H.alt_tag = nullptr; H.alt_key = nullptr; H.alt_cap = 0;
  int lg = 0;
  while ((1u << lg) < s.hcap_max) lg++;
  H.shift = 32 - lg;
  A.stack = smem + T::O_RING; A.stack_cap = T::RING;
  if (T::WIN_TILES > 0) {
    A.tx = A.ty = T::WIN_TILES;
    A.ox = A.oy = c.c - 16 - 32 * (T::WIN_TILES / 2);
    A.stride = A.tx + 1;
    A.halo = 0;
    A.bits = smem + T::O_BITS;
    A.tested = smem + T::O_TESTED;
    return true;
  }
  bbox_window(c, vol, &A.ox, &A.oy, &A.tx, &A.ty);
  A.stride = A.tx + 1;
  A.halo = 1;
  const size_t words = (size_t)A.ty * 32 * A.stride;
  const size_t twords = (size_t)((A.tx * A.ty + 31) / 32);
  const bool bits_in_smem = words <= (size_t)smem_bits_words;
  const bool tested_in_smem = twords <= (size_t)T::TESTED;
  A.bits = bits_in_smem ? (smem + T::O_BITS) : g_bits;
  A.tested = tested_in_smem ? (smem + T::O_TESTED) : g_tested;
  if (!bits_in_smem && words > s.bits_words) return false;
  if (!tested_in_smem && twords > s.tested_words) return false;
  return true;
}

template <class T>
__global__ void __launch_bounds__(T::THREADS, T::MIN_BLOCKS) area_block_kernel(BatchArgs a, SlabCfg s, int smem_bits_words, int cnt_next, int cnt_over,
                                                                                 const uint32_t* count_ptr, int stats_base) {
  extern __shared__ __align__(16) uint32_t smem[];
  BlockGroup g(reinterpret_cast<int*>(smem));
  Ctl& ctl = *reinterpret_cast<Ctl*>(smem + T::O_CTL);
  int* next = reinterpret_cast<int*>(smem + T::O_CTL + 42);
  const V3 w = mk(a.wx, a.wy, a.wz);
  const uint32_t count = count_ptr ? *count_ptr : a.nq;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *next = (int)atomicAdd(&a.counters[cnt_next], 1u);
    __syncthreads();
    const uint32_t i = (uint32_t)*next;
    if (i >= count) break;
    const uint32_t q = a.list_in ? a.list_in[i] : i;
    PlaneCtx c;
    const V3 p = mk(__ldg(a.pos + 3 * q), __ldg(a.pos + 3 * q + 1), __ldg(a.pos + 3 * q + 2));
    const V3 n = mk(__ldg(a.nrm + 3 * q), __ldg(a.nrm + 3 * q + 1), __ldg(a.nrm + 3 * q + 2));
    if (!plane_init(c, a.vol, p, n, w)) {
      if (threadIdx.x == 0) zero_query(a, q);
      continue;
    }
    Arena<uint32_t> A;
    WideHash H;
    int st = Q_OVERFLOW;
    if (block_arena<T>(A, H, s, smem, smem_bits_words, c, a.vol)) st = run_area_flood(g, a, c, A, ctl, q, smem + T::O_CTL + 43);
    if (threadIdx.x == 0) finish_query(a, ctl, q, st, cnt_over, stats_base);
  }
}

// ------------------------------------------------------------------------------------------------
// MID tier: one 256-thread CTA per query, 256x256-cell window in shared memory as one neighbour byte per cell
// (64 KB) + a foreground bitmap (10 KB): the walk is two dependent shared-memory loads per step, run by one
// warp in lock step (xs_query.cuh, flood_component_nbr).  Takes the few per cent of sweep queries that do not
// fit a warp.
// ------------------------------------------------------------------------------------------------
template <int THREADS_, int WSIDE_>
struct MidTierT {
  static constexpr int THREADS = THREADS_;
  static constexpr int WSIDE = WSIDE_;             // WSIDE x WSIDE cells; uint16 cell index u | v<<8
  static constexpr int RING = 1024;                // DFS stack ring entries (uint16) in shared memory
  static constexpr int O_CTL = 64;
  static constexpr int O_RING = O_CTL + 56;
  static constexpr int O_LUT = O_RING + RING / 2;  // 256 x int16
  static constexpr int FSTRIDE = WSIDE / 32 + 2;
  static constexpr int O_FBITS = O_LUT + 128;
  static constexpr int O_CELLS = O_FBITS + (WSIDE + 2) * FSTRIDE;
  static constexpr int WORDS = O_CELLS + WSIDE * WSIDE / 4 + 128;   // the walk may touch up to a row past either end
  static constexpr int SMEM_BYTES = WORDS * 4;
  static_assert(WSIDE % 32 == 0 && WSIDE <= 256, "uint16 cell packing");
};
// 256 x 256-cell window, 256 threads, 77.5 KB: one such CTA fits beside two warp-tier CTAs on an SM.  (Tried in round 2:
// a 160-cell window with 128 threads and TWO CTAs per SM, so that two walks hide each other's latency — the mid tier got
// slower, 0.39 -> 0.50 ms, its parallel phases having half the threads, and the warp tier beside it 0.42 -> 0.55 ms.)
using MidTier = MidTierT<256, 256>;

__host__ __device__ inline size_t mid_slab_bytes(const SlabCfg& s) {
  size_t bytes = (size_t)s.max_n * (2 + 4 + 2) + (size_t)s.hcap_max * 8;
  return ((bytes + 255) / 256) * 256;
}

template <class MidTier>
__global__ void __maxnreg__(96) area_mid_kernel(BatchArgs a, SlabCfg s, int cnt_next, int cnt_over, int stats_base, int n_direct) {
  extern __shared__ __align__(16) uint32_t smem[];
  BlockGroup g(reinterpret_cast<int*>(smem));
  Ctl& ctl = *reinterpret_cast<Ctl*>(smem + MidTier::O_CTL);
  int* next = reinterpret_cast<int*>(smem + MidTier::O_CTL + 42);
  const V3 w = mk(a.wx, a.wy, a.wz);
  // n_direct >= 0: take queries 0..n_direct-1 as they are (tiny batches); else the head of the size-ordered list
  const uint32_t n_big = n_direct >= 0 ? (uint32_t)n_direct : a.counters[CNT_NBIG];
  uint32_t late_hint = 0u;
  // per-CTA slab in HBM: order (u16), kept masks (u32), stack spill (u16), first-touch table
  char* slab = s.base + (size_t)blockIdx.x * s.stride;
  Arena<uint16_t> A;
  A.mask = reinterpret_cast<uint32_t*>(slab);
  A.order = reinterpret_cast<uint16_t*>(slab + (size_t)s.max_n * 4);
  A.spill = reinterpret_cast<uint16_t*>(slab + (size_t)s.max_n * 6);
  WideHashSeq H;
  uint32_t* const g_tag = reinterpret_cast<uint32_t*>(slab + (size_t)s.max_n * 8);
  uint32_t* const g_key = g_tag + s.hcap_max;
  A.max_n = s.max_n;
  A.stack = reinterpret_cast<uint16_t*>(smem + MidTier::O_RING); A.stack_cap = MidTier::RING;
  A.cells = reinterpret_cast<uint8_t*>(smem + MidTier::O_CELLS); A.wside = MidTier::WSIDE; A.mside = MidTier::WSIDE; A.moff = 0; A.mlim = MidTier::WSIDE;
  A.bits = nullptr; A.tested = nullptr; A.tx = A.ty = MidTier::WSIDE / 32; A.stride = 0; A.halo = 1;
  A.compact_pairs = 0;          // masks live in the HBM slab, except for sections small enough for the (dead) bitmap
  A.pairs2 = nullptr; A.pairs2_cap = 0;
  A.mask_alt = smem + MidTier::O_FBITS + 32; A.mask_alt_cap = (MidTier::WSIDE + 2) * MidTier::FSTRIDE - 32;
  A.fbits = smem + MidTier::O_FBITS; A.fstride = MidTier::FSTRIDE;
  A.nlut = reinterpret_cast<const int16_t*>(smem + MidTier::O_LUT);
  nbr_lut_fill(g, reinterpret_cast<int16_t*>(smem + MidTier::O_LUT), MidTier::WSIDE);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      // 1. the queries classified as large: the first n_big entries of the size-ordered list, largest first
      // 2. then whatever the warp tier has handed over SO FAR (it may still be running: this kernel never waits for
      //    it — a later launch of the same kernel, after the warp tier, takes the stragglers).  An entry is claimed by
      //    swapping it for XS_LATE_CLAIMED.
      uint32_t q = XS_LATE_EMPTY;
      if (a.counters[cnt_next] < n_big) {
        const uint32_t i = atomicAdd(&a.counters[cnt_next], 1u);
        if (i < n_big) q = a.list_in ? a.list_in[i] : i;
      }
      if (q == XS_LATE_EMPTY) {
        const uint32_t tail = *reinterpret_cast<volatile uint32_t*>(&a.counters[CNT_OVER0]);
        for (uint32_t j = late_hint; j < tail; j++) {
          const uint32_t e = *reinterpret_cast<volatile uint32_t*>(&a.late[j]);
          if (e >= XS_LATE_CLAIMED) continue;
          if (atomicCAS(&a.late[j], e, XS_LATE_CLAIMED) == e) { q = e; late_hint = j + 1; break; }
        }
      }
      *next = (int)q;
    }
    __syncthreads();
    const uint32_t q = (uint32_t)*next;
    if (q == XS_LATE_EMPTY) break;
    PlaneCtx c;
    const V3 p = mk(__ldg(a.pos + 3 * q), __ldg(a.pos + 3 * q + 1), __ldg(a.pos + 3 * q + 2));
    const V3 n = mk(__ldg(a.nrm + 3 * q), __ldg(a.nrm + 3 * q + 1), __ldg(a.nrm + 3 * q + 2));
    if (!plane_init(c, a.vol, p, n, w)) {
      if (threadIdx.x == 0) zero_query(a, q);
      continue;
    }
    A.ox = A.oy = c.c - (MidTier::WSIDE / 64) * 32 - 16;   // the centre cell sits in the middle of tile (2,2) / (4,4)
    H.htag = g_tag; H.hkey = g_key;
    H.cap = s.hcap_max; H.cap_max = s.hcap_max; H.factor = s.hash_factor; H.max_probe = s.max_probe;
    H.hx = a.vol.hx; H.hy = a.vol.hy;
    // the 64 KB cell window is dead once the walk is done: 8192-entry first-touch table in shared memory
    H.alt_tag = nullptr; H.alt_key = nullptr; H.alt_cap = 0;
    int lg = 0;
    while ((1u << lg) < s.hcap_max) lg++;
    H.shift = 32 - lg;
    const int st = run_area_flood(g, a, c, A, ctl, q, smem + MidTier::O_CTL + 43);
    if (threadIdx.x == 0) finish_query(a, ctl, q, st, cnt_over, stats_base);
  }
}

// ------------------------------------------------------------------------------------------------
// WIDE tier: the mid tier's neighbour-byte walk with the largest window one SM holds — 416 x 416 cells (169 KB of
// neighbour bytes + 25 KB of foreground bitmap), 32-bit cell indices, one 512-thread CTA per SM.  Takes the sections
// that leave the mid tier's 256-cell window or exceed its 16 Ki samples (BASELINE config 0: a 500^3 sphere cut through
// its centre, 1.3e5 samples) before they fall back to the big tier's bitmap walk (3-4x slower per step).
// ------------------------------------------------------------------------------------------------
struct WideTier {
  static constexpr int THREADS = 512;
  static constexpr int WSIDE = 416;                // 13 x 13 tiles; cell index u + 416 v
  static constexpr int RING = 2048;                // DFS stack ring entries (uint32) in shared memory
  static constexpr int O_CTL = 64;
  static constexpr int O_RING = O_CTL + 56;
  static constexpr int O_LUT = O_RING + RING;      // 256 x int16
  static constexpr int FSTRIDE = WSIDE / 32 + 2;
  static constexpr int O_FBITS = O_LUT + 128;
  static constexpr int O_CELLS = O_FBITS + (WSIDE + 2) * FSTRIDE;
  static constexpr int WORDS = O_CELLS + WSIDE * WSIDE / 4 + 128;
  static constexpr int ALT_CAP = 16384;            // first-touch table aliased onto the (dead) neighbour bytes
  static_assert(2 * ALT_CAP * 4 <= WSIDE * WSIDE, "alt table fits the cell window");
};
constexpr int WIDE_SMEM_BYTES = WideTier::WORDS * 4;

__global__ void __launch_bounds__(WideTier::THREADS, 1) area_wide_kernel(BatchArgs a, SlabCfg s, int cnt_next, int cnt_over, const uint32_t* count_ptr, int stats_base) {
  extern __shared__ __align__(16) uint32_t smem[];
  BlockGroup g(reinterpret_cast<int*>(smem));
  Ctl& ctl = *reinterpret_cast<Ctl*>(smem + WideTier::O_CTL);
  int* next = reinterpret_cast<int*>(smem + WideTier::O_CTL + 42);
  const V3 w = mk(a.wx, a.wy, a.wz);
  const uint32_t count = count_ptr ? *count_ptr : a.nq;
  char* slab = s.base + (size_t)blockIdx.x * s.stride;
  Arena<uint32_t> A;
  uint32_t* p = reinterpret_cast<uint32_t*>(slab);
  A.order = p; p += s.max_n;
  A.mask = p; p += s.max_n;
  A.spill = p; p += s.max_n;
  uint32_t* const g_tag = p; p += s.hcap_max;
  uint32_t* const g_key = p;
  A.max_n = s.max_n;
  A.stack = smem + WideTier::O_RING; A.stack_cap = WideTier::RING;
  A.cells = reinterpret_cast<uint8_t*>(smem + WideTier::O_CELLS); A.wside = WideTier::WSIDE; A.mside = WideTier::WSIDE; A.moff = 0; A.mlim = WideTier::WSIDE;
  A.bits = nullptr; A.tested = nullptr; A.tx = A.ty = WideTier::WSIDE / 32; A.stride = 0; A.halo = 1;
  A.compact_pairs = 0;          // masks live in the HBM slab, except for sections small enough for the (dead) bitmap
  A.pairs2 = nullptr; A.pairs2_cap = 0;
  A.mask_alt = smem + WideTier::O_FBITS + 32; A.mask_alt_cap = (WideTier::WSIDE + 2) * WideTier::FSTRIDE - 32;
  A.fbits = smem + WideTier::O_FBITS; A.fstride = WideTier::FSTRIDE;
  A.nlut = reinterpret_cast<const int16_t*>(smem + WideTier::O_LUT);
  nbr_lut_fill(g, reinterpret_cast<int16_t*>(smem + WideTier::O_LUT), WideTier::WSIDE);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) *next = (int)atomicAdd(&a.counters[cnt_next], 1u);
    __syncthreads();
    const uint32_t i = (uint32_t)*next;
    if (i >= count) break;
    const uint32_t q = a.list_in ? a.list_in[i] : i;
    PlaneCtx c;
    const V3 pp = mk(__ldg(a.pos + 3 * q), __ldg(a.pos + 3 * q + 1), __ldg(a.pos + 3 * q + 2));
    const V3 n = mk(__ldg(a.nrm + 3 * q), __ldg(a.nrm + 3 * q + 1), __ldg(a.nrm + 3 * q + 2));
    if (!plane_init(c, a.vol, pp, n, w)) {
      if (threadIdx.x == 0) zero_query(a, q);
      continue;
    }
    A.ox = A.oy = c.c - (WideTier::WSIDE / 64) * 32 - 16;   // the centre cell sits in the middle of the middle tile
    WideHash H;
    H.htag = g_tag; H.hkey = g_key;
    H.cap = s.hcap_max; H.cap_max = s.hcap_max; H.factor = s.hash_factor; H.max_probe = s.max_probe;
    H.hx = a.vol.hx; H.hy = a.vol.hy;
    H.alt_tag = smem + WideTier::O_CELLS; H.alt_key = H.alt_tag + WideTier::ALT_CAP; H.alt_cap = WideTier::ALT_CAP;
    int lg = 0;
    while ((1u << lg) < s.hcap_max) lg++;
    H.shift = 32 - lg;
    const int st = run_area_flood(g, a, c, A, ctl, q, smem + WideTier::O_CTL + 43);
    if (threadIdx.x == 0) finish_query(a, ctl, q, st, cnt_over, stats_base);
  }
}

// ------------------------------------------------------------------------------------------------
// P2 kernels: one 256-thread CTA per chunk of 256 consecutive ranks of a parked query (grid-stride over the chunks
// parked since `chunk_start`).  The query's plane is rebuilt from its vertex and normal (plane_init is deterministic),
// its window origin comes from the P2Rec, its first-touch table is its own region of the hash pool (L2-resident).
// ------------------------------------------------------------------------------------------------
struct P2Chunk {
  Arena<uint32_t> A;        // only order / mask / ox / oy are meaningful
  WideHash H;
  const P2Rec* rec;
  uint32_t rec_index, n, q;
  int r;                    // this thread's rank (may be >= n)
};
__device__ __forceinline__ void p2_chunk_setup(const BatchArgs& a, uint32_t chunk, P2Chunk& k) {
  const P2Pools& p = a.p2;
  k.rec_index = __ldg(p.chunk_rec + chunk);
  const P2Rec* rec = p.recs + k.rec_index;
  k.rec = rec;
  k.n = rec->n; k.q = rec->q;
  const uint32_t cap = rec->hash_cap;
  k.A.order = p.order + rec->order_off; k.A.ox = rec->ox; k.A.oy = rec->oy;
  k.A.mask = p.mask + rec->order_off;
  k.H.htag = p.hash + rec->hash_off; k.H.hkey = k.H.htag + cap;
  k.H.cap = cap; k.H.cap_max = cap; k.H.factor = 0; k.H.max_probe = (int)min(cap, 1u << 20);
  k.H.hx = a.vol.hx; k.H.hy = a.vol.hy;
  k.H.alt_tag = nullptr; k.H.alt_key = nullptr; k.H.alt_cap = 0;
  k.H.shift = 32 - (31 - __clz(cap));
  k.r = (int)((chunk - rec->chunk0) * P2_CHUNK + threadIdx.x);
}

// contact bits + first-touch claims.  All chunks of all parked sections claim at once, and neighbouring samples probe
// the same blocks (~16 claims per distinct block): sent straight to the tables in L2 that is ~9 M atomics on ~0.5 M
// addresses, and the L2 atomic units become the bottleneck (146 us for the bench sweep).  So a chunk first settles its
// claims among its own 256 samples in a shared-memory table (tag -> smallest rank), and only the winners — one claim
// per distinct block and chunk — go to the section's table; a sample is a tentative owner of a block when it won both.
constexpr int P2_LCAP = 2048;               // local table slots (a chunk touches a few hundred distinct blocks)
constexpr int P2_LPROBE = 32;
__global__ void __launch_bounds__(P2_CHUNK) p2_claims_kernel(BatchArgs a) {
  __shared__ uint32_t s_tag[P2_LCAP], s_key[P2_LCAP];
  __shared__ uint16_t s_list[P2_LCAP];
  __shared__ uint32_t s_tent[P2_CHUNK];
  __shared__ int s_wsum[P2_CHUNK / 32];
  const uint32_t nchunks = min(a.p2.cursors[P2C_CHUNK], a.p2.chunk_cap), chunk_start = a.p2.cursors[P2C_DONE];
  if (a.p2.cursors[P2C_FULL]) return;
  for (uint32_t chunk = chunk_start + blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    P2Chunk k;
    p2_chunk_setup(a, chunk, k);
    const PlaneCtx& c = k.rec->c;
    __syncthreads();
    for (int i = threadIdx.x; i < P2_LCAP; i += P2_CHUNK) { s_tag[i] = XS_HEMPTY; s_key[i] = XS_HEMPTY; }
    s_tent[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t ct = 0u, tentative = 0u;
    bool ovf = false;
    const bool live = k.r < (int)k.n;
    if (live) {
      V3 cur;
      const Sample s = sample_at(c, k.A, k.r, &cur);
      ct = contact_bits_rounded(c, cur);
      const uint32_t tag = k.H.tag(s.rx >> 1, s.ry >> 1, s.rz >> 1);
      uint32_t m = probe_candidates(c, a.vol, s);
      while (m) {
        const int kk = __ffs(m) - 1;
        m &= m - 1u;
        int ox, oy, oz;
        decode_k(kk, ox, oy, oz);
        const uint32_t t = tag + k.H.tag_delta(ox, oy, oz);
        uint32_t h = (t * 2654435761u) >> (32 - 11);
        bool placed = false;
        // the local key is (thread << 5 | offset): ordered by rank (a sample probes a block once), and the slot's winner
        // is known with its offset — it gets its tentative bit from whoever sends the slot's claim on (no second walk
        // over the 27 offsets to ask the table who won)
        for (int probes = 0; probes < P2_LPROBE; probes++) {
          const uint32_t old = atomicCAS(&s_tag[h], XS_HEMPTY, t);
          if (old == XS_HEMPTY || old == t) { atomicMin(&s_key[h], (threadIdx.x << 5) | (uint32_t)kk); placed = true; break; }
          h = (h + 1u) & (P2_LCAP - 1u);
        }
        if (!placed) {                                           // local table crowded: this claim goes to the section's table itself
          const int won = k.H.claim_tag(t, (uint32_t)k.r);
          if (won == 0) ovf = true;
          if (won == 2) tentative |= 1u << kk;
        }
      }
    }
    __syncthreads();
    // the occupied slots (a few hundred of 2048), listed densely so that every thread sends one or two claims instead of
    // walking eight mostly empty slots with their dependent L2 round trips
    {
      uint32_t mine = 0u;
#pragma unroll
      for (int j = 0; j < P2_LCAP / P2_CHUNK; j++) mine |= (s_tag[threadIdx.x * (P2_LCAP / P2_CHUNK) + j] != XS_HEMPTY ? 1u : 0u) << j;
      const int cnt = __popc(mine);
      int inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += t;
      }
      if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = inc;
      __syncthreads();
      int base = 0, total = 0;
#pragma unroll
      for (int w = 0; w < P2_CHUNK / 32; w++) { const int t = s_wsum[w]; if (w < (int)(threadIdx.x >> 5)) base += t; total += t; }
      int o = base + inc - cnt;
#pragma unroll
      for (int j = 0; j < P2_LCAP / P2_CHUNK; j++)
        if ((mine >> j) & 1u) s_list[o++] = (uint16_t)(threadIdx.x * (P2_LCAP / P2_CHUNK) + j);
      __syncthreads();
      const uint32_t rank0 = (uint32_t)(k.r - (int)threadIdx.x);  // rank of the chunk's first sample
      for (int e = threadIdx.x; e < total; e += P2_CHUNK) {
        const int i = s_list[e];
        const uint32_t key = s_key[i];
        const int won = k.H.claim_tag(s_tag[i], rank0 + (key >> 5));
        if (won == 0) ovf = true;
        if (won == 2) atomicOr(&s_tent[key >> 5], 1u << (key & 31u));   // else: a smaller rank of another chunk holds the block
      }
    }
    __syncthreads();
    if (live) k.A.mask[k.r] = tentative | s_tent[threadIdx.x];
    ct = __reduce_or_sync(0xffffffffu, ct);
    if ((threadIdx.x & 31) == 0 && ct) atomicOr(&a.p2.recs[k.rec_index].contact, ct);
    if (ovf) a.p2.cursors[P2C_HASHFULL] = 1u;
  }
}

// ownership -> adjacency -> plane distance: kept masks, items / polygons per chunk (emit_items pass 1).  A sample has 0-5
// tentative blocks, most have one or two: the (sample, offset) pairs of the chunk are listed densely in shared memory and
// evaluated one per thread, so that the three dependent L2 round trips of a pair (table tag, table rank, occupancy byte)
// are paid once or twice per chunk instead of up to five times by a few lanes.
constexpr int P2_PAIRS = 2048;
__device__ __forceinline__ bool p2_keep_pair(const BatchArgs& a, const P2Chunk& k, const PlaneCtx& c, int r, int kk, int* polys) {
  const Sample s = sample_at(c, k.A, r, (V3*)nullptr);
  const int bx = s.rx >> 1, by = s.ry >> 1, bz = s.rz >> 1;
  int ox, oy, oz;
  decode_k(kk, ox, oy, oz);
  if (k.H.owner(bx + ox, by + oy, bz + oz) != (uint32_t)r) return false;          // someone with a smaller rank touched it first
  const uint32_t centre = cube_at(a.vol, bx, by, bz);
  const uint32_t cand = cube_at(a.vol, bx + ox, by + oy, bz + oz);
  if (!blocks_adjacent(centre, cand, ox, oy, oz)) return false;                   // visited, but contributes nothing
  const V3 ctr = vadd(mk((float)(2 * (bx + ox)), (float)(2 * (by + oy)), (float)(2 * (bz + oz))), mk(0.5f, 0.5f, 0.5f));
  if (fabsf(vdot(vsub(ctr, c.g.pos), c.g.n)) > XS_FAR2) return false;             // block value is exactly 0
  *polys = item_polys(cand);
  return true;
}
__global__ void __launch_bounds__(P2_CHUNK) p2_emit1_kernel(BatchArgs a) {
  __shared__ int sh_sum[2][P2_CHUNK / 32];
  __shared__ int s_wsum[P2_CHUNK / 32];
  __shared__ uint32_t s_keep[P2_CHUNK];
  __shared__ uint16_t s_pairs[P2_PAIRS];
  const uint32_t nchunks = min(a.p2.cursors[P2C_CHUNK], a.p2.chunk_cap), chunk_start = a.p2.cursors[P2C_DONE];
  if (a.p2.cursors[P2C_FULL] || a.p2.cursors[P2C_HASHFULL]) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t chunk = chunk_start + blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    P2Chunk k;
    p2_chunk_setup(a, chunk, k);
    const PlaneCtx& c = k.rec->c;
    int my_items = 0, my_polys = 0;
    const bool live = k.r < (int)k.n;
    const uint32_t tent = live ? k.A.mask[k.r] : 0u;
    const int cnt = __popc(tent);
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    __syncthreads();                       // (the shared arrays of the previous chunk have been read)
    if (lane == 31) s_wsum[warp] = inc;
    s_keep[threadIdx.x] = 0u;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < P2_CHUNK / 32; w++) { const int t = s_wsum[w]; if (w < warp) base += t; total += t; }
    const int r0 = k.r - (int)threadIdx.x;                       // rank of thread 0
    if (total <= P2_PAIRS) {
      int o = base + inc - cnt;
      uint32_t m = tent;
      while (m) {
        const int kk = __ffs(m) - 1;
        m &= m - 1u;
        s_pairs[o++] = (uint16_t)((threadIdx.x << 5) | kk);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < total; e += P2_CHUNK) {
        const int pr = s_pairs[e], t = pr >> 5, kk = pr & 31;
        int polys;
        if (p2_keep_pair(a, k, c, r0 + t, kk, &polys)) { atomicOr(&s_keep[t], 1u << kk); my_items++; my_polys += polys; }
      }
      __syncthreads();
      if (live) k.A.mask[k.r] = s_keep[threadIdx.x];
    } else if (live) {
      uint32_t m = tent, keep = 0u;
      while (m) {
        const int kk = __ffs(m) - 1;
        m &= m - 1u;
        int polys;
        if (p2_keep_pair(a, k, c, k.r, kk, &polys)) { keep |= 1u << kk; my_items++; my_polys += polys; }
      }
      k.A.mask[k.r] = keep;
    }
    __syncwarp();
    my_items = __reduce_add_sync(0xffffffffu, my_items);
    my_polys = __reduce_add_sync(0xffffffffu, my_polys);
    if (lane == 0) { sh_sum[0][warp] = my_items; sh_sum[1][warp] = my_polys; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int ti = 0, tp = 0;
      for (int w = 0; w < P2_CHUNK / 32; w++) { ti += sh_sum[0][w]; tp += sh_sum[1][w]; }
      a.p2.chunk_items[chunk] = (uint32_t)ti; a.p2.chunk_polys[chunk] = (uint32_t)tp;
    }
  }
}

// exclusive prefix sums of the chunk totals, continuing from the records already emitted (one CTA)
__global__ void __launch_bounds__(1024) p2_scan_kernel(BatchArgs a) {
  __shared__ unsigned long long wsum[2][32];
  const P2Pools& p = a.p2;
  const uint32_t nchunks = min(p.cursors[P2C_CHUNK], p.chunk_cap), chunk_start = p.cursors[P2C_DONE];
  if (p.cursors[P2C_FULL] || p.cursors[P2C_HASHFULL] || nchunks <= chunk_start) return;
  const uint32_t cnt = nchunks - chunk_start;
  const uint32_t per = (cnt + 1023u) / 1024u;
  const uint32_t lo = min(cnt, threadIdx.x * per), hi = min(cnt, lo + per);
  unsigned long long si = 0, sp = 0;
  for (uint32_t i = lo; i < hi; i++) { si += p.chunk_items[chunk_start + i]; sp += p.chunk_polys[chunk_start + i]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long ii = si, ip = sp;                       // inclusive over the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long ti = __shfl_up_sync(0xffffffffu, ii, d), tp = __shfl_up_sync(0xffffffffu, ip, d);
    if (lane >= d) { ii += ti; ip += tp; }
  }
  if (lane == 31) { wsum[0][warp] = ii; wsum[1][warp] = ip; }
  __syncthreads();
  if (warp == 0) {
    unsigned long long wi = wsum[0][lane], wp = wsum[1][lane];
    const unsigned long long mi = wi, mp = wp;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long ti = __shfl_up_sync(0xffffffffu, wi, d), tp = __shfl_up_sync(0xffffffffu, wp, d);
      if (lane >= d) { wi += ti; wp += tp; }
    }
    wsum[0][lane] = wi - mi; wsum[1][lane] = wp - mp;        // exclusive over the warps
    if (lane == 31) {
      const unsigned long long ri = (unsigned long long)a.out.counters[1] + wi, rp = (unsigned long long)a.out.counters[0] + wp;
      p.item_off[nchunks] = (uint32_t)min(ri, 0xffffffffull); p.poly_off[nchunks] = (uint32_t)min(rp, 0xffffffffull);
      if (ri > a.out.item_cap || rp > a.out.poly_cap) a.out.counters[2] = 1u;    // the host grows the buffers and runs the batch again
    }
  }
  __syncthreads();
  const unsigned long long bi = a.out.counters[1], bp = a.out.counters[0];
  si = bi + wsum[0][warp] + ii - si; sp = bp + wsum[1][warp] + ip - sp;
  for (uint32_t i = lo; i < hi; i++) {
    p.item_off[chunk_start + i] = (uint32_t)si; p.poly_off[chunk_start + i] = (uint32_t)sp;
    si += p.chunk_items[chunk_start + i]; sp += p.chunk_polys[chunk_start + i];
  }
  __syncthreads();
  if (threadIdx.x == 0) { a.out.counters[1] = p.item_off[nchunks]; a.out.counters[0] = p.poly_off[nchunks]; }
}

// item and polygon records in (rank, k) order (emit_items pass 2); the first chunk of a query publishes its QInfo
__global__ void __launch_bounds__(P2_CHUNK) p2_emit2_kernel(BatchArgs a) {
  __shared__ int scratch[40];
  const P2Pools& p = a.p2;
  const uint32_t nchunks = min(p.cursors[P2C_CHUNK], p.chunk_cap), chunk_start = p.cursors[P2C_DONE];
  if (p.cursors[P2C_FULL] || p.cursors[P2C_HASHFULL] || a.out.counters[2]) return;
  BlockGroup g(scratch);
  const EmitOut& out = a.out;
  for (uint32_t chunk = chunk_start + blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    P2Chunk k;
    p2_chunk_setup(a, chunk, k);
    const PlaneCtx& c = k.rec->c;
    const uint32_t q = k.q;
    const uint32_t keep = (k.r < (int)k.n) ? k.A.mask[k.r] : 0u;
    int bx = 0, by = 0, bz = 0, ni = 0, np = 0;
    if (keep) {
      const Sample s = sample_at(c, k.A, k.r, (V3*)nullptr);
      bx = s.rx >> 1; by = s.ry >> 1; bz = s.rz >> 1;
      uint32_t km = keep;
      while (km) {
        const int kk = __ffs(km) - 1;
        km &= km - 1u;
        ni++;
        int ox, oy, oz;
        decode_k(kk, ox, oy, oz);
        np += item_polys(cube_at(a.vol, bx + ox, by + oy, bz + oz));
      }
    }
    int ti, tp;
    uint32_t io = p.item_off[chunk] + (uint32_t)g.exscan(ni, ti);
    uint32_t po = p.poly_off[chunk] + (uint32_t)g.exscan(np, tp);
    uint32_t km = keep;
    bool first = true;
    while (km) {
      const int kk = __ffs(km) - 1;
      km &= km - 1u;
      int ox, oy, oz;
      decode_k(kk, ox, oy, oz);
      const int cx = bx + ox, cy = by + oy, cz = bz + oz;
      const uint32_t cand = cube_at(a.vol, cx, cy, cz);
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
    if (chunk == k.rec->chunk0 && threadIdx.x == 0) {
      const uint32_t nch = (k.n + P2_CHUNK - 1) / P2_CHUNK;
      QInfo qi;
      qi.item_off = p.item_off[chunk]; qi.n_items = p.item_off[chunk + nch] - p.item_off[chunk];
      qi.contact = *reinterpret_cast<volatile uint32_t*>(&p.recs[k.rec_index].contact); qi.status = out.ready;
      out.geom[q] = c.g;
      out.qinfo[q] = qi;
      if (a.done_list) a.done_list[atomicAdd(&a.counters[CNT_DONE2], 1u)] = q;
      atomicAdd(&a.stats[1], (unsigned long long)qi.n_items);
      atomicAdd(&a.stats[17], (unsigned long long)qi.n_items);
    }
  }
  // the last CTA through declares the chunks done: the next chain of P2 kernels starts behind them (nobody parks a
  // section while a chain runs: the flood kernels of this stream come before or after it)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&p.cursors[P2C_EXITED], 1u) == gridDim.x - 1u) { p.cursors[P2C_EXITED] = 0u; p.cursors[P2C_DONE] = nchunks; }
  }
}

// ------------------------------------------------------------------------------------------------
// polygon kernel.  Two thirds of the records are cubes the plane does not cut (value exactly 0) and the rest differ in
// vertex count, so one-thread-per-record leaves most lanes idle.  Each warp therefore FILTERS 32 records at a time
// (plane_cube_miss: the reach of the cube along the normal, one dot product), writes the zeros, queues the survivors in
// shared memory and evaluates them 32 at a time, all lanes busy; unit voxels and 2x2x2 blocks run through the same
// instruction stream (run-time cube side).  The evaluation is ~3000 instructions long: it exists ONCE in the kernel (the
// loop drains the queue through the same call site), a second inlined copy thrashed the instruction cache.
// ------------------------------------------------------------------------------------------------
constexpr int POLY_THREADS = 256;
__global__ void __launch_bounds__(POLY_THREADS, 4) poly_eval_kernel(const PolyRec* __restrict__ polys, const PlaneGeom* __restrict__ geom,
                                                                 float* __restrict__ vals, const uint32_t* __restrict__ count_ptr,
                                                                 uint32_t start, uint32_t cap) {
  __shared__ uint32_t queue_all[POLY_THREADS / 32][64];
  uint32_t* queue = queue_all[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  uint32_t n = *count_ptr;
  if (n > cap) n = cap;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int qn = 0;
  uint32_t base = start + wid * 32u;
#pragma unroll 1
  for (;;) {
    const bool more = base < n;                 // warp-uniform
    if (more) {
      const uint32_t i = base + (uint32_t)lane;
      bool near = false;
      if (i < n) {
        const PolyRec r = polys[i];
        const PlaneGeom* gp = geom + r.q;
        const V3 pos = mk(__ldg(&gp->pos.x), __ldg(&gp->pos.y), __ldg(&gp->pos.z));
        const V3 nn = mk(__ldg(&gp->n.x), __ldg(&gp->n.y), __ldg(&gp->n.z));
        near = !plane_cube_miss((int)r.kind, (float)r.x, (float)r.y, (float)r.z, pos, nn);
        if (!near) vals[i] = 0.0f;
      }
      const uint32_t ball = __ballot_sync(0xffffffffu, near);
      if (near) queue[qn + __popc(ball & ((1u << lane) - 1u))] = i;
      qn += __popc(ball);
      base += nwarps * 32u;
      __syncwarp();
      if (qn < 32) continue;
    } else if (qn == 0) {
      break;
    }
    if (lane < qn) {
      const uint32_t j = queue[lane];
      const PolyRec r = polys[j];
      vals[j] = poly_eval(r, geom[r.q]);
    }
    const uint32_t rest = (lane + 32 < qn) ? queue[lane + 32] : 0u;
    __syncwarp();
    queue[lane] = rest;
    qn = qn > 32 ? qn - 32 : 0;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// combine kernel: one CTA per query; item values in parallel, then the reference's ordered sums
// ------------------------------------------------------------------------------------------------
#ifndef XS_COMBINE_TILED_FROM
#define XS_COMBINE_TILED_FROM 256u
#endif
__global__ void __launch_bounds__(128) combine_kernel(const ItemRec* __restrict__ items, float* __restrict__ vals, const QInfo* __restrict__ qinfo,
                                                      const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count, uint32_t nq,
                                                      float* __restrict__ area, uint8_t* __restrict__ contact, uint32_t want_status) {
  const uint32_t count = list ? *list_count : nq;
  for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
    const uint32_t q = list ? list[i] : i;
    const QInfo qi = qinfo[q];
    if (qi.status != want_status) continue;     // pending, or owned by the other pipeline
    const ItemRec* it = items + qi.item_off;
    const uint32_t n = qi.n_items;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
      const ItemRec e = it[k];
      vals[e.poly] = item_value(e, vals);       // an item's polygons are its own: safe to overwrite its first slot
    }
    __syncthreads();
    if (n > XS_COMBINE_TILED_FROM) {
      // a large section (up to ~1e5 items): the ordered sum is one dependent chain of float additions whatever we do, so
      // make the chain as short as it can be.  Per tile of 1024 items: all threads stage (value, first) pairs in shared
      // memory; every thread sums the items of the samples that START in its eight slots (a sample's own sum is a short
      // chain of 1-27 additions, independent of every other sample's) and the sums are compacted in order; thread 0 then
      // runs the one chain that is inherently serial — total += sum of the next sample — one addition per SAMPLE with
      // nothing else on the dependency path (it used to walk every item through two shared loads and a select).
      __shared__ float tile_v[1024];
      __shared__ uint8_t tile_f[1024];
      __shared__ __align__(16) float dense[1024];
      __shared__ int wcount[4];
      __shared__ float carry[2];
      if (threadIdx.x == 0) { carry[0] = 0.0f; carry[1] = 0.0f; }
      const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
      for (uint32_t b = 0; b < n; b += 1024u) {
        __syncthreads();
        const int len = (int)min(1024u, n - b);
        for (int j = threadIdx.x; j < len; j += blockDim.x) { const ItemRec e = it[b + j]; tile_v[j] = vals[e.poly]; tile_f[j] = (uint8_t)((e.meta >> 5) & 1u); }
        __syncthreads();
        const int j0 = (int)threadIdx.x * 8;
        int mine = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) mine += (j0 + t < len && tile_f[j0 + t]) ? 1 : 0;
        int inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) wcount[wrp] = inc;
        __syncthreads();
        int off = inc - mine, nseg = 0;
        for (int i = 0; i < 4; i++) { if (i < wrp) off += wcount[i]; nseg += wcount[i]; }
        for (int t = 0; t < 8; t++) {
          const int j = j0 + t;
          if (j < len && tile_f[j]) {
            float sub = fadd(0.0f, tile_v[j]);
            for (int k = j + 1; k < len && !tile_f[k]; k++) sub = fadd(sub, tile_v[k]);
            dense[off++] = sub;
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          float total = carry[0], sub = carry[1];
          for (int j = 0; j < len && !tile_f[j]; j++) sub = fadd(sub, tile_v[j]);      // the tail of a sample that began in the previous tile
          int i = 0;
          for (; i + 4 <= nseg; i += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(&dense[i]);
            total = fadd(total, sub); total = fadd(total, d4.x); total = fadd(total, d4.y); total = fadd(total, d4.z);
            sub = d4.w;
          }
          for (; i < nseg; i++) { total = fadd(total, sub); sub = dense[i]; }
          carry[0] = total; carry[1] = sub;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) { area[q] = fadd(carry[0], carry[1]); contact[q] = (uint8_t)qi.contact; }
      __syncthreads();
      continue;
    }
    if (threadIdx.x < 32) {
      OrderedSum s;
      s.init();
      for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t k = b + threadIdx.x;
        float v = 0.0f;
        int first = 0;
        if (k < n) { const ItemRec e = it[k]; v = vals[e.poly]; first = (e.meta >> 5) & 1; }
        const int len = (n - b) < 32u ? (int)(n - b) : 32;
        for (int j = 0; j < len; j++) {
          const float vj = __shfl_sync(0xffffffffu, v, j);
          const int fj = __shfl_sync(0xffffffffu, first, j);
          s.add(vj, fj != 0);
        }
      }
      if (threadIdx.x == 0) { area[q] = s.finish(); contact[q] = (uint8_t)qi.contact; }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// pre-pass: a cheap size estimate per query -> queries ordered largest first.
// classify_kernel: one warp per query probes the lattice on a 16x16 grid of stride 4 around the centre cell (256
//   probes, 8 per lane); key = number of foreground probes (each stands for 16 cells), clamped to 255, 0 for queries
//   the reference answers with zeros.  Keys are histogrammed.
// order_kernel: counting sort by descending key (order inside a key is arbitrary).  Queries with key >= T go straight
//   to the mid tier, which therefore starts at t = 0 beside the warp tier; the warp tier takes the rest largest first
//   (longest-processing-time-first keeps the tail of its persistent warps short).  The estimate only steers
//   scheduling: a query the warp tier cannot fit is still handed over, and every tier computes the same bits.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) classify_kernel(VolView vol, const float* __restrict__ pos, const float* __restrict__ nrm,
                                                       float wx, float wy, float wz, uint32_t nq, uint8_t* __restrict__ key,
                                                       uint32_t* __restrict__ hist, uint32_t* __restrict__ late, QInfo* __restrict__ qinfo,
                                                       uint32_t* __restrict__ bad_normal) {
  const int lane = threadIdx.x & 31;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += (gridDim.x * blockDim.x) >> 5) {
    PlaneCtx c;
    const V3 p = mk(__ldg(pos + 3 * q), __ldg(pos + 3 * q + 1), __ldg(pos + 3 * q + 2));
    const V3 n = mk(__ldg(nrm + 3 * q), __ldg(nrm + 3 * q + 1), __ldg(nrm + 3 * q + 2));
    int cnt = 0;
    const bool ok = plane_init(c, vol, p, n, mk(wx, wy, wz));
    if (ok) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int pi = lane * 8 + k;
        cnt += cell_fg(c, vol, c.c + ((pi & 15) - 8) * 4 + 2, c.c + ((pi >> 4) - 8) * 4 + 2) ? 1 : 0;
      }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
      const uint32_t k = ok ? (uint32_t)(cnt < 1 ? 1 : (cnt > 255 ? 255 : cnt)) : 0u;
      key[q] = (uint8_t)k;
      atomicAdd(&hist[k], 1u);
      if (n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) *bad_normal = 1u;   // reported as an argument error after the batch
      late[q] = XS_LATE_EMPTY;
      qinfo[q].status = QS_PENDING;   // (a stale READY from the previous batch would let the wrong pipeline's combine kernel write area[q])
    }
  }
}

__global__ void __launch_bounds__(256) order_kernel(const uint8_t* __restrict__ key, const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor,
                                                    uint32_t* __restrict__ sorted, uint32_t nq, uint32_t* __restrict__ n_big, int big_key) {
  __shared__ uint32_t start[256];
  {
    uint32_t s = 0;
    for (int j = threadIdx.x + 1; j < 256; j++) s += hist[j];   // queries with a larger key come first
    start[threadIdx.x] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_big = start[big_key - 1];
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const uint32_t k = key[q];
    sorted[start[k] + atomicAdd(&cursor[k], 1u)] = q;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static unsigned long long g_alloc_gen = 0;   // bumped whenever a device buffer moves: captured launch sequences are stale then
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return XS3D_OK;
    g_alloc_gen++;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return fail(XS3D_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
    cap = want;
    return XS3D_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// warp tier configuration (shared memory per warp = WarpNLayout::WORDS * 4 bytes)
constexpr int T0_WARPS = 10, T0_MAXN = 384, T0_HCAP = 1024;   // 3 CTAs per SM -> 30 warps per SM (2 beside a mid-tier CTA)
constexpr int T0_CTAS_PER_SM = 3;
constexpr int MID_BIG_KEY = 26;   // >= 26 of 256 probes foreground (~420 cells): the section will not fit a warp (T0_MAXN); tuned on the
                                  //   bench sweep — a query misjudged as small is handed over by the warp tier, only later
using T0NLayout = WarpNLayout<T0_MAXN, T0_HCAP>;
#define T0_KERNEL area_warpn_kernel<T0_WARPS, T0_MAXN, T0_HCAP>
constexpr int T0_SMEM_BYTES = (128 + T0_WARPS * T0NLayout::WORDS) * 4;

struct xs3d_volume {
  int device = 0;
  int sx = 0, sy = 0, sz = 0, hx = 0, hy = 0, hz = 0;
  cudaStream_t stream = nullptr;
  DevBuf cubes, raw, dbits;        // dbits: bit stream built on the device for images of awkward sizes
  bool loaded = false;
  // labels (slice)
  DevBuf labels;
  int itemsize = 0, labels_c_order = 0;
  bool labels_loaded = false;
  // batch workspace
  DevBuf q_pos, q_nrm, q_area, q_contact, lists, counters, slab1, slab2, sec_idx, sec_val, misc;
  DevBuf polys, items, vals, geom, qinfo, slab_mid, keys;
  DevBuf polys2, items2, vals2;                        // record buffers of the second (mid / big tier) pipeline
  // captured launch sequences of area_batch_device (CUDA graphs), keyed by everything baked into their nodes
  struct BatchGraph {
    unsigned long long gen = 0; uint64_t n = 0; const void* ptr[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float w[3] = {0, 0, 0}; int single = 0, mid_key = 0, factor = 0; uint64_t caps[5] = {0, 0, 0, 0, 0};
    cudaGraphExec_t exec = nullptr; int launches = 0; unsigned long long used = 0;
  };
  std::vector<BatchGraph> graphs;
  unsigned long long graph_clock = 0;
  DevBuf p2_order, p2_mask, p2_hash, p2_chunks, p2_recs;   // pools of the CTA tiers' second half (P2Pools)
  uint64_t p2_order_cap = 0, p2_rec_cap = 0, p2_hash_min_words = 0;
  int p2_factor = 8;                                   // first-touch table slots per sample (x4 when a table fills up)
  DevBuf slice_bufs[6];                                // batched slices: plan, row summaries, fill meta, host staging; path B workspace; skeleton tangents
  DevBuf visited;                                      // path B: one bit per voxel, all zero between calls
  void* slice_state = nullptr;                         // xs_slice.cu's host-side plan
  void (*slice_state_free)(void*) = nullptr;
  uint64_t poly_cap2 = 0, item_cap2 = 0;
  cudaEvent_t evb[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // batch timing, launching stream
  cudaEvent_t evs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // batch timing, second stream
  cudaStream_t stream2 = nullptr;                      // the mid tier runs here, beside the warp tier
  cudaStream_t stream3 = nullptr;                      // third warp-tier CTA per SM (moves in when the mid-tier CTA retires)
  cudaEvent_t evf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // fork, warp-tier grid A done, second stream done, warp-tier grid B done, mid tier's first launch done
  SlabCfg cfg_mid{}, cfg_wide{};
  DevBuf slab_wide;
  int wide_ctas = 0;
  int mid_ctas = 0;
  int mid_key = 0;
  uint64_t poly_cap = 0, item_cap = 0;
  SlabCfg cfg1{}, cfg2{};
  uint32_t* h_counters = nullptr;   // pinned: CNT_WORDS counters + 8 u64 stats
  uint8_t* h_bits = nullptr;        // pinned staging of the bit-packed image (upload_packed_and_ingest)
  size_t h_bits_cap = 0;
  DevBuf staged[2];                 // bit-packed images uploaded ahead of their ingest (xs3d_volume_stage_packed)
  DevBuf staged_cubes[2];           // ... and the occupancy bytes built from them (swapped with `cubes` by commit_staged)
  bool staged_ready[2] = {false, false};
  uint8_t* h_stage_bits[2] = {nullptr, nullptr};   // pinned: where xs3d_volume_pack_and_stage packs
  size_t h_stage_cap[2] = {0, 0};
  cudaStream_t stage_stream = nullptr;
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  cudaEvent_t stage_copy_ev[2] = {nullptr, nullptr};   // recorded behind the last host-to-device copy of a slot: its host buffer may be rewritten
  bool stage_copy_pending[2] = {false, false};
  uint64_t stats[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t tier_stats[2] = {0, 0};   // samples / items handled by the CTA tiers
  uint32_t wide_over = 0;            // queries the wide tier passed on to the big tier
  bool expect_wide = false;          // the previous batch had sections that left the mid tier's window: launch the wide / big
                                     //   tiers in the first round of the next one (no host round trip in between)
  uint64_t prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int sm_count = 0, smem_optin = 0;
  bool kernels_configured = false;
  bool own_stream = true;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_block = nullptr;    // blocking host wait (XS3D_BLOCKING_SYNC)
  cudaEvent_t ev_done = nullptr;     // recorded behind a batch's result copies: the host waits for THIS, not for the stream
  void (*after_enqueue)(void*) = nullptr;   // xs3d_volume_set_after_enqueue
  void* after_enqueue_ctx = nullptr;
  float timing[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // ms, see xs3d_volume_timing
  int timing_pending = 0;   // 1: batch events not yet read, 2: same for a tiny (`single`) batch
};

// Wait for `st` on the host.  XS3D_BLOCKING_SYNC=1: sleep on a blocking event instead of spinning in the driver — for
// processes that share their cores with others (several ranks on one host while rank 0's pool packs the next image).
static cudaError_t host_wait(xs3d_volume* v, cudaStream_t st) {
  static const bool blocking = getenv("XS3D_BLOCKING_SYNC") != nullptr;
  static const int poll_us = getenv("XS3D_POLL_SLEEP_US") ? atoi(getenv("XS3D_POLL_SLEEP_US")) : 0;
  if (poll_us > 0) {
    // dev knob: poll with short sleeps instead of spinning (frees the hyperthread sibling for the packing pool)
    for (;;) {
      const cudaError_t q = cudaStreamQuery(st);
      if (q != cudaErrorNotReady) return q;
      timespec ts{0, poll_us * 1000L};
      nanosleep(&ts, nullptr);
    }
  }
  if (!blocking) return cudaStreamSynchronize(st);
  if (!v->ev_block) {
    cudaError_t e = cudaEventCreateWithFlags(&v->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(v->ev_block, st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(v->ev_block);
}

static VolView view_of(const xs3d_volume* v) {
  VolView r;
  r.cubes = (const uint8_t*)v->cubes.p;
  r.sx = v->sx; r.sy = v->sy; r.sz = v->sz; r.hx = v->hx; r.hy = v->hy; r.hz = v->hz;
  return r;
}

static int block_smem_bytes(const xs3d_volume* v) { return v->smem_optin; }
static int block_smem_bits_words(const xs3d_volume* v) { return v->smem_optin / 4 - BigTier::FIXED_WORDS; }

static int configure_kernels(xs3d_volume* v) {
  if (v->kernels_configured) return XS3D_OK;
  const int t0_bytes = T0_SMEM_BYTES;
  if (t0_bytes > v->smem_optin) return fail(XS3D_ERR_UNSUPPORTED, "device shared memory too small for the warp tier");
  CK(cudaFuncSetAttribute(T0_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, t0_bytes));
  CK(cudaFuncSetAttribute(area_block_kernel<BigTier>, cudaFuncAttributeMaxDynamicSharedMemorySize, block_smem_bytes(v)));
  CK(cudaFuncSetAttribute(area_mid_kernel<MidTier>, cudaFuncAttributeMaxDynamicSharedMemorySize, MidTier::SMEM_BYTES));
  if (WIDE_SMEM_BYTES > v->smem_optin) return fail(XS3D_ERR_UNSUPPORTED, "device shared memory too small for the wide tier");
  CK(cudaFuncSetAttribute(area_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WIDE_SMEM_BYTES));
  v->kernels_configured = true;
  return XS3D_OK;
}

extern "C" int xs3d_volume_create(int device, uint64_t sx, uint64_t sy, uint64_t sz, xs3d_volume** out) {
  *out = nullptr;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(XS3D_ERR_ARG, "no such CUDA device");
  if (sx > 60000 || sy > 60000 || sz > 60000) return fail(XS3D_ERR_UNSUPPORTED, "volume side > 60000");
  const uint64_t hx = (sx + 1) >> 1, hy = (sy + 1) >> 1, hz = (sz + 1) >> 1;
  if (hx * hy * hz >= 0xffffffffull) return fail(XS3D_ERR_UNSUPPORTED, "volume has >= 2^32 2x2x2 blocks");
  CK(cudaSetDevice(device));
  xs3d_volume* v = new xs3d_volume();
  v->device = device;
  v->sx = (int)sx; v->sy = (int)sy; v->sz = (int)sz; v->hx = (int)hx; v->hy = (int)hy; v->hz = (int)hz;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  v->sm_count = prop.multiProcessorCount;
  v->smem_optin = (int)prop.sharedMemPerBlockOptin;
  CK(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
  {
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&v->stream2, cudaStreamNonBlocking, prio_hi));   // latency-bound tier: its CTAs are placed first
    CK(cudaStreamCreateWithPriority(&v->stream3, cudaStreamNonBlocking, prio_lo));
  }
  for (int i = 0; i < 8; i++) CK(cudaEventCreate(&v->ev[i]));
  for (int i = 0; i < 5; i++) CK(cudaEventCreateWithFlags(&v->evf[i], cudaEventDisableTiming));
  for (int i = 0; i < 6; i++) CK(cudaEventCreate(&v->evb[i]));
  for (int i = 0; i < 6; i++) CK(cudaEventCreate(&v->evs[i]));
  RET(v->cubes.ensure(hx * hy * hz + 64));
  CK(cudaHostAlloc((void**)&v->h_counters, 512, cudaHostAllocDefault));
  *out = v;
  return XS3D_OK;
}

extern "C" int xs3d_volume_destroy(xs3d_volume* v) {
  if (!v) return XS3D_OK;
  cudaSetDevice(v->device);
  if (v->stream) { cudaStreamSynchronize(v->stream); if (v->own_stream) cudaStreamDestroy(v->stream); }
  if (v->stream2) { cudaStreamSynchronize(v->stream2); cudaStreamDestroy(v->stream2); }
  if (v->stream3) { cudaStreamSynchronize(v->stream3); cudaStreamDestroy(v->stream3); }
  if (v->stage_stream) { cudaStreamSynchronize(v->stage_stream); cudaStreamDestroy(v->stage_stream); }
  for (int i = 0; i < 2; i++) { if (v->stage_ev[i]) cudaEventDestroy(v->stage_ev[i]); if (v->stage_copy_ev[i]) cudaEventDestroy(v->stage_copy_ev[i]); v->staged[i].release(); v->staged_cubes[i].release(); if (v->h_stage_bits[i]) cudaFreeHost(v->h_stage_bits[i]); }
  for (int i = 0; i < 8; i++) if (v->ev[i]) cudaEventDestroy(v->ev[i]);
  if (v->ev_block) cudaEventDestroy(v->ev_block);
  for (int i = 0; i < 5; i++) if (v->evf[i]) cudaEventDestroy(v->evf[i]);
  for (int i = 0; i < 6; i++) if (v->evb[i]) cudaEventDestroy(v->evb[i]);
  for (int i = 0; i < 6; i++) if (v->evs[i]) cudaEventDestroy(v->evs[i]);
  DevBuf* bufs[] = { &v->cubes, &v->raw, &v->dbits, &v->labels, &v->q_pos, &v->q_nrm, &v->q_area, &v->q_contact, &v->lists,
                     &v->counters, &v->slab1, &v->slab2, &v->sec_idx, &v->sec_val, &v->misc,
                     &v->polys, &v->items, &v->vals, &v->geom, &v->qinfo, &v->slab_mid, &v->keys, &v->polys2, &v->items2, &v->vals2, &v->slab_wide,
                     &v->p2_order, &v->p2_mask, &v->p2_hash, &v->p2_chunks, &v->p2_recs };
  for (auto& g : v->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  for (DevBuf* b : bufs) b->release();
  for (DevBuf& b : v->slice_bufs) b.release();
  v->visited.release();
  if (v->slice_state && v->slice_state_free) v->slice_state_free(v->slice_state);
  if (v->h_counters) cudaFreeHost(v->h_counters);
  if (v->h_bits) cudaFreeHost(v->h_bits);
  delete v;
  return XS3D_OK;
}

extern "C" int xs3d_volume_shape(xs3d_volume* v, uint64_t shape[3]) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  shape[0] = v->sx; shape[1] = v->sy; shape[2] = v->sz;
  return XS3D_OK;
}

// uint8 image in HBM -> bit stream (bit i <-> byte i != 0), in the image's own linear order: the device-side twin of
// xs_pack_bits for images whose fastest side is not a multiple of 16 (the 16-byte-row kernels above do not apply; the
// any-size bit ingest kernels then build the occupancy bytes).  32 voxels per thread: two 16-byte streaming loads, one
// 4-byte store; purely linear, so it runs at copy speed whatever the shape.  `img` must be 16-byte aligned.
__global__ void __launch_bounds__(256) pack_bits_device_kernel(const uint8_t* __restrict__ img, size_t voxels, uint32_t* __restrict__ bits) {
  const size_t words = (voxels + 31) / 32;
  for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < words; wi += (size_t)gridDim.x * blockDim.x) {
    const size_t base = wi * 32;
    uint32_t m = 0;
    if (base + 32 <= voxels) {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(img + base));
      const uint4 b = __ldcs(reinterpret_cast<const uint4*>(img + base + 16));
      const uint32_t w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
#pragma unroll
      for (int j = 0; j < 8; j++) m |= ((((__vcmpne4(w[j], 0u) & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * j);
    } else {
      for (size_t i = base; i < voxels; i++) m |= (uint32_t)(img[i] != 0) << (i - base);
    }
    bits[wi] = m;
  }
}

static int launch_ingest_bits_to(xs3d_volume* v, const uint8_t* d_bits, int c_order, uint8_t* dst, cudaStream_t stream);

// C-ordered 1-byte image (sz % 16 == 0, 16-byte aligned) -> occupancy bytes through the TMA tile kernel.  Returns false
// when the tensor map cannot be made (then the caller uses ingest_c16_kernel); XS3D_NO_TMA=1 forces that too.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool launch_ingest_c_tma(xs3d_volume* v, const uint8_t* d_img, uint8_t* cubes, uint32_t want4, int eq, cudaStream_t stream) {
  static const bool off = getenv("XS3D_NO_TMA") != nullptr;
  if (off) return false;
  static TensorMapEncodeFn encode = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) encode = (TensorMapEncodeFn)fn;
    if (cudaFuncSetAttribute(ingest_c_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(ingest_c_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES) != cudaSuccess) encode = nullptr;
    cudaGetLastError();
  }
  if (!encode) return false;
  CUtensorMap tmap;
  const cuuint64_t dims[3] = { (cuuint64_t)v->sz, (cuuint64_t)v->sy, (cuuint64_t)v->sx };           // innermost first: z, y, x
  const cuuint64_t strides[2] = { (cuuint64_t)v->sz, (cuuint64_t)v->sz * (cuuint64_t)v->sy };       // bytes, for y and x
  const cuuint32_t box[3] = { 2 * BC_BZ, 2, 2 * BC_BX };
  const cuuint32_t estr[3] = { 1, 1, 1 };
  if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  const size_t tiles = (size_t)((v->hx + BC_BX - 1) / BC_BX) * ((v->hz + BC_BZ - 1) / BC_BZ) * v->hy;
  const int grid = (int)std::min<size_t>(tiles, (size_t)v->sm_count * 3);
  if (eq) ingest_c_tma_kernel<true><<<grid, 256, TMA_SMEM_BYTES, stream>>>(tmap, cubes, v->hx, v->hy, v->hz, want4);
  else ingest_c_tma_kernel<false><<<grid, 256, TMA_SMEM_BYTES, stream>>>(tmap, cubes, v->hx, v->hy, v->hz, want4);
  return cudaGetLastError() == cudaSuccess;
}

static int launch_ingest(xs3d_volume* v, const uint8_t* d_img, int c_order) {
  const size_t nblocks = (size_t)v->hx * v->hy * v->hz;
  if (nblocks == 0) return XS3D_OK;
  const bool fast = !c_order && (v->sx % 16 == 0) && ((uintptr_t)d_img % 16 == 0);
  if (fast) {
    const size_t total = nblocks / 8;
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)v->sm_count * 32);
    ingest_f16_kernel<<<grid, 256, 0, v->stream>>>(d_img, (uint8_t*)v->cubes.p, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz);
  } else if (c_order && (v->sz % 16 == 0) && ((uintptr_t)d_img % 16 == 0) && launch_ingest_c_tma(v, d_img, (uint8_t*)v->cubes.p, 0u, 0, v->stream)) {
    // (launched)
  } else if (c_order && (v->sz % 16 == 0) && ((uintptr_t)d_img % 16 == 0)) {
    const size_t tiles = (size_t)((v->hx + BC_BX - 1) / BC_BX) * ((v->hz + BC_BZ - 1) / BC_BZ) * v->hy;
    ingest_c16_kernel<<<(int)std::min<size_t>(tiles, (size_t)v->sm_count * 32), 256, 0, v->stream>>>(d_img, (uint8_t*)v->cubes.p, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, 0u, 0);
  } else if ((uintptr_t)d_img % 16 == 0 && nblocks >= 4096) {
    // awkward sizes: bytes -> bits at copy speed, then the any-size bit ingest (1.25x the traffic of the direct kernels,
    // against the one-thread-per-byte kernel's ~10 % of the HBM peak)
    const size_t voxels = (size_t)v->sx * v->sy * v->sz;
    RET(v->dbits.ensure(((voxels + 31) / 32) * 4 + 64));
    const size_t words = (voxels + 31) / 32;
    pack_bits_device_kernel<<<(int)std::min<size_t>((words + 255) / 256, (size_t)v->sm_count * 32), 256, 0, v->stream>>>(d_img, voxels, (uint32_t*)v->dbits.p);
    CK(cudaGetLastError());
    RET(launch_ingest_bits_to(v, (const uint8_t*)v->dbits.p, c_order, (uint8_t*)v->cubes.p, v->stream));
  } else {
    int grid = (int)std::min<size_t>((nblocks + 255) / 256, (size_t)v->sm_count * 32);
    if (c_order)
      ingest_generic_kernel<true><<<grid, 256, 0, v->stream>>>(d_img, (uint8_t*)v->cubes.p, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz,
                                                                (long long)v->sy * v->sz, (long long)v->sz, 1ll);
    else
      ingest_generic_kernel<false><<<grid, 256, 0, v->stream>>>(d_img, (uint8_t*)v->cubes.p, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz,
                                                                 1ll, (long long)v->sx, (long long)v->sx * v->sy);
  }
  CK(cudaGetLastError());
  return XS3D_OK;
}

extern "C" int xs_pack_bits_parts(const uint8_t* in, size_t lo, size_t hi, uint8_t* out, int nparts,
                                  void (*done)(void* ctx, size_t lo, size_t hi), void* ctx);   // xs_hostpack.cpp

extern "C" int xs_pack_threads(void);                                                 // xs_hostpack.cpp

// Host image -> occupancy bytes through the bit-packed upload: the host cores pack the image to one bit per voxel
// (one pool job, its parts sent as they complete, so that packing and the transfers overlap), 1/8 of the bytes cross PCIe, the device builds
// the occupancy bytes from the bits.  Worth it from a few MiB up, and only with enough host threads to out-run the link
// (measured on a B200 box: 8 threads pack 94 GB/s, 14 threads 130 GB/s; the link carries 54 GB/s); otherwise the image goes up as it is.
// XS3D_NO_PACK=1 / XS3D_FORCE_PACK=1 override.
constexpr size_t PACK_MIN_VOXELS = (size_t)8 << 20;
constexpr int PACK_MIN_THREADS = 6;
static bool use_packed_upload(size_t voxels) {
  static const bool off = getenv("XS3D_NO_PACK") != nullptr;
  static const bool force = getenv("XS3D_FORCE_PACK") != nullptr;
  if (off || voxels < PACK_MIN_VOXELS) return false;
  return force || xs_pack_threads() >= PACK_MIN_THREADS;
}
// dev aid: XS3D_TRACE=1 prints host-side timestamps (microseconds) of the upload / launch / wait phases to stderr
static void trace(const xs3d_volume* v, const char* tag) {
  static const bool on = getenv("XS3D_TRACE") != nullptr;
  if (!on) return;
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  fprintf(stderr, "xs3d-trace %p %s %.1f\n", (const void*)v, tag, ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3);
}

// bit-packed image in device memory -> occupancy bytes `dst`, on `stream`
static int launch_ingest_bits_to(xs3d_volume* v, const uint8_t* d_bits, int c_order, uint8_t* dst, cudaStream_t stream) {
  const size_t nblocks = (size_t)v->hx * v->hy * v->hz;
  const uintptr_t al = (uintptr_t)d_bits;
  if (!c_order && (v->sx % 32 == 0) && al % 4 == 0) {
    const size_t total = nblocks / 16;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)v->sm_count * 32);
    ingest_bits_f32_kernel<<<grid, 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz);
  } else if (!c_order && (v->sx % 16 == 0) && al % 2 == 0) {
    const size_t total = nblocks / 8;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)v->sm_count * 32);
    ingest_bits_f16_kernel<<<grid, 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz);
  } else if (c_order && al % 4 == 0) {
    const size_t tiles = (size_t)((v->hx + BC_BX - 1) / BC_BX) * ((v->hz + BC_BZ - 1) / BC_BZ) * v->hy;
    ingest_bits_c_kernel<<<(int)std::min<size_t>(tiles, (size_t)v->sm_count * 32), 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz);
  } else if (!c_order) {
    const size_t total = (size_t)((v->hx + 7) / 8) * v->hy * v->hz;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)v->sm_count * 32);
    ingest_bits_fx_kernel<<<grid, 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz);
  } else {
    const int grid = (int)std::min<size_t>((nblocks + 255) / 256, (size_t)v->sm_count * 32);
    if (c_order)
      ingest_bits_generic_kernel<<<grid, 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz,
                                                           (long long)v->sy * v->sz, (long long)v->sz, 1ll);
    else
      ingest_bits_generic_kernel<<<grid, 256, 0, stream>>>(d_bits, dst, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz,
                                                           1ll, (long long)v->sx, (long long)v->sx * v->sy);
  }
  CK(cudaGetLastError());
  return XS3D_OK;
}
// ... into the volume's occupancy bytes on its launching stream (timed by ev[6], ev[7] like the byte ingest)
static int launch_ingest_bits(xs3d_volume* v, const uint8_t* d_bits, int c_order) {
  CK(cudaEventRecord(v->ev[6], v->stream));
  RET(launch_ingest_bits_to(v, d_bits, c_order, (uint8_t*)v->cubes.p, v->stream));
  CK(cudaEventRecord(v->ev[7], v->stream));
  return XS3D_OK;
}

struct PackUpload { const uint8_t* h_bits; uint8_t* d_bits; cudaStream_t stream; cudaError_t err; };
// a part of the bit stream is ready on the host: send it (runs on the thread that called xs_pack_bits_parts)
static void pack_part_ready(void* ctx, size_t lo, size_t hi) {
  PackUpload* u = (PackUpload*)ctx;
  if (u->err != cudaSuccess) return;
  u->err = cudaMemcpyAsync(u->d_bits + lo / 8, u->h_bits + lo / 8, (hi - lo + 7) / 8, cudaMemcpyHostToDevice, u->stream);
}

static int upload_packed_and_ingest(xs3d_volume* v, const uint8_t* binimg, int c_order) {
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  const size_t nbytes = (voxels + 7) / 8;
  if (v->h_bits_cap < nbytes + 64) {
    if (v->h_bits) cudaFreeHost(v->h_bits);
    v->h_bits = nullptr; v->h_bits_cap = 0;
    CK(cudaHostAlloc((void**)&v->h_bits, nbytes + 4096, cudaHostAllocDefault));
    v->h_bits_cap = nbytes + 4096;
  }
  RET(v->raw.ensure(nbytes + 4096));
  uint8_t* d_bits = (uint8_t*)v->raw.p;
  // one job on the host pool, reported in eight parts (256 KiB-grain aligned, hence byte aligned in the bit stream):
  // each part's copy is enqueued as soon as it is packed, so only the last 2 MiB transfer is exposed
  trace(v, "pack-begin");
  PackUpload up{ v->h_bits, d_bits, v->stream, cudaSuccess };
  xs_pack_bits_parts(binimg, 0, voxels, v->h_bits, 8, pack_part_ready, &up);
  CK(up.err);
  trace(v, "pack-end");
  RET(launch_ingest_bits(v, d_bits, c_order));
  return XS3D_OK;
}

extern "C" int xs3d_volume_load(xs3d_volume* v, const uint8_t* binimg, int mem, int c_order) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) { v->loaded = true; return XS3D_OK; }
  if (!binimg) return fail(XS3D_ERR_ARG, "null image");
  const uint8_t* d_img = binimg;
  if (mem == XS3D_HOST && use_packed_upload(voxels)) {
    RET(upload_packed_and_ingest(v, binimg, c_order));
  } else {
    if (mem == XS3D_HOST) {
      RET(v->raw.ensure(voxels + 64));
      CK(cudaMemcpyAsync(v->raw.p, binimg, voxels, cudaMemcpyHostToDevice, v->stream));
      d_img = (const uint8_t*)v->raw.p;
    }
    CK(cudaEventRecord(v->ev[6], v->stream));
    RET(launch_ingest(v, d_img, c_order));
    CK(cudaEventRecord(v->ev[7], v->stream));
  }
  CK(cudaStreamSynchronize(v->stream));
  CK(cudaEventElapsedTime(&v->timing[4], v->ev[6], v->ev[7]));
  v->loaded = true;
  return XS3D_OK;
}

// The two halves of the bit-packed upload as separate calls (a pipeline can pack image k+1 on the host while image k is
// being swept, or hold bit-packed masks in the first place).
extern "C" int xs3d_pack_bits(const uint8_t* binimg, uint64_t voxels, uint8_t* bits) {
  if (voxels && (!binimg || !bits)) return fail(XS3D_ERR_ARG, "null argument");
  xs_pack_bits_parts(binimg, 0, (size_t)voxels, bits, 1, nullptr, nullptr);
  return XS3D_OK;
}

extern "C" int xs3d_volume_load_packed(xs3d_volume* v, const uint8_t* bits, int mem, int c_order) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) { v->loaded = true; return XS3D_OK; }
  if (!bits) return fail(XS3D_ERR_ARG, "null image");
  const uint8_t* d_bits = bits;
  if (mem == XS3D_HOST) {
    const size_t nbytes = (voxels + 7) / 8;
    RET(v->raw.ensure(nbytes + 4096));
    CK(cudaMemcpyAsync(v->raw.p, bits, nbytes, cudaMemcpyHostToDevice, v->stream));
    d_bits = (const uint8_t*)v->raw.p;
  }
  RET(launch_ingest_bits(v, d_bits, c_order));
  CK(cudaStreamSynchronize(v->stream));
  CK(cudaEventElapsedTime(&v->timing[4], v->ev[6], v->ev[7]));
  v->loaded = true;
  return XS3D_OK;
}

// Upload AND ingest ahead of use: the copy and the bits -> occupancy-byte kernel run on the volume's own upload stream
// (beside whatever the launching stream is doing, e.g. the sweep of the previous image) into one of two spare sets of
// buffers; xs3d_volume_commit_staged then only makes the launching stream wait for that work and swaps the spare
// occupancy bytes in.  stage may be called from another host thread than commit (the caller orders the two calls of a
// slot; a slot may be staged again once it has been committed).
static int stage_prepare(xs3d_volume* v, int slot, size_t nbytes) {
  if (!v->stage_stream) {
    CK(cudaStreamCreateWithFlags(&v->stage_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&v->stage_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&v->stage_copy_ev[i], cudaEventDisableTiming));
  }
  RET(v->staged[slot].ensure(nbytes + 4096));
  RET(v->staged_cubes[slot].ensure((size_t)v->hx * v->hy * v->hz + 64));
  return XS3D_OK;
}
// The previous image staged through this slot may still be on its way up: its host buffer (the volume's pinned buffer,
// or the caller's) must not be rewritten before those copies are done.  Called at the top of the two staging calls.
static int stage_wait_copies(xs3d_volume* v, int slot) {
  if (v->stage_copy_pending[slot]) {
    CK(cudaEventSynchronize(v->stage_copy_ev[slot]));
    v->stage_copy_pending[slot] = false;
  }
  return XS3D_OK;
}
static int stage_finish(xs3d_volume* v, int slot, int c_order) {
  CK(cudaEventRecord(v->stage_copy_ev[slot], v->stage_stream));
  v->stage_copy_pending[slot] = true;
  RET(launch_ingest_bits_to(v, (const uint8_t*)v->staged[slot].p, c_order, (uint8_t*)v->staged_cubes[slot].p, v->stage_stream));
  CK(cudaEventRecord(v->stage_ev[slot], v->stage_stream));
  v->staged_ready[slot] = true;
  return XS3D_OK;
}

extern "C" int xs3d_volume_stage_packed(xs3d_volume* v, const uint8_t* bits, int slot, int c_order) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) return XS3D_OK;
  if (!bits) return fail(XS3D_ERR_ARG, "null image");
  const size_t nbytes = (voxels + 7) / 8;
  RET(stage_wait_copies(v, slot));
  RET(stage_prepare(v, slot, nbytes));
  // in 1 MiB pieces: the small host-to-device copies of a batch running meanwhile (its queries) then wait for one
  // piece (~20 us) on the copy engine, not for the whole image
  for (size_t off = 0; off < nbytes; off += (size_t)1 << 20)
    CK(cudaMemcpyAsync((uint8_t*)v->staged[slot].p + off, bits + off, std::min<size_t>((size_t)1 << 20, nbytes - off), cudaMemcpyHostToDevice, v->stage_stream));
  return stage_finish(v, slot, c_order);
}

// The same for a boolean image that is still bytes: packed by the host pool into a pinned buffer of the volume, every
// eighth of the bit stream copied (upload stream) as soon as it is packed — the transfer ends with the packing.
extern "C" int xs3d_volume_pack_and_stage(xs3d_volume* v, const uint8_t* binimg, int slot, int c_order) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) return XS3D_OK;
  if (!binimg) return fail(XS3D_ERR_ARG, "null image");
  const size_t nbytes = (voxels + 7) / 8;
  RET(stage_wait_copies(v, slot));
  RET(stage_prepare(v, slot, nbytes));
  if (v->h_stage_cap[slot] < nbytes + 64) {
    if (v->h_stage_bits[slot]) cudaFreeHost(v->h_stage_bits[slot]);
    v->h_stage_bits[slot] = nullptr; v->h_stage_cap[slot] = 0;
    CK(cudaHostAlloc((void**)&v->h_stage_bits[slot], nbytes + 4096, cudaHostAllocDefault));
    v->h_stage_cap[slot] = nbytes + 4096;
  }
  PackUpload up{ v->h_stage_bits[slot], (uint8_t*)v->staged[slot].p, v->stage_stream, cudaSuccess };
  xs_pack_bits_parts(binimg, 0, voxels, v->h_stage_bits[slot], 8, pack_part_ready, &up);
  CK(up.err);
  return stage_finish(v, slot, c_order);
}

// Several processes (one per GPU) that can all read ONE host image (shared memory, or a replica each): every process
// packs and uploads only its part of the bit stream, the parts are all-gathered between the GPUs by the host layer on a
// stream of its own, and each GPU builds its occupancy bytes from the whole stream — the packing (the host-memory-bound
// part of the end-to-end step) is spread over the cores of every rank instead of the source rank's.
//   staged_bits        : the slot's device bit buffer (at least min_bytes long: the all-gather wants world * part bytes)
//   pack_and_stage_part: voxels [lo, hi) -> bits, copied to their place in that buffer on the upload stream
//   wait_staged_bits   : `cuda_stream` waits for that copy (then runs the all-gather into the buffer)
//   ingest_staged      : bits -> spare occupancy bytes on `cuda_stream`, behind the all-gather; marks the slot staged
extern "C" int xs3d_volume_staged_bits(xs3d_volume* v, int slot, uint64_t min_bytes, void** device_ptr, uint64_t* bytes) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1 || !device_ptr || !bytes) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  const size_t nbytes = (voxels + 7) / 8;
  RET(stage_prepare(v, slot, std::max<size_t>(nbytes, (size_t)min_bytes)));
  *device_ptr = v->staged[slot].p;
  *bytes = (uint64_t)nbytes;
  return XS3D_OK;
}
extern "C" int xs3d_volume_pack_and_stage_part(xs3d_volume* v, const uint8_t* binimg, int slot, uint64_t lo_voxel, uint64_t hi_voxel) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) return XS3D_OK;
  if (!binimg) return fail(XS3D_ERR_ARG, "null image");
  if ((hi_voxel > lo_voxel && lo_voxel % 4096 != 0) || hi_voxel > voxels || lo_voxel > hi_voxel) return fail(XS3D_ERR_ARG, "part must start at a multiple of 4096 voxels and end inside the image");
  const size_t nbytes = (voxels + 7) / 8;
  RET(stage_wait_copies(v, slot));
  RET(stage_prepare(v, slot, nbytes));
  if (v->h_stage_cap[slot] < nbytes + 64) {
    if (v->h_stage_bits[slot]) cudaFreeHost(v->h_stage_bits[slot]);
    v->h_stage_bits[slot] = nullptr; v->h_stage_cap[slot] = 0;
    CK(cudaHostAlloc((void**)&v->h_stage_bits[slot], nbytes + 4096, cudaHostAllocDefault));
    v->h_stage_cap[slot] = nbytes + 4096;
  }
  // the ingest of the image that went through this slot before read the device bit buffer on the caller's stream
  CK(cudaStreamWaitEvent(v->stage_stream, v->stage_ev[slot], 0));
  if (hi_voxel > lo_voxel) {
    PackUpload up{ v->h_stage_bits[slot], (uint8_t*)v->staged[slot].p, v->stage_stream, cudaSuccess };
    xs_pack_bits_parts(binimg, (size_t)lo_voxel, (size_t)hi_voxel, v->h_stage_bits[slot], 4, pack_part_ready, &up);
    CK(up.err);
  }
  CK(cudaEventRecord(v->stage_copy_ev[slot], v->stage_stream));
  v->stage_copy_pending[slot] = true;
  return XS3D_OK;
}
extern "C" int xs3d_volume_wait_staged_bits(xs3d_volume* v, int slot, void* cuda_stream) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1 || !v->stage_stream) return fail(XS3D_ERR_ARG, "nothing staged in this slot");
  CK(cudaSetDevice(v->device));
  CK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, v->stage_copy_ev[slot], 0));
  return XS3D_OK;
}
extern "C" int xs3d_volume_ingest_staged(xs3d_volume* v, int slot, int c_order, void* cuda_stream) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1 || !v->stage_stream) return fail(XS3D_ERR_ARG, "slot not prepared (xs3d_volume_staged_bits first)");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) return XS3D_OK;
  RET(launch_ingest_bits_to(v, (const uint8_t*)v->staged[slot].p, c_order, (uint8_t*)v->staged_cubes[slot].p, (cudaStream_t)cuda_stream));
  CK(cudaEventRecord(v->stage_ev[slot], (cudaStream_t)cuda_stream));
  v->staged_ready[slot] = true;
  return XS3D_OK;
}

// For a host layer that fills the spare occupancy bytes of a slot itself (an NCCL broadcast into them on its own stream):
// the buffer, the "wait for what was staged" hook for that stream, and the call that marks the slot as staged.
extern "C" int xs3d_volume_staged_cubes(xs3d_volume* v, int slot, void** device_ptr, uint64_t* bytes) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  RET(stage_prepare(v, slot, (voxels + 7) / 8));
  *device_ptr = v->staged_cubes[slot].p;
  *bytes = (uint64_t)v->hx * v->hy * v->hz;
  return XS3D_OK;
}
extern "C" int xs3d_volume_wait_staged(xs3d_volume* v, int slot, void* cuda_stream) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1 || !v->stage_stream) return fail(XS3D_ERR_ARG, "nothing staged in this slot");
  CK(cudaSetDevice(v->device));
  CK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, v->stage_ev[slot], 0));
  return XS3D_OK;
}
// The spare occupancy bytes of `slot` were written on `cuda_stream` (everything enqueued there so far): commit_staged
// will wait for exactly that.
extern "C" int xs3d_volume_mark_staged(xs3d_volume* v, int slot, void* cuda_stream) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1 || !v->stage_stream) return fail(XS3D_ERR_ARG, "slot not prepared (xs3d_volume_staged_cubes first)");
  CK(cudaSetDevice(v->device));
  CK(cudaEventRecord(v->stage_ev[slot], (cudaStream_t)cuda_stream));
  v->staged_ready[slot] = true;
  return XS3D_OK;
}

extern "C" int xs3d_volume_commit_staged(xs3d_volume* v, int slot) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (slot < 0 || slot > 1) return fail(XS3D_ERR_ARG, "slot must be 0 or 1");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0) { v->loaded = true; return XS3D_OK; }
  if (!v->stage_stream || !v->staged_ready[slot]) return fail(XS3D_ERR_ARG, "nothing staged in this slot");
  // every earlier batch has finished with the current occupancy bytes (the batch calls return after a synchronize), so
  // the buffers can simply change places; what the launching stream does next waits for the staged ingest
  CK(cudaStreamWaitEvent(v->stream, v->stage_ev[slot], 0));
  std::swap(v->cubes, v->staged_cubes[slot]);
  v->staged_ready[slot] = false;
  v->loaded = true;
  return XS3D_OK;
}

extern "C" int xs3d_volume_select_label(xs3d_volume* v, uint64_t label) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (!v->labels_loaded) return fail(XS3D_ERR_ARG, "volume has no labels loaded");
  CK(cudaSetDevice(v->device));
  const size_t nblocks = (size_t)v->hx * v->hy * v->hz;
  if (nblocks == 0) { v->loaded = true; return XS3D_OK; }
  const long long stx = v->labels_c_order ? (long long)v->sy * v->sz : 1ll;
  const long long sty = v->labels_c_order ? (long long)v->sz : (long long)v->sx;
  const long long stz = v->labels_c_order ? 1ll : (long long)v->sx * v->sy;
  const int grid = (int)std::min<size_t>((nblocks + 255) / 256, (size_t)v->sm_count * 32);
  uint8_t* cubes = (uint8_t*)v->cubes.p;
  CK(cudaEventRecord(v->ev[6], v->stream));
  static const bool label_generic = getenv("XS3D_LABEL_GENERIC") != nullptr;      // dev aid: force the one-thread-per-byte kernel
  const int vpt = 16 / std::max(1, v->itemsize);
  if (!label_generic && !v->labels_c_order && v->sx % vpt == 0 && (uintptr_t)v->labels.p % 16 == 0) {
    const size_t total = nblocks / (size_t)(vpt / 2);
    const int fgrid = (int)std::min<size_t>((total + 255) / 256, (size_t)v->sm_count * 32);
    switch (v->itemsize) {
      case 1: ingest_label_f16_kernel<uint8_t><<<fgrid, 256, 0, v->stream>>>((const uint8_t*)v->labels.p, (uint8_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
      case 2: ingest_label_f16_kernel<uint16_t><<<fgrid, 256, 0, v->stream>>>((const uint16_t*)v->labels.p, (uint16_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
      case 4: ingest_label_f16_kernel<uint32_t><<<fgrid, 256, 0, v->stream>>>((const uint32_t*)v->labels.p, (uint32_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
      default: ingest_label_f16_kernel<unsigned long long><<<fgrid, 256, 0, v->stream>>>((const unsigned long long*)v->labels.p, (unsigned long long)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
    }
  } else if (!label_generic && v->labels_c_order && (uintptr_t)v->labels.p % 16 == 0 && v->sz % (v->itemsize == 1 ? 16 : 8) == 0) {
    const size_t tiles = (size_t)((v->hx + BC_BX - 1) / BC_BX) * ((v->hz + BC_BZ - 1) / BC_BZ) * v->hy;
    const int cgrid = (int)std::min<size_t>(tiles, (size_t)v->sm_count * 32);
    switch (v->itemsize) {
      case 1:
        if (!launch_ingest_c_tma(v, (const uint8_t*)v->labels.p, cubes, (uint32_t)(label & 0xff) * 0x01010101u, 1, v->stream))
          ingest_c16_kernel<<<cgrid, 256, 0, v->stream>>>((const uint8_t*)v->labels.p, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, (uint32_t)(label & 0xff) * 0x01010101u, 1);
        break;
      case 2: ingest_label_c_kernel<uint16_t><<<cgrid, 256, 0, v->stream>>>((const uint16_t*)v->labels.p, (uint16_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
      case 4: ingest_label_c_kernel<uint32_t><<<cgrid, 256, 0, v->stream>>>((const uint32_t*)v->labels.p, (uint32_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
      default: ingest_label_c_kernel<unsigned long long><<<cgrid, 256, 0, v->stream>>>((const unsigned long long*)v->labels.p, (unsigned long long)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz); break;
    }
  } else
  switch (v->itemsize) {
    case 1: ingest_label_kernel<uint8_t><<<grid, 256, 0, v->stream>>>((const uint8_t*)v->labels.p, (uint8_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, stx, sty, stz); break;
    case 2: ingest_label_kernel<uint16_t><<<grid, 256, 0, v->stream>>>((const uint16_t*)v->labels.p, (uint16_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, stx, sty, stz); break;
    case 4: ingest_label_kernel<uint32_t><<<grid, 256, 0, v->stream>>>((const uint32_t*)v->labels.p, (uint32_t)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, stx, sty, stz); break;
    default: ingest_label_kernel<unsigned long long><<<grid, 256, 0, v->stream>>>((const unsigned long long*)v->labels.p, (unsigned long long)label, cubes, v->sx, v->sy, v->sz, v->hx, v->hy, v->hz, stx, sty, stz); break;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(v->ev[7], v->stream));
  CK(cudaStreamSynchronize(v->stream));
  CK(cudaEventElapsedTime(&v->timing[4], v->ev[6], v->ev[7]));
  v->loaded = true;
  return XS3D_OK;
}

extern "C" int xs3d_volume_cubes(xs3d_volume* v, void** device_ptr, uint64_t* bytes) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  *device_ptr = v->cubes.p;
  *bytes = (uint64_t)v->hx * v->hy * v->hz;
  return XS3D_OK;
}
extern "C" int xs3d_volume_mark_loaded(xs3d_volume* v) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  v->loaded = true;
  return XS3D_OK;
}

// Slab sizing.  T1: moderate per-CTA slabs (one per SM).  T2: a single worst-case slab.
static SlabCfg make_cfg(const xs3d_volume* v, bool worst_case) {
  SlabCfg s{};
  const long long m = std::max(v->sx, std::max(v->sy, v->sz));
  const long long psx = 2 * 97 * m / 56 + 1;
  const size_t win = (size_t)(psx + 2 + 64);
  s.bits_words = ((win + 31) / 32 + 1) * (win + 32);
  s.tested_words = (((win + 31) / 32) * ((win + 31) / 32) + 31) / 32 + 1;
  if (!worst_case) {
    // big tier, one slab per SM: large enough for the biggest planar section of volumes up to ~600^3 (so the
    // reference's perf.cpp-style sweeps through solid volumes run 148 planes at a time), capped at 2^19 samples
    unsigned long long mn = (unsigned long long)(1.5 * (double)m * (double)m) + 16ull * m + 4096ull;
    if (mn > (1ull << 19)) mn = 1ull << 19;
    if (mn < (1ull << 16)) mn = 1ull << 16;
    s.max_n = (int)mn;
    unsigned long long cap = 1024;
    while (cap < 4 * mn) cap <<= 1;
    s.hcap_max = (uint32_t)cap;
    s.hash_factor = 4;
    s.max_probe = 1024;
  } else {
    // the largest planar section of a box is < 1.5 * m^2 lattice cells (plus rim)
    unsigned long long mn = (unsigned long long)(1.5 * (double)m * (double)m) + 16ull * m + 4096ull;
    if (mn > 0x7fffffffull / 32) mn = 0x7fffffffull / 32;
    s.max_n = (int)mn;
    // the same table serves path A (keys = blocks) and path B (keys = voxels): never more than 27 per sample
    const unsigned long long nvoxels = (unsigned long long)v->sx * v->sy * v->sz;
    const unsigned long long keys = std::min<unsigned long long>(27ull * mn, nvoxels + 1);
    unsigned long long cap = 1024;
    while (cap < 2 * keys) cap <<= 1;
    s.hcap_max = (uint32_t)std::min<unsigned long long>(cap, 1ull << 31);
    s.hash_factor = 64;
    s.max_probe = 1 << 30;     // the table can never fill; the bound only turns a logic error into a loud failure
  }
  return s;
}

// words of the first-touch table pool: 2 x pow2(>= factor x samples) per parked section, at least 2 x 1024
static uint64_t p2_hash_words(const xs3d_volume* v) {
  const uint64_t w = std::max<uint64_t>(v->p2_order_cap * 4ull * (uint64_t)v->p2_factor + 4096ull * v->p2_rec_cap, v->p2_hash_min_words);
  return std::min<uint64_t>(w, 0xffffffffull);
}

static int ensure_batch_workspace(xs3d_volume* v, uint64_t n) {
  RET(configure_kernels(v));
  RET(v->q_pos.ensure(n * 12 + 64));
  RET(v->q_nrm.ensure(n * 12 + 64));
  RET(v->q_area.ensure(n * 4 + 64));
  RET(v->q_contact.ensure(n + 64));
  RET(v->lists.ensure(n * 7 * 4 + 64));
  RET(v->keys.ensure(n + 64));
  RET(v->counters.ensure(CNT_TOTAL_WORDS * 4));
  if (!v->slab_mid.p) {
    SlabCfg m{};
    m.max_n = 1 << 14; m.hcap_max = 0; m.bits_words = 0; m.tested_words = 0; m.hash_factor = 8; m.max_probe = 256;   // (no per-CTA table: P2 has the sections' tables)
    m.stride = mid_slab_bytes(m);
    v->mid_ctas = v->sm_count;
    if (const char* e = getenv("XS3D_MID_CTAS")) v->mid_ctas = std::max(1, atoi(e));          // tuning knobs (dev)
    if (const char* e = getenv("XS3D_MID_KEY")) v->mid_key = std::min(255, std::max(1, atoi(e)));
    RET(v->slab_mid.ensure(m.stride * (size_t)v->mid_ctas));
    m.base = (char*)v->slab_mid.p;
    v->cfg_mid = m;
  }
  if (!v->slab_wide.p) {
    SlabCfg m{};
    m.max_n = 1 << 17; m.hcap_max = 1u << 19; m.bits_words = 0; m.tested_words = 0; m.hash_factor = 4; m.max_probe = 1024;
    m.stride = slab_bytes(m);
    v->wide_ctas = std::min(v->sm_count, 32);
    RET(v->slab_wide.ensure(m.stride * (size_t)v->wide_ctas));
    m.base = (char*)v->slab_wide.p;
    v->cfg_wide = m;
  }
  RET(v->geom.ensure(n * sizeof(PlaneGeom) + 64));
  RET(v->qinfo.ensure(n * sizeof(QInfo) + 64));
  if (v->poly_cap == 0) {
    // first guess: ~1k polygons per query (a neurite-scale section needs 200-600); grown on demand
    v->poly_cap = std::min<uint64_t>(std::max<uint64_t>(1ull << 20, n * 1024ull), 1ull << 26);
    v->item_cap = v->poly_cap / 2;
  }
  RET(v->polys.ensure(v->poly_cap * sizeof(PolyRec) + 64));
  RET(v->vals.ensure(v->poly_cap * 4 + 64));
  RET(v->items.ensure(v->item_cap * sizeof(ItemRec) + 64));
  if (v->poly_cap2 == 0) {
    v->poly_cap2 = std::min<uint64_t>(std::max<uint64_t>(1ull << 19, n * 128ull), 1ull << 26);
    v->item_cap2 = v->poly_cap2 / 2;
  }
  RET(v->polys2.ensure(v->poly_cap2 * sizeof(PolyRec) + 64));
  RET(v->vals2.ensure(v->poly_cap2 * 4 + 64));
  RET(v->items2.ensure(v->item_cap2 * sizeof(ItemRec) + 64));
  {
    // P2 pools: room for the largest section the volume can hold plus a sweep's worth of mid-tier sections; a full
    // pool doubles and the batch runs again (like the record buffers)
    const long long m = std::max(v->sx, std::max(v->sy, v->sz));
    const uint64_t worst_n = (uint64_t)(1.5 * (double)m * (double)m) + 16ull * m + 4096ull;
    if (v->p2_order_cap == 0) {
      v->p2_order_cap = std::max<uint64_t>(1ull << 20, worst_n + (1ull << 19));
      // dev / test knobs: start with small pools or sparse tables so that the grow-and-rerun paths are exercised
      if (const char* e = getenv("XS3D_P2_ORDER_CAP")) v->p2_order_cap = std::max<uint64_t>(256, strtoull(e, nullptr, 10));
      if (const char* e = getenv("XS3D_P2_FACTOR")) v->p2_factor = std::max(1, atoi(e));
    }
    v->p2_order_cap = std::min<uint64_t>(v->p2_order_cap, 0x7fffffffull);
    v->p2_rec_cap = std::max<uint64_t>(v->p2_rec_cap, n);
    RET(v->p2_order.ensure(v->p2_order_cap * 4 + 64));
    RET(v->p2_mask.ensure(v->p2_order_cap * 4 + 64));
    RET(v->p2_hash.ensure(p2_hash_words(v) * 4 + 64));
    RET(v->p2_chunks.ensure((v->p2_order_cap / P2_CHUNK + v->p2_rec_cap + 64) * 5 * 4 + 64));
    RET(v->p2_recs.ensure(v->p2_rec_cap * sizeof(P2Rec) + 64));
  }
  if (!v->slab1.p) {
    v->cfg1 = make_cfg(v, false);
    v->cfg1.stride = slab_bytes(v->cfg1);
    RET(v->slab1.ensure(v->cfg1.stride * (size_t)v->sm_count));
    v->cfg1.base = (char*)v->slab1.p;
  }
  return XS3D_OK;
}

static int ensure_worst_case_slab(xs3d_volume* v) {
  if (v->slab2.p) return XS3D_OK;
  v->cfg2 = make_cfg(v, true);
  v->cfg2.stride = slab_bytes(v->cfg2);
  RET(v->slab2.ensure(v->cfg2.stride));
  v->cfg2.base = (char*)v->slab2.p;
  return XS3D_OK;
}

// Core of the batched area path; all pointers are device pointers.  Two independent pipelines, no host
// synchronisation until the end:
//   launching stream: pre-pass (size estimate + ordering) -> warp tier -> polygon kernel -> combine kernel
//   second stream   : mid tier (queries classified as large + whatever the warp tier hands over while both run)
//                     -> mid tier again once the warp tier has ended (late hand-overs) -> big tier (what the mid tier
//                     could not fit) -> polygon kernel -> combine kernel
// Each pipeline has its own record buffers and emission counters; a query's QInfo.status says which pipeline owns it.
// Rare paths after the final sync: queries that overflowed the big tier's slab re-run with the worst-case slab; a
// full record buffer grows and the batch re-runs.  Tiny batches (`single`) go straight to the big tier.
static int area_batch_device(xs3d_volume* v, uint64_t n, const float* d_pos, const float* d_nrm, const float w[3],
                             float* d_area, uint8_t* d_contact, bool single, float* h_area = nullptr, uint8_t* h_contact = nullptr) {
  uint32_t* d_cnt = (uint32_t*)v->counters.p;
  unsigned long long* d_stats = (unsigned long long*)(d_cnt + CNT_WORDS);
  uint32_t* d_emit1 = d_cnt + 64;     // [0] polygons [1] items [2] capacity flag
  uint32_t* d_emit2 = d_cnt + 68;
  uint32_t* list0 = (uint32_t*)v->lists.p;
  uint32_t* list1 = list0 + n;
  uint32_t* list2 = list1 + n;
  uint32_t* list3 = list2 + n;
  uint32_t* sorted = list3 + n;
  uint32_t* listw = sorted + n;
  uint32_t* done2 = listw + n;
  cudaStream_t s1 = v->stream, s2 = v->stream2;
  int launches = 0;
  for (int attempt = 0; attempt < 8; attempt++) {
    BatchArgs a;
    a.vol = view_of(v);
    a.pos = d_pos; a.nrm = d_nrm; a.wx = w[0]; a.wy = w[1]; a.wz = w[2];
    a.nq = (uint32_t)n; a.counters = d_cnt; a.stats = d_stats;
    a.late = list0;
    a.done_list = nullptr;
    {
      P2Pools& p = a.p2;
      p.cursors = d_cnt + 72;
      p.order = (uint32_t*)v->p2_order.p; p.mask = (uint32_t*)v->p2_mask.p; p.order_cap = (uint32_t)v->p2_order_cap;
      p.hash = (uint32_t*)v->p2_hash.p;
      p.hash_cap = (uint32_t)p2_hash_words(v);
      p.chunk_cap = (uint32_t)(v->p2_order_cap / P2_CHUNK + v->p2_rec_cap + 32);
      uint32_t* c5 = (uint32_t*)v->p2_chunks.p;
      const size_t cs = (size_t)p.chunk_cap + 1;
      p.chunk_rec = c5; p.chunk_items = c5 + cs; p.chunk_polys = c5 + 2 * cs; p.item_off = c5 + 3 * cs; p.poly_off = c5 + 4 * cs;
      p.recs = (P2Rec*)v->p2_recs.p; p.rec_cap = (uint32_t)v->p2_rec_cap;
      p.hash_factor = v->p2_factor;
    }
    // the second half of the CTA tiers (claims, ownership, records) for everything parked since chunk `c0`
    auto launch_p2 = [&](cudaStream_t st) -> int {
      const int pgrid = v->sm_count * 8;
      p2_claims_kernel<<<pgrid, P2_CHUNK, 0, st>>>(a);
      p2_emit1_kernel<<<pgrid, P2_CHUNK, 0, st>>>(a);
      p2_scan_kernel<<<1, 1024, 0, st>>>(a);
      p2_emit2_kernel<<<pgrid, P2_CHUNK, 0, st>>>(a);
      CK(cudaGetLastError());
      launches += 4;
      return XS3D_OK;
    };
    EmitOut out1, out2;
    out1.polys = (PolyRec*)v->polys.p; out1.items = (ItemRec*)v->items.p; out1.geom = (PlaneGeom*)v->geom.p; out1.qinfo = (QInfo*)v->qinfo.p;
    out1.counters = d_emit1; out1.ready = QS_READY;
    out1.poly_cap = (uint32_t)std::min<uint64_t>(v->poly_cap, 0xffffffffull); out1.item_cap = (uint32_t)std::min<uint64_t>(v->item_cap, 0xffffffffull);
    out2 = out1;
    out2.polys = (PolyRec*)v->polys2.p; out2.items = (ItemRec*)v->items2.p; out2.counters = d_emit2; out2.ready = QS_READY2;
    out2.poly_cap = (uint32_t)std::min<uint64_t>(v->poly_cap2, 0xffffffffull); out2.item_cap = (uint32_t)std::min<uint64_t>(v->item_cap2, 0xffffffffull);
    float* vals1 = (float*)v->vals.p;
    float* vals2 = (float*)v->vals2.p;
    const unsigned cgrid_combine = (unsigned)std::min<uint64_t>(n, 65535ull * 16);
    // ---- everything from here to the read-back of the counters is one fixed launch sequence for given buffers and n.
    //      XS3D_GRAPH=1 captures it into a CUDA graph the first time and replays it afterwards (one driver call instead
    //      of ~50 on the host; timing events become external event nodes so that xs3d_volume_timing keeps working).
    //      OFF by default: measured on B200 (profiles/r2j_cuda_graph_replay.txt) the replayed sweep takes 1.05 ms against
    //      0.74 ms launched directly — the graph executor does not reproduce the co-residency of the three streams (the
    //      mid tier runs alone first: 0.21 ms, then the warp tier: 0.49 ms), which is what the launch order and the stream
    //      priorities of the direct path arrange.
    const bool speculate_wide = !single && v->expect_wide;
    bool capturing = false;
    auto rec_t = [&](cudaEvent_t ev, cudaStream_t st) -> cudaError_t {
      return capturing ? cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal) : cudaEventRecord(ev, st);
    };
    auto enqueue = [&]() -> int {
    CK(cudaMemsetAsync(d_cnt, 0, CNT_TOTAL_WORDS * 4, s1));
    CK(rec_t(v->evb[0], s1));
    if (!single) {
      // pre-pass: size estimate per query -> list ordered largest first; its head goes straight to the mid tier
      const int cgrid = (int)std::min<uint64_t>((n + 7) / 8, (uint64_t)v->sm_count * 8);
      classify_kernel<<<cgrid, 256, 0, s1>>>(a.vol, d_pos, d_nrm, w[0], w[1], w[2], (uint32_t)n, (uint8_t*)v->keys.p, d_cnt + CNT_HIST, list0, out1.qinfo, d_cnt + CNT_BADN);
      CK(cudaGetLastError());
      const int ogrid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)v->sm_count);
      order_kernel<<<ogrid, 256, 0, s1>>>((const uint8_t*)v->keys.p, d_cnt + CNT_HIST, d_cnt + CNT_CURSOR, sorted, (uint32_t)n, d_cnt + CNT_NBIG, v->mid_key ? v->mid_key : MID_BIG_KEY);
      CK(cudaGetLastError());
      launches += 2;
      CK(rec_t(v->evb[1], s1));
      // ---- residency plan per SM (228 KB of shared memory): two warp-tier CTAs (2 x 74.4 KB) + one mid-tier CTA
      //      (78.5 KB).  The warp tier is therefore launched as TWO grids pulling from the same queue: A (2 CTAs per SM)
      //      on the launching stream and B (1 CTA per SM) on a third stream.  The mid tier's stream has the higher
      //      priority, so its CTAs are placed before B's; B's CTAs move in as mid-tier CTAs retire.  Whichever of A and
      //      the mid tier is dispatched first, both fit.
      // ---- second stream: mid tier beside the warp tier (it never waits for it)
      CK(cudaEventRecord(v->evf[0], s1));
      CK(cudaStreamWaitEvent(s2, v->evf[0], 0));
      a.out = out2; a.done_list = done2; a.list_in = sorted; a.list_out = list1;
      CK(rec_t(v->evs[0], s2));
      area_mid_kernel<MidTier><<<v->mid_ctas, MidTier::THREADS, MidTier::SMEM_BYTES, s2>>>(a, v->cfg_mid, CNT_NEXT1, CNT_OVER1, 10, -1);
      CK(cudaGetLastError());
      CK(rec_t(v->evs[1], s2));
      CK(cudaEventRecord(v->evf[4], s2));
      CK(rec_t(v->evs[5], s2));
      // (dev knob XS3D_P2_EARLY=1: a P2 chain right here, for the sections the first launch parked, beside the end of the warp
      //  tier instead of behind the late hand-overs.  Measured on B200: no gain — the chain takes issue slots from the warp
      //  tier's last queries, 0.435 instead of 0.418 ms, and the step stays at 0.69 ms: profiles/r2t_negative_results.txt)
      static const bool p2_early = getenv("XS3D_P2_EARLY") != nullptr;
      if (p2_early) {
        a.out = out2; a.done_list = done2;
        RET(launch_p2(s2));
      }
      // profiling aid: under ncu every kernel runs alone, so the split of the warp tier into grids A and B would show
      // grid B (10 warps per SM) doing all the work; XS3D_SINGLE_GRID=1 launches the warp tier as one grid of 3 CTAs per SM
      const bool single_grid = getenv("XS3D_SINGLE_GRID") != nullptr;
      // ---- third stream: warp-tier grid B, the third warp-tier CTA per SM.  It starts when the mid tier's first launch has
      //      ended: launched beside it, its CTAs sometimes took the third slot of every SM before the mid tier's did, and
      //      the mid tier then ran after the warp tier instead of beside it (0.77 instead of 0.69 ms, for a whole process)
      CK(cudaStreamWaitEvent(v->stream3, v->evf[0], 0));
      CK(cudaStreamWaitEvent(v->stream3, v->evf[4], 0));
      a.out = out1; a.done_list = nullptr; a.list_in = sorted; a.list_out = list0;
      const int gridB = (int)std::min<uint64_t>((n + T0_WARPS - 1) / T0_WARPS, (uint64_t)v->sm_count * (T0_CTAS_PER_SM - 2));
      static const bool no_grid_b = getenv("XS3D_NO_GRIDB") != nullptr;
      if (!single_grid && !no_grid_b) {
        T0_KERNEL<<<gridB, T0_WARPS * 32, T0_SMEM_BYTES, v->stream3>>>(a);
        CK(cudaGetLastError());
      }
      CK(cudaEventRecord(v->evf[3], v->stream3));
      // ---- launching stream: warp-tier grid A
      const int gridA = (int)std::min<uint64_t>((n + T0_WARPS - 1) / T0_WARPS, (uint64_t)v->sm_count * (single_grid ? T0_CTAS_PER_SM : 2));
      T0_KERNEL<<<gridA, T0_WARPS * 32, T0_SMEM_BYTES, s1>>>(a);
      CK(cudaGetLastError());
      CK(cudaEventRecord(v->evf[1], s1));
      CK(cudaStreamWaitEvent(s1, v->evf[3], 0));       // the polygon kernel needs grid B's records too
      CK(rec_t(v->evb[2], s1));
      launches += 3;
      // ---- second stream: late hand-overs (after the warp tier), big tier, its own polygon + combine kernels
      CK(cudaStreamWaitEvent(s2, v->evf[1], 0));
      CK(cudaStreamWaitEvent(s2, v->evf[3], 0));
      a.out = out2; a.done_list = done2; a.list_in = sorted; a.list_out = list1;
      area_mid_kernel<MidTier><<<v->mid_ctas, MidTier::THREADS, MidTier::SMEM_BYTES, s2>>>(a, v->cfg_mid, CNT_NEXT1, CNT_OVER1, 10, -1);
      CK(cudaGetLastError());
      CK(rec_t(v->evs[2], s2));
      // Sections that leave the mid tier's window are rare (none in the 512^3 bench, four in the 1024^3 one): normally
      // they get a second round after the host has seen the counters (below).  A sweep that had some is likely to be
      // followed by another that has some (same object, same scale), so after such a batch the wide and big tiers are
      // launched here, behind the mid tier's second launch, and their sections join the same P2 chain: no host round
      // trip, no second chain of launches (1024^3 sweep: 0.33 ms of exposed tail).  Empty lists cost two launches.
      if (speculate_wide) {
        a.out = out2; a.done_list = done2; a.list_in = list1; a.list_out = listw;
        area_wide_kernel<<<v->wide_ctas, WideTier::THREADS, WIDE_SMEM_BYTES, s2>>>(a, v->cfg_wide, CNT_NEXTW, CNT_OVERW, d_cnt + CNT_OVER1, 10);
        CK(cudaGetLastError());
        a.list_in = listw; a.list_out = list2;
        area_block_kernel<BigTier><<<std::min(v->sm_count, 16), BigTier::THREADS, block_smem_bytes(v), s2>>>(a, v->cfg1, block_smem_bits_words(v), CNT_NEXT2, CNT_OVER2, d_cnt + CNT_OVERW, 10);
        CK(cudaGetLastError());
        launches += 2;
      }
      // the second half of every section the CTA tiers parked
      a.out = out2; a.done_list = done2;
      RET(launch_p2(s2));
      CK(rec_t(v->evs[3], s2));
      poly_eval_kernel<<<v->sm_count * 4, 256, 0, s2>>>(out2.polys, out2.geom, vals2, d_emit2, 0u, out2.poly_cap);
      CK(cudaGetLastError());
      combine_kernel<<<(unsigned)std::min<uint64_t>(n, 1024), 128, 0, s2>>>(out2.items, vals2, out2.qinfo, done2, d_cnt + CNT_DONE2, (uint32_t)n, d_area, d_contact, QS_READY2);
      CK(cudaGetLastError());
      CK(rec_t(v->evs[4], s2));
      CK(cudaEventRecord(v->evf[2], s2));
      launches += 3;
    } else {
      CK(rec_t(v->evb[1], s1));
      // tiny batches (single calls): mid tier first — a section of a few thousand cells takes a fraction of the big
      // tier's time there — then the big tier for what does not fit its window
      a.out = out1; a.done_list = nullptr; a.list_in = nullptr; a.list_out = list1;
      area_mid_kernel<MidTier><<<(int)n, MidTier::THREADS, MidTier::SMEM_BYTES, s1>>>(a, v->cfg_mid, CNT_NEXT1, CNT_OVER1, 10, (int)n);
      CK(cudaGetLastError());
      a.out = out1; a.done_list = nullptr;
      RET(launch_p2(s1));
      CK(rec_t(v->evb[2], s1));
      launches += 1;
    }
    poly_eval_kernel<<<v->sm_count * 8, 256, 0, s1>>>(out1.polys, out1.geom, vals1, d_emit1, 0u, out1.poly_cap);
    CK(cudaGetLastError());
    CK(rec_t(v->evb[3], s1));
    combine_kernel<<<cgrid_combine, 128, 0, s1>>>(out1.items, vals1, out1.qinfo, nullptr, nullptr, (uint32_t)n, d_area, d_contact, QS_READY);
    CK(cudaGetLastError());
    CK(rec_t(v->evb[4], s1));
    launches += 2;
    if (!single) CK(cudaStreamWaitEvent(s1, v->evf[2], 0));   // join
    CK(rec_t(v->evb[5], s1));
    CK(cudaMemcpyAsync(v->h_counters, d_cnt, 512, cudaMemcpyDeviceToHost, s1));
    if (h_area) {            // host callers: the results come down behind the kernels, under the same (single) host wait
      CK(cudaMemcpyAsync(h_area, d_area, n * 4, cudaMemcpyDeviceToHost, s1));
      CK(cudaMemcpyAsync(h_contact, d_contact, n, cudaMemcpyDeviceToHost, s1));
    }
    return XS3D_OK;
    };   // enqueue
    static const bool no_graph = getenv("XS3D_GRAPH") == nullptr;
    // (a pageable result buffer makes cudaMemcpyAsync synchronous, which a capture forbids: only pinned / device results)
    bool use_graph = !no_graph && s1 != nullptr && s1 != cudaStreamLegacy && s1 != cudaStreamPerThread;   // (the default streams cannot be captured)
    if (use_graph && h_area) {
      cudaPointerAttributes pa;
      if (cudaPointerGetAttributes(&pa, h_area) != cudaSuccess || pa.type != cudaMemoryTypeHost) use_graph = false;
      if (use_graph && (cudaPointerGetAttributes(&pa, h_contact) != cudaSuccess || pa.type != cudaMemoryTypeHost)) use_graph = false;
      cudaGetLastError();
    }
    if (!use_graph) {
      launches = 0;
      RET(enqueue());
    } else {
      xs3d_volume::BatchGraph key;
      key.gen = g_alloc_gen; key.n = n;
      const void* kp[7] = { d_pos, d_nrm, d_area, d_contact, h_area, h_contact, v->cubes.p };
      for (int i = 0; i < 7; i++) key.ptr[i] = kp[i];
      key.w[0] = w[0]; key.w[1] = w[1]; key.w[2] = w[2];
      key.single = (single ? 1 : 0) | (speculate_wide ? 2 : 0); key.mid_key = v->mid_key; key.factor = v->p2_factor;
      key.caps[0] = v->poly_cap; key.caps[1] = v->poly_cap2; key.caps[2] = v->p2_order_cap; key.caps[3] = v->p2_rec_cap; key.caps[4] = (uint64_t)(uintptr_t)s1;
      xs3d_volume::BatchGraph* hit = nullptr;
      for (auto& g : v->graphs)
        if (g.exec && g.gen == key.gen && g.n == key.n && !memcmp(g.ptr, key.ptr, sizeof(key.ptr)) && !memcmp(g.w, key.w, 12) && g.single == key.single &&
            g.mid_key == key.mid_key && g.factor == key.factor && !memcmp(g.caps, key.caps, sizeof(key.caps))) { hit = &g; break; }
      if (!hit) {
        launches = 0;
        CK(cudaStreamBeginCapture(s1, cudaStreamCaptureModeThreadLocal));
        capturing = true;
        const int rc = enqueue();
        capturing = false;
        cudaGraph_t graph = nullptr;
        const cudaError_t ee = cudaStreamEndCapture(s1, &graph);
        if (rc != XS3D_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        CK(ee);
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CK(ei);
        // keep at most 8 sequences per volume: drop the least recently used one (and every stale one)
        for (auto& g : v->graphs)
          if (g.exec && g.gen != key.gen) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        xs3d_volume::BatchGraph* slot = nullptr;
        for (auto& g : v->graphs) if (!g.exec) { slot = &g; break; }
        if (!slot && v->graphs.size() < 8) { v->graphs.emplace_back(); slot = &v->graphs.back(); }
        if (!slot) {
          slot = &v->graphs[0];
          for (auto& g : v->graphs) if (g.used < slot->used) slot = &g;
          cudaGraphExecDestroy(slot->exec);
        }
        key.exec = exec; key.launches = launches;
        *slot = key;
        hit = slot;
      }
      hit->used = ++v->graph_clock;
      launches = hit->launches;
      CK(cudaGraphLaunch(hit->exec, s1));
    }
    trace(v, "launched");
    if (v->after_enqueue && !use_graph) {
      // the caller enqueues follow-up work on the launching stream now (e.g. the uploads — or, in bench.py, the L2-evicting
      // write — of its NEXT step): it runs on the GPU right behind this batch, while the host is still on its way back with
      // this batch's results.  The host therefore waits for an event behind the result copies, not for the whole stream.
      if (!v->ev_done) CK(cudaEventCreateWithFlags(&v->ev_done, cudaEventDisableTiming));
      CK(cudaEventRecord(v->ev_done, s1));
      v->after_enqueue(v->after_enqueue_ctx);
      CK(cudaEventSynchronize(v->ev_done));
    } else {
      CK(host_wait(v, s1));
    }
    trace(v, "synced");
    v->timing_pending = single ? 2 : 1;   // event durations are read on demand (xs3d_volume_timing): ten driver calls off the hot path
    v->timing[3] = 0.0f;
    if (v->h_counters[CNT_BADN]) return fail(XS3D_ERR_ARG, "normal vector must not be a null vector (all zeros).");
    const uint32_t over0 = v->h_counters[CNT_OVER0], over1 = v->h_counters[CNT_OVER1];
    const uint32_t* he1 = v->h_counters + 64;
    const uint32_t* he2 = v->h_counters + 68;
    uint32_t over3 = 0;
    const uint32_t* hp2 = v->h_counters + 72;
    bool p2_full = hp2[P2C_FULL] != 0, p2_hashfull = hp2[P2C_HASHFULL] != 0;
    bool capacity = he1[2] != 0 || he2[2] != 0 || p2_full || p2_hashfull;
    const EmitOut& ob = single ? out1 : out2;            // where the CTA tiers' records go
    float* valsb = single ? vals1 : vals2;
    uint32_t* d_emitb = single ? d_emit1 : d_emit2;
    v->timing[2] = 0.0f;
    if (!single) v->expect_wide = over1 > 0;
    if (!capacity && over1 > 0 && !speculate_wide) {
      // SECOND ROUND: sections that left the mid tier's 256-cell window (or exceeded its 16 Ki samples) — the wide tier
      // (same walk, 416-cell window), the big tier's whole-plane bitmap walk for what leaves that too, their second half,
      // polygons and sums.  Rare in sweeps of thin objects (none in the 512^3 bench, four in the 1024^3 one), so the common
      // path does not pay for five launches that find nothing; a large single section pays one extra host round trip.
      const uint32_t used = single ? he1[0] : he2[0];
      a.out = ob; a.done_list = nullptr;
      const int wg = single ? (int)std::min<uint64_t>(over1, (uint64_t)v->wide_ctas) : v->wide_ctas;
      CK(cudaEventRecord(v->ev[2], s1));
      a.list_in = list1; a.list_out = listw;
      area_wide_kernel<<<wg, WideTier::THREADS, WIDE_SMEM_BYTES, s1>>>(a, v->cfg_wide, CNT_NEXTW, CNT_OVERW, d_cnt + CNT_OVER1, 10);
      CK(cudaGetLastError());
      a.list_in = listw; a.list_out = list2;
      area_block_kernel<BigTier><<<(int)std::min<uint64_t>(over1, (uint64_t)v->sm_count), BigTier::THREADS, block_smem_bytes(v), s1>>>(a, v->cfg1, block_smem_bits_words(v), CNT_NEXT2, CNT_OVER2, d_cnt + CNT_OVERW, 10);
      CK(cudaGetLastError());
      RET(launch_p2(s1));
      poly_eval_kernel<<<v->sm_count * 8, 256, 0, s1>>>(ob.polys, ob.geom, valsb, d_emitb, used, ob.poly_cap);
      combine_kernel<<<(unsigned)std::min<uint64_t>(over1, 1024), 128, 0, s1>>>(ob.items, valsb, ob.qinfo, list1, d_cnt + CNT_OVER1, (uint32_t)n, d_area, d_contact, ob.ready);
      CK(cudaGetLastError());
      CK(cudaEventRecord(v->ev[3], s1));
      launches += 4;
      CK(cudaMemcpyAsync(v->h_counters, d_cnt, 512, cudaMemcpyDeviceToHost, s1));
      if (h_area) {
        CK(cudaMemcpyAsync(h_area, d_area, n * 4, cudaMemcpyDeviceToHost, s1));
        CK(cudaMemcpyAsync(h_contact, d_contact, n, cudaMemcpyDeviceToHost, s1));
      }
      CK(host_wait(v, s1));
      CK(cudaEventElapsedTime(&v->timing[2], v->ev[2], v->ev[3]));
      p2_full = hp2[P2C_FULL] != 0; p2_hashfull = hp2[P2C_HASHFULL] != 0;
      capacity = he1[2] != 0 || he2[2] != 0 || p2_full || p2_hashfull;
    }
    const uint32_t over2 = v->h_counters[CNT_OVER2];
    v->wide_over = v->h_counters[CNT_OVERW];
    if (!capacity && over2 > 0) {
      // worst-case slab, one CTA, only for the queries the big tier could not fit (records go where the big tier's went)
      RET(ensure_worst_case_slab(v));
      const uint32_t used = single ? he1[0] : he2[0];
      a.out = ob; a.done_list = nullptr; a.list_in = list2; a.list_out = list3;
      CK(cudaEventRecord(v->ev[0], s1));
      area_block_kernel<BigTier><<<1, BigTier::THREADS, block_smem_bytes(v), s1>>>(a, v->cfg2, block_smem_bits_words(v), CNT_NEXT3, CNT_OVER3, d_cnt + CNT_OVER2, 10);
      CK(cudaGetLastError());
      RET(launch_p2(s1));
      CK(cudaEventRecord(v->ev[1], s1));
      poly_eval_kernel<<<v->sm_count * 8, 256, 0, s1>>>(ob.polys, ob.geom, valsb, d_emitb, used, ob.poly_cap);
      combine_kernel<<<1024, 128, 0, s1>>>(ob.items, valsb, ob.qinfo, list2, d_cnt + CNT_OVER2, (uint32_t)n, d_area, d_contact, ob.ready);
      CK(cudaGetLastError());
      launches += 3;
      CK(cudaMemcpyAsync(v->h_counters, d_cnt, 512, cudaMemcpyDeviceToHost, s1));
      if (h_area) {
        CK(cudaMemcpyAsync(h_area, d_area, n * 4, cudaMemcpyDeviceToHost, s1));
        CK(cudaMemcpyAsync(h_contact, d_contact, n, cudaMemcpyDeviceToHost, s1));
      }
      CK(cudaStreamSynchronize(s1));
      CK(cudaEventElapsedTime(&v->timing[3], v->ev[0], v->ev[1]));
      over3 = v->h_counters[CNT_OVER3];
      p2_full = hp2[P2C_FULL] != 0; p2_hashfull = hp2[P2C_HASHFULL] != 0;
      capacity = he1[2] != 0 || he2[2] != 0 || p2_full || p2_hashfull;
    }
    const unsigned long long* hs = (const unsigned long long*)(v->h_counters + CNT_WORDS);
    v->stats[0] = hs[0]; v->stats[1] = hs[1]; v->stats[2] = hs[2] * 1024ull;
    v->stats[3] = (single ? 0u : v->h_counters[CNT_NBIG]) + over0; v->stats[4] = over1; v->stats[5] = (uint64_t)launches; v->stats[6] = over2;
    v->stats[7] = (uint64_t)he1[0] + (uint64_t)he2[0];
    v->stats[8] = over3; v->stats[9] = 0;
    v->tier_stats[0] = hs[16]; v->tier_stats[1] = hs[17];
    for (int i = 0; i < 12; i++) v->prof[i] = hs[4 + i];
    if (capacity) {
      // a record buffer filled up: double it and run the batch again
      if (he1[2] != 0) { v->poly_cap *= 2; v->item_cap *= 2; }
      if (he2[2] != 0) { v->poly_cap2 *= 2; v->item_cap2 *= 2; }
      if (p2_full) {
        // the cursors kept counting past the capacities: they say how much this batch needs
        if (v->p2_order_cap >= 0x7fffffffull) return fail(XS3D_ERR_CAPACITY, "sample pool would exceed 2^31 entries; split the batch");
        const uint64_t need_order = (uint64_t)hp2[P2C_ORDER] + (uint64_t)hp2[P2C_ORDER] / 4 + 1024, need_hash = (uint64_t)hp2[P2C_HASH] + (uint64_t)hp2[P2C_HASH] / 4 + 4096;
        v->p2_order_cap = std::min<uint64_t>(std::max<uint64_t>(2 * v->p2_order_cap, need_order), 0x7fffffffull);
        v->p2_hash_min_words = std::max<uint64_t>(v->p2_hash_min_words, need_hash);
      }
      if (p2_hashfull) {
        if (v->p2_factor >= 512) return fail(XS3D_ERR_CAPACITY, "a first-touch table filled up at 512 slots per sample");
        v->p2_factor *= 4;
      }
      if (p2_full || p2_hashfull) RET(ensure_batch_workspace(v, n));
      if (v->poly_cap > (1ull << 31) || v->poly_cap2 > (1ull << 31)) return fail(XS3D_ERR_CAPACITY, "polygon record buffer would exceed 2^31 entries; split the batch");
      RET(v->polys.ensure(v->poly_cap * sizeof(PolyRec) + 64));
      RET(v->vals.ensure(v->poly_cap * 4 + 64));
      RET(v->items.ensure(v->item_cap * sizeof(ItemRec) + 64));
      RET(v->polys2.ensure(v->poly_cap2 * sizeof(PolyRec) + 64));
      RET(v->vals2.ensure(v->poly_cap2 * 4 + 64));
      RET(v->items2.ensure(v->item_cap2 * sizeof(ItemRec) + 64));
      continue;
    }
    if (over3 > 0) return fail(XS3D_ERR_CAPACITY, "a query exceeded the worst-case working set");
    return XS3D_OK;
  }
  return fail(XS3D_ERR_CAPACITY, "record buffers kept overflowing");
}

extern "C" int xs3d_area_batch(xs3d_volume* v, uint64_t n, const float* pos, const float* normal, const float anisotropy[3],
                               int mem, float* area, uint8_t* contact) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (!v->loaded) return fail(XS3D_ERR_ARG, "volume has no image loaded");
  if (n == 0) return XS3D_OK;
  if (n >= 0x7fffffffull) return fail(XS3D_ERR_ARG, "too many queries in one batch");
  if (!pos || !normal || !anisotropy || !area || !contact) return fail(XS3D_ERR_ARG, "null argument");
  CK(cudaSetDevice(v->device));
  if ((size_t)v->sx * v->sy * v->sz == 0) {
    if (mem == XS3D_HOST) { memset(area, 0, n * 4); memset(contact, 0, n); }
    else { CK(cudaMemsetAsync(area, 0, n * 4, v->stream)); CK(cudaMemsetAsync(contact, 0, n, v->stream)); CK(cudaStreamSynchronize(v->stream)); }
    return XS3D_OK;
  }
  RET(ensure_batch_workspace(v, n));
  // an all-zero normal is an argument error (the reference raises ValueError, xs3d/__init__.py:72-75): host buffers and
  // tiny device batches are checked here, large device batches by the pre-pass kernel
  if (mem == XS3D_HOST || n < 4) {
    float tmp[9];
    const float* hn = normal;
    if (mem != XS3D_HOST) { CK(cudaMemcpyAsync(tmp, normal, n * 12, cudaMemcpyDeviceToHost, v->stream)); CK(cudaStreamSynchronize(v->stream)); hn = tmp; }
    for (uint64_t i = 0; i < n; i++)
      if (hn[3 * i] == 0.0f && hn[3 * i + 1] == 0.0f && hn[3 * i + 2] == 0.0f) return fail(XS3D_ERR_ARG, "normal vector must not be a null vector (all zeros).");
  }
  const float* d_pos = pos; const float* d_nrm = normal;
  float* d_area = area; uint8_t* d_contact = contact;
  if (mem == XS3D_HOST) {
    CK(cudaMemcpyAsync(v->q_pos.p, pos, n * 12, cudaMemcpyHostToDevice, v->stream));
    CK(cudaMemcpyAsync(v->q_nrm.p, normal, n * 12, cudaMemcpyHostToDevice, v->stream));
    d_pos = (const float*)v->q_pos.p; d_nrm = (const float*)v->q_nrm.p;
    d_area = (float*)v->q_area.p; d_contact = (uint8_t*)v->q_contact.p;
  }
  RET(area_batch_device(v, n, d_pos, d_nrm, anisotropy, d_area, d_contact, n < 4, mem == XS3D_HOST ? area : nullptr, mem == XS3D_HOST ? contact : nullptr));
  return XS3D_OK;
}

// One sweep with everything on the host: queries are enqueued BEFORE the (much larger) image so that, when several
// sweeps are in flight on different volumes/streams, a sweep's query upload never queues behind another sweep's
// 128 MiB image in the copy engine; then ingest, the query kernels and the read-back — one call, one final sync.
extern "C" int xs3d_area_sweep_host(xs3d_volume* v, const uint8_t* binimg, int c_order, uint64_t n, const float* pos, const float* normal,
                                    const float anisotropy[3], float* area, uint8_t* contact) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (!binimg || !pos || !normal || !anisotropy || !area || !contact) return fail(XS3D_ERR_ARG, "null argument");
  if (n >= 0x7fffffffull) return fail(XS3D_ERR_ARG, "too many queries in one batch");
  CK(cudaSetDevice(v->device));
  const size_t voxels = (size_t)v->sx * v->sy * v->sz;
  if (voxels == 0 || n == 0) {
    memset(area, 0, n * 4); memset(contact, 0, n);
    v->loaded = true;
    return XS3D_OK;
  }
  for (uint64_t i = 0; i < n; i++)
    if (normal[3 * i] == 0.0f && normal[3 * i + 1] == 0.0f && normal[3 * i + 2] == 0.0f) return fail(XS3D_ERR_ARG, "normal vector must not be a null vector (all zeros).");
  RET(ensure_batch_workspace(v, n));
  CK(cudaMemcpyAsync(v->q_pos.p, pos, n * 12, cudaMemcpyHostToDevice, v->stream));
  CK(cudaMemcpyAsync(v->q_nrm.p, normal, n * 12, cudaMemcpyHostToDevice, v->stream));
  if (use_packed_upload(voxels)) {
    RET(upload_packed_and_ingest(v, binimg, c_order));
  } else {
    RET(v->raw.ensure(voxels + 64));
    CK(cudaMemcpyAsync(v->raw.p, binimg, voxels, cudaMemcpyHostToDevice, v->stream));
    CK(cudaEventRecord(v->ev[6], v->stream));
    RET(launch_ingest(v, (const uint8_t*)v->raw.p, c_order));
    CK(cudaEventRecord(v->ev[7], v->stream));
  }
  v->loaded = true;
  RET(area_batch_device(v, n, (const float*)v->q_pos.p, (const float*)v->q_nrm.p, anisotropy, (float*)v->q_area.p, (uint8_t*)v->q_contact.p, n < 4, area, contact));
  CK(cudaEventElapsedTime(&v->timing[4], v->ev[6], v->ev[7]));
  trace(v, "results");
  return XS3D_OK;
}

extern "C" int xs3d_volume_stats(xs3d_volume* v, uint64_t stats[10]) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  for (int i = 0; i < 10; i++) stats[i] = v->stats[i];
  stats[9] = v->tier_stats[0];   // samples handled by the CTA tiers (of stats[0])
  return XS3D_OK;
}

extern "C" int xs3d_volume_timing(xs3d_volume* v, float ms[12]) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (v->timing_pending) {
    const bool single = v->timing_pending == 2;
    v->timing_pending = 0;
    CK(cudaSetDevice(v->device));
    const float ingest = v->timing[4], worst = v->timing[3], second = v->timing[2];
    for (int i = 0; i < 12; i++) v->timing[i] = 0.0f;
    v->timing[4] = ingest; v->timing[3] = worst; v->timing[2] = second;
    CK(cudaEventElapsedTime(&v->timing[9], v->evb[0], v->evb[1]));                 // pre-pass
    CK(cudaEventElapsedTime(&v->timing[single ? 1 : 0], v->evb[1], v->evb[2]));   // warp tier (tiny batches: mid tier + its second half)
    CK(cudaEventElapsedTime(&v->timing[5], v->evb[2], v->evb[3]));                 // polygon kernel
    CK(cudaEventElapsedTime(&v->timing[6], v->evb[3], v->evb[4]));                 // combine kernel
    CK(cudaEventElapsedTime(&v->timing[8], v->evb[4], v->evb[5]));                 // waiting for the second stream
    if (!single) {
      CK(cudaEventElapsedTime(&v->timing[1], v->evs[0], v->evs[1]));               // mid tier beside the warp tier
      CK(cudaEventElapsedTime(&v->timing[10], v->evs[5], v->evs[2]));              // ... its second launch (incl. waiting for the warp tier)
      CK(cudaEventElapsedTime(&v->timing[7], v->evs[1], v->evs[5]));               // second half of the CTA tiers (P2 kernels), first chain
      {
        float t2 = 0.0f;
        CK(cudaEventElapsedTime(&t2, v->evs[2], v->evs[3]));                       // ... second chain (late hand-overs)
        v->timing[7] += t2;
      }
      CK(cudaEventElapsedTime(&v->timing[11], v->evs[3], v->evs[4]));              // second stream's polygon + combine kernels
    }
  }
  for (int i = 0; i < 12; i++) ms[i] = v->timing[i];
  return XS3D_OK;
}

// `fn(ctx)` is called by xs3d_area_batch / xs3d_area_sweep_host on the calling thread once a batch — kernels and result
// copies — is enqueued, before the host waits for it: work the callback enqueues on the volume's stream runs right behind
// the batch, and does not delay the batch's results (the host waits for an event recorded before the callback).  null: off.
extern "C" int xs3d_volume_set_after_enqueue(xs3d_volume* v, void (*fn)(void*), void* ctx) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  v->after_enqueue = fn;
  v->after_enqueue_ctx = ctx;
  return XS3D_OK;
}

extern "C" int xs3d_volume_set_stream(xs3d_volume* v, void* cuda_stream) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  CK(cudaSetDevice(v->device));
  CK(cudaStreamSynchronize(v->stream));
  if (v->own_stream) CK(cudaStreamDestroy(v->stream));
  v->stream = (cudaStream_t)cuda_stream;
  v->own_stream = false;
  return XS3D_OK;
}

extern "C" int xs3d_volume_profile(xs3d_volume* v, uint64_t prof[12]) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  for (int i = 0; i < 12; i++) prof[i] = v->prof[i];
  return XS3D_OK;
}

// ---- path B ----------------------------------------------------------------------------------------
// xs_section.cu: grid-wide kernels (lattice bitmap -> union-find over row runs -> component -> voxels -> polygon kernel
// -> compaction); leaves the non-zero (index, value) pairs on the device.
int xs_section_fast_sparse(xs3d_volume* v, const float p[3], const float n[3], const float w[3],
                           uint64_t** d_idx, float** d_val, uint64_t* nnz, uint8_t* contact);
static int section_fast_sparse(xs3d_volume* v, const float p[3], const float n[3], const float w[3],
                               uint64_t** d_idx, float** d_val, uint64_t* nnz, uint8_t* contact) {
  RET(configure_kernels(v));
  return xs_section_fast_sparse(v, p, n, w, d_idx, d_val, nnz, contact);
}

int xs_section_slow_sparse(xs3d_volume* v, const float p[3], const float n[3], const float w[3], int method,
                           uint64_t** d_idx, float** d_val, uint64_t* nnz, uint8_t* contact);   // xs_slow.cu
int xs_area_slow(xs3d_volume* v, const float p[3], const float n[3], const float w[3], float* area, uint8_t* contact);   // xs_slow.cu

extern "C" int xs3d_section_sparse(xs3d_volume* v, const float pos[3], const float normal[3], const float anisotropy[3],
                                   int method, uint64_t* idx, float* val, uint64_t cap, uint64_t* nnz, uint8_t* contact) {
  if (!v) return fail(XS3D_ERR_ARG, "null volume");
  if (!v->loaded) return fail(XS3D_ERR_ARG, "volume has no image loaded");
  CK(cudaSetDevice(v->device));
  *nnz = 0; *contact = 0;
  if ((size_t)v->sx * v->sy * v->sz == 0) return XS3D_OK;
  uint64_t* d_idx = nullptr; float* d_val = nullptr;
  if (method == 0) RET(section_fast_sparse(v, pos, normal, anisotropy, &d_idx, &d_val, nnz, contact));
  else RET(xs_section_slow_sparse(v, pos, normal, anisotropy, method, &d_idx, &d_val, nnz, contact));
  // both paths leave exactly the non-zero voxels on the device
  const uint64_t k = *nnz;
  if (k > cap) return fail(XS3D_ERR_ARG, "sparse output buffers too small (nnz returned)");
  if (k) {
    CK(cudaMemcpyAsync(idx, d_idx, k * 8, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaMemcpyAsync(val, d_val, k * 4, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
  }
  return XS3D_OK;
}

// ---- process-wide default volume backing the drop-in single-call entry points ------------------------
static std::mutex g_mu;
static xs3d_volume* g_default = nullptr;   // last image seen by a single-call entry point

static int default_volume(uint64_t sx, uint64_t sy, uint64_t sz, xs3d_volume** out) {
  if (g_default && ((uint64_t)g_default->sx != sx || (uint64_t)g_default->sy != sy || (uint64_t)g_default->sz != sz)) {
    xs3d_volume_destroy(g_default);
    g_default = nullptr;
  }
  if (!g_default) {
    int dev = 0;
    cudaGetDevice(&dev);
    RET(xs3d_volume_create(dev, sx, sy, sz, &g_default));
  }
  *out = g_default;
  return XS3D_OK;
}

// fastxs3d.set_shape (xs3d.hpp:1295-1298 sizes the reusable visited buffer): here it creates the process-wide default
// volume of that shape with its batch workspaces, so that the first query does not pay for the allocations.
extern "C" int xs3d_set_shape(uint64_t sx, uint64_t sy, uint64_t sz) {
  std::lock_guard<std::mutex> lk(g_mu);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(XS3D_ERR_CUDA, "no CUDA device");
  if (sx * sy * sz == 0) return XS3D_OK;
  xs3d_volume* v;
  RET(default_volume(sx, sy, sz, &v));
  CK(cudaSetDevice(v->device));
  return ensure_batch_workspace(v, 1);
}
// set_shape + upload of the image (a warm start for callers that go on with the same image; later calls still upload
// the image they are given)
extern "C" int xs3d_set_shape_image(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz, int c_order) {
  RET(xs3d_set_shape(sx, sy, sz));
  std::lock_guard<std::mutex> lk(g_mu);
  if (sx * sy * sz == 0) return XS3D_OK;
  xs3d_volume* v;
  RET(default_volume(sx, sy, sz, &v));
  return xs3d_volume_load(v, binimg, XS3D_HOST, c_order);
}
extern "C" int xs3d_clear_shape(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_default) xs3d_volume_destroy(g_default);
  g_default = nullptr;
  return XS3D_OK;
}

extern "C" int xs3d_area(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz, const float pos[3], const float normal[3],
                         const float anisotropy[3], int slow_method, int use_persistent_data, float* area, uint8_t* contact) {
  std::lock_guard<std::mutex> lk(g_mu);
  *area = 0.0f; *contact = 0;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(XS3D_ERR_CUDA, "no CUDA device");
  if (sx * sy * sz == 0) return XS3D_OK;
  xs3d_volume* v;
  RET(default_volume(sx, sy, sz, &v));
  // The reference's persistent data is only a reusable visited buffer (xs3d.hpp:963-971): `binimg` is always the image
  // that is evaluated.  Here the reusable part is the default volume with its workspaces (created by xs3d_set_shape);
  // the image itself goes up on every call, whatever use_persistent_data says.
  (void)use_persistent_data;
  RET(xs3d_volume_load(v, binimg, XS3D_HOST, 0));
  if (slow_method) return xs_area_slow(v, pos, normal, anisotropy, area, contact);
  return xs3d_area_batch(v, 1, pos, normal, anisotropy, XS3D_HOST, area, contact);
}

extern "C" int xs3d_section(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz, const float pos[3], const float normal[3],
                            const float anisotropy[3], int method, float* out, uint8_t* contact) {
  std::lock_guard<std::mutex> lk(g_mu);
  *contact = 0;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(XS3D_ERR_CUDA, "no CUDA device");
  if (sx * sy * sz == 0) return XS3D_OK;
  xs3d_volume* v;
  RET(default_volume(sx, sy, sz, &v));
  RET(xs3d_volume_load(v, binimg, XS3D_HOST, 0));
  uint64_t* d_idx = nullptr; float* d_val = nullptr; uint64_t nnz = 0;
  if (method == 0) RET(section_fast_sparse(v, pos, normal, anisotropy, &d_idx, &d_val, &nnz, contact));
  else RET(xs_section_slow_sparse(v, pos, normal, anisotropy, method, &d_idx, &d_val, &nnz, contact));
  if (nnz) {
    std::vector<uint64_t> idx(nnz);
    std::vector<float> val(nnz);
    CK(cudaMemcpyAsync(idx.data(), d_idx, nnz * 8, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaMemcpyAsync(val.data(), d_val, nnz * 4, cudaMemcpyDeviceToHost, v->stream));
    CK(cudaStreamSynchronize(v->stream));
    for (uint64_t i = 0; i < nnz; i++) out[idx[i]] = val[i];     // scatter into the caller's zero-filled image (xs3d.hpp:526-528)
  }
  return XS3D_OK;
}

// accessors for the other translation units
xs::VolView xs_view(const xs3d_volume* v) { return view_of(v); }
cudaStream_t xs_stream(const xs3d_volume* v) { return v->stream; }
int xs_sm_count(const xs3d_volume* v) { return v->sm_count; }
int xs_fail(int code, const char* msg) { return fail(code, msg); }
int xs_buf(xs3d_volume* v, int which, size_t bytes, void** p) {
  DevBuf* b = which >= 4 ? &v->slice_bufs[which - 4] : (which == 0 ? &v->sec_idx : (which == 1 ? &v->sec_val : (which == 2 ? &v->misc : &v->counters)));
  int r = b->ensure(bytes);
  *p = b->p;
  return r;
}
uint32_t* xs_pinned(xs3d_volume* v) { return v->h_counters; }
int xs_default_volume(uint64_t sx, uint64_t sy, uint64_t sz, xs3d_volume** out) {
  return default_volume(sx, sy, sz, out);
}
std::mutex& xs_mutex() { return g_mu; }
int xs_set_device(const xs3d_volume* v) { CK(cudaSetDevice(v->device)); return XS3D_OK; }
int xs_visited_bits(xs3d_volume* v, uint32_t** p) {
  const size_t bytes = (((size_t)v->sx * v->sy * v->sz + 31) / 32) * 4 + 64;
  if (v->visited.cap < bytes) {
    RET(v->visited.ensure(bytes));
    CK(cudaMemsetAsync(v->visited.p, 0, v->visited.cap, v->stream));
  }
  *p = (uint32_t*)v->visited.p;
  return XS3D_OK;
}
int xs_launch_poly_eval(xs3d_volume* v, const PolyRec* polys, const PlaneGeom* geom, float* vals, const uint32_t* count_ptr, uint32_t cap) {
  poly_eval_kernel<<<v->sm_count * 8, 256, 0, v->stream>>>(polys, geom, vals, count_ptr, 0u, cap);
  CK(cudaGetLastError());
  return XS3D_OK;
}
void** xs_slice_state_slot(xs3d_volume* v, void (*free_fn)(void*)) { v->slice_state_free = free_fn; return &v->slice_state; }
void xs_labels(xs3d_volume* v, void** p, int* itemsize, int* c_order, bool* loaded) {
  *p = v->labels.p; *itemsize = v->itemsize; *c_order = v->labels_c_order; *loaded = v->labels_loaded;
}
int xs_labels_set(xs3d_volume* v, const void* labels, int itemsize, int mem, int c_order) {
  const size_t bytes = (size_t)v->sx * v->sy * v->sz * (size_t)itemsize;
  v->itemsize = itemsize; v->labels_c_order = c_order;
  if (bytes == 0) { v->labels_loaded = true; return XS3D_OK; }
  RET(v->labels.ensure(bytes + 64));
  if (labels) {            // (null: allocate only — the caller fills the buffer, xs3d_volume_labels_buffer)
    CK(cudaMemcpyAsync(v->labels.p, labels, bytes, mem == XS3D_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, v->stream));
    CK(cudaStreamSynchronize(v->stream));
  }
  v->labels_loaded = true;
  return XS3D_OK;
}
