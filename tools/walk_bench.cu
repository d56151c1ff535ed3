// Dev micro-benchmark: cycles per step of the neighbour-mask walk (xs_query.cuh, dfs_run_nbr) on a synthetic disc.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I cross-section_b200/csrc tools/walk_bench.cu -o gpurun_out/walk_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "xs_query.cuh"
using namespace xs;

constexpr int WLOG = 8, W = 256, FS = W / 32 + 2;
constexpr int O_LUT = 0, O_F = 128, O_RING = O_F + (W + 2) * FS, RING = 4096, O_CELLS = O_RING + RING / 2, WORDS = O_CELLS + W * W / 4 + 128;

__global__ void __launch_bounds__(256) walk_kernel(int radius, int variant, uint16_t* order, uint16_t* spill, long long* out) {
  extern __shared__ __align__(16) uint32_t smem[];
  BlockGroup g(reinterpret_cast<int*>(smem + WORDS));
  Arena<uint16_t> A;
  A.wlog = WLOG; A.fbits = smem + O_F; A.fstride = FS; A.nlut = reinterpret_cast<const int16_t*>(smem + O_LUT);
  A.cells = reinterpret_cast<uint8_t*>(smem + O_CELLS);
  A.stack = reinterpret_cast<uint16_t*>(smem + O_RING); A.stack_cap = RING; A.spill = spill + (size_t)blockIdx.x * 65536;
  A.order = order + (size_t)blockIdx.x * 65536; A.max_n = 60000;
  A.ox = A.oy = 0; A.tx = A.ty = 8;
  nbr_lut_fill(g, reinterpret_cast<int16_t*>(smem + O_LUT), WLOG);
  for (int i = threadIdx.x; i < (W + 2) * FS; i += blockDim.x) A.fbits[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < W * W; i += blockDim.x) {
    const int u = i & 255, v = i >> 8;
    const int dx = u - 128, dy = v - 128;
    if (dx * dx + dy * dy <= radius * radius) atomicOr(&frow(A, v)[u >> 5], 1u << (u & 31));
  }
  __syncthreads();
  nbr_masks(g, A, ~0ull);
  __syncthreads();
  if (threadIdx.x < 32) {
    const int ci0 = 128 + (128 << 8);
    if (threadIdx.x == 0) {
      A.order[0] = (uint16_t)ci0; A.stack[0] = (uint16_t)ci0;
      for (int k = 0; k < 8; k++) nbr_visit_one(A.cells, ci0, k, WLOG);
    }
    __syncwarp();
    int n = 1, depth = 1, lo = 0, ci = ci0;
    const long long t0 = clock64();
    const int st = dfs_run_nbr_warp(A, n, depth, lo, ci, (int)threadIdx.x, smem + WORDS);
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = n; out[blockIdx.x * 4 + 2] = st; }
  }
}

int main() {
  uint16_t *order, *spill; long long* out;
  const int ctas = 296;
  cudaMalloc(&order, (size_t)ctas * 65536 * 2); cudaMalloc(&spill, (size_t)ctas * 65536 * 2); cudaMalloc(&out, ctas * 32);
  const int smem_bytes = (WORDS + 64) * 4;
  cudaFuncSetAttribute(walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int grid : {1, 148, 296})
    for (int r : {8, 16, 32, 64, 120}) {
      walk_kernel<<<grid, 256, smem_bytes>>>(r, 0, order, spill, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4];
      cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
      printf("grid %3d radius %3d: n %6lld status %lld cycles %9lld -> %.1f cycles/sample (%s)\n", grid, r, h[1], h[2], h[0], (double)h[0] / (double)h[1], cudaGetErrorString(e));
    }
  return 0;
}
