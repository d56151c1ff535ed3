// xs_device.h — internal glue between the translation units of libxs3d_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "xs_query.cuh"

struct xs3d_volume;

xs::VolView xs_view(const xs3d_volume* v);
cudaStream_t xs_stream(const xs3d_volume* v);
int xs_sm_count(const xs3d_volume* v);
int xs_fail(int code, const char* msg);
// growable device buffers owned by the volume: 0 sparse idx, 1 sparse val, 2 misc, 3 counters, 4-7 batched slices
// (plan, row summaries, fill meta, host staging), 8 path B workspace, 9 skeleton tangents
int xs_buf(xs3d_volume* v, int which, size_t bytes, void** p);
uint32_t* xs_pinned(xs3d_volume* v);   // 256 bytes of pinned host memory
int xs_default_volume(uint64_t sx, uint64_t sy, uint64_t sz, xs3d_volume** out);
std::mutex& xs_mutex();
void xs_labels(xs3d_volume* v, void** p, int* itemsize, int* c_order, bool* loaded);
int xs_labels_set(xs3d_volume* v, const void* labels, int itemsize, int mem, int c_order);
int xs_set_device(const xs3d_volume* v);
// one bit per voxel, zeroed when (re)allocated; path B leaves it all-zero after every call
int xs_visited_bits(xs3d_volume* v, uint32_t** p);
// poly_eval_kernel (xs_device.cu) on the volume's stream: vals[i] = area of polys[i] for i < min(*count_ptr, cap)
int xs_launch_poly_eval(xs3d_volume* v, const xs::PolyRec* polys, const xs::PlaneGeom* geom, float* vals, const uint32_t* count_ptr, uint32_t cap);
// opaque per-volume state of the batched slice path (freed with `free_fn` when the volume is destroyed)
void** xs_slice_state_slot(xs3d_volume* v, void (*free_fn)(void*));

#define XS_CK(call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      char b__[512];                                                                                 \
      snprintf(b__, sizeof(b__), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return xs_fail(1, b__);                                                                        \
    }                                                                                                \
  } while (0)
#define XS_RET(call)          \
  do {                        \
    int r__ = (call);         \
    if (r__ != 0) return r__; \
  } while (0)
