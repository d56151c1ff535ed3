/* xs3d_b200.h — C ABI of libxs3d_b200.so, the B200-native replacement for the reference's pybind11
 * extension `fastxs3d` (/root/reference/src/fastxs3d.cpp:218-226).
 *
 * Plain C: pointers, sizes, status codes.  No torch / numpy / C++ types cross this boundary.
 * Every entry point returns XS3D_OK (0) or an error code; xs3d_last_error() returns the message of
 * the last failure on the calling thread.  There is NO CPU implementation behind these symbols:
 * without a CUDA device every compute entry point fails with XS3D_ERR_CUDA.
 *
 * Array conventions (same as the reference, xs3d/__init__.py:77-83, fastxs3d.cpp:20-26):
 *   binimg : uint8 (bool viewed as uint8), 3-D, Fortran order  -> index x + sx*(y + sy*z); nonzero = foreground
 *   pos / normal / anisotropy : float32[3] per query (batched: float32[n][3], row-major)
 *   area   : float32 (the reference's float32 result, bit-identical)
 *   contact: uint8 bitfield  bit0 -x, bit1 +x, bit2 -y, bit3 +y, bit4 -z, bit5 +z
 * `mem` arguments say where the caller's buffers live: XS3D_HOST (0) or XS3D_DEVICE (1, same device
 * as the volume; stream-ordered on the volume's stream and synchronised before returning).
 */
#ifndef XS3D_B200_H_
#define XS3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XS3D_OK 0
#define XS3D_ERR_CUDA 1        /* CUDA runtime failure / no device */
#define XS3D_ERR_ARG 2         /* invalid argument */
#define XS3D_ERR_UNSUPPORTED 3 /* e.g. volume with >= 2^32 voxel blocks */
#define XS3D_ERR_CAPACITY 4    /* internal working set exceeded even on the last-resort path */

#define XS3D_HOST 0
#define XS3D_DEVICE 1

typedef struct xs3d_volume xs3d_volume;   /* a volume resident on one GPU + its workspaces */

const char* xs3d_last_error(void);
int xs3d_version(void);
int xs3d_device_count(int* count);

/* ---- resident volumes (the GPU meaning of the reference's set_shape/clear_shape, xs3d.hpp:1295-1303,
 *      and the handle the batched entry points work on) ------------------------------------------------ */
int xs3d_volume_create(int device, uint64_t sx, uint64_t sy, uint64_t sz, xs3d_volume** out);
int xs3d_volume_destroy(xs3d_volume* vol);
/* Upload (if mem == XS3D_HOST) and ingest a binary image into the 2x2x2 occupancy layout.
 * c_order != 0: the buffer is C-contiguous [sx][sy][sz] (index z + sz*(y + sy*x)); this replaces the
 * host-side np.asfortranarray copy of xs3d/__init__.py:77. */
int xs3d_volume_load(xs3d_volume* vol, const uint8_t* binimg, int mem, int c_order);
/* The two halves of the bit-packed upload as separate calls.  xs3d_pack_bits: boolean image (1 byte per voxel, any
 * memory order) -> one bit per voxel, bit (i & 7) of byte (i >> 3) <-> byte i != 0, packed by the library's host thread
 * pool (`bits` holds (voxels + 7) / 8 bytes).  xs3d_volume_load_packed: ingest such a bit stream (host or device memory;
 * c_order as in xs3d_volume_load).  xs3d_volume_load does both itself for large host images; the split lets a caller
 * pack image k+1 while image k is being swept, or keep its masks bit-packed (1/8 of the reference's bool arrays). */
int xs3d_pack_bits(const uint8_t* binimg, uint64_t voxels, uint8_t* bits);
int xs3d_volume_load_packed(xs3d_volume* vol, const uint8_t* bits, int mem, int c_order);
/* Upload and ingest ahead of use.  stage copies a bit-packed HOST image (pinned, for the copy to be asynchronous) into
 * the spare buffers of `slot` (0 or 1) and builds occupancy bytes from it, all on the volume's own upload stream, and
 * returns at once; commit makes the launching stream wait for that work and swaps the spare occupancy bytes in (no
 * kernel, no host synchronisation).  A pipeline stages image k+1 (from any host thread) while image k is being swept;
 * a slot may be staged again once it has been committed; that call first waits (on the host) until the copies of the
 * slot's previous image have left its host buffer — `bits` must stay untouched until then, i.e. until the next stage
 * call on the same slot returns, or until a batch that followed commit_staged(slot) has returned.  c_order as in
 * xs3d_volume_load. */
int xs3d_volume_stage_packed(xs3d_volume* vol, const uint8_t* bits, int slot, int c_order);
/* stage for an image that is still bytes (1 per voxel, host memory, pinned or not): packed by the host pool, each eighth
 * of the bit stream copied as soon as it is packed, so that the transfer ends with the packing.  Returns when the image
 * has been read (the caller may reuse it); copy and ingest may still be in flight. */
int xs3d_volume_pack_and_stage(xs3d_volume* vol, const uint8_t* binimg, int slot, int c_order);
int xs3d_volume_commit_staged(xs3d_volume* vol, int slot);
/* For a host layer that moves staged occupancy bytes between GPUs itself (NCCL broadcast on its own stream, beside the
 * sweep of the previous image): the spare occupancy buffer of a slot; a hook that makes `cuda_stream` wait for what
 * stage_packed / pack_and_stage put there; and the call that declares the buffer written by everything enqueued on
 * `cuda_stream` so far, so that commit_staged waits for exactly that. */
int xs3d_volume_staged_cubes(xs3d_volume* vol, int slot, void** device_ptr, uint64_t* bytes);
int xs3d_volume_wait_staged(xs3d_volume* vol, int slot, void* cuda_stream);
int xs3d_volume_mark_staged(xs3d_volume* vol, int slot, void* cuda_stream);
/* One host image that several processes (one per GPU) can all read — shared memory, or a replica each: every process
 * packs and uploads only ITS part of the bit stream, so the host-memory-bound packing is spread over the cores of all
 * ranks.  staged_bits: the slot's device bit buffer, at least min_bytes long ((voxels + 7) / 8 bytes are the image; an
 * all-gather of equal parts may need a little more).  pack_and_stage_part: voxels [lo, hi) (lo a multiple of 4096) ->
 * bits -> their place in that buffer, on the volume's upload stream; returns when the part has been read.
 * wait_staged_bits: `cuda_stream` waits for that copy; the host layer then all-gathers the parts into the buffer on that
 * stream.  ingest_staged: bits -> the slot's spare occupancy bytes on `cuda_stream` (behind the all-gather) and marks
 * the slot staged, so that commit_staged waits for exactly that.  c_order as in xs3d_volume_load. */
int xs3d_volume_staged_bits(xs3d_volume* vol, int slot, uint64_t min_bytes, void** device_ptr, uint64_t* bytes);
int xs3d_volume_pack_and_stage_part(xs3d_volume* vol, const uint8_t* binimg, int slot, uint64_t lo_voxel, uint64_t hi_voxel);
int xs3d_volume_wait_staged_bits(xs3d_volume* vol, int slot, void* cuda_stream);
int xs3d_volume_ingest_staged(xs3d_volume* vol, int slot, int c_order, void* cuda_stream);
/* Attach a label volume for xs3d_projection_v (any 1/2/4/8-byte dtype, C or F order; fastxs3d.cpp:135,199-215). */
int xs3d_volume_load_labels(xs3d_volume* vol, const void* labels, int itemsize, int mem, int c_order);
/* The device buffer of the label volume, (re)allocated for that item size / order without an upload: for a host layer
 * that fills it itself (an NCCL broadcast from the rank that uploaded the labels).  Counts as loaded afterwards. */
int xs3d_volume_labels_buffer(xs3d_volume* vol, int itemsize, int c_order, void** device_ptr, uint64_t* bytes);
/* Make the resident image "labels == label" (labels attached with xs3d_volume_load_labels): the occupancy bytes
 * are rebuilt on the device from the label volume, so one uploaded segmentation serves every object in it. */
int xs3d_volume_select_label(xs3d_volume* vol, uint64_t label);
/* Device pointer + size of the ingested occupancy bytes (for an NCCL broadcast by the host layer),
 * and the call that marks them valid after such a broadcast wrote them. */
int xs3d_volume_cubes(xs3d_volume* vol, void** device_ptr, uint64_t* bytes);
int xs3d_volume_mark_loaded(xs3d_volume* vol);
int xs3d_volume_shape(xs3d_volume* vol, uint64_t shape[3]);

/* ---- batched entry point (new): many (vertex, normal) queries against one resident volume.
 *      Same per-query semantics as fastxs3d.area(..., slow_method=False)  (xs3d.hpp:1305-1365). */
int xs3d_area_batch(xs3d_volume* vol, uint64_t n, const float* pos, const float* normal,
                    const float anisotropy[3], int mem, float* area, uint8_t* contact);
/* Skeleton sweeps (the use the reference is written for, README.md:59-71): unit tangent per skeleton
 * vertex from (vertices float32[n][3] in voxel coordinates, edges int64[m][2]), computed on the device — the normals
 * of the section planes, ready for xs3d_area_batch with the same `mem`.  tangent = normalised sum of the incident edge
 * directions, each flipped to agree with the vertex's first incident half-edge; (0, 0, 1) for an isolated vertex. */
int xs3d_skeleton_tangents(xs3d_volume* vol, uint64_t n_vertices, const float* vertices, uint64_t n_edges,
                           const int64_t* edges, int mem, float* normals);
/* One whole sweep from host buffers in a single call: upload the queries, upload + ingest `binimg` into `vol`
 * (which must have the image's shape), run the batch, read areas + contacts back.  Equivalent to
 * xs3d_volume_load(HOST) followed by xs3d_area_batch(HOST), but with the copies ordered so that several sweeps in
 * flight on different volumes overlap their uploads with each other's kernels. */
int xs3d_area_sweep_host(xs3d_volume* vol, const uint8_t* binimg, int c_order, uint64_t n, const float* pos,
                         const float* normal, const float anisotropy[3], float* area, uint8_t* contact);
/* Counters of the last xs3d_area_batch on this volume:
 * stats[0] lattice samples, [1] contributing blocks (items), [2] lattice cells tested, [3] queries escalated
 * warp->mid tier, [4] mid->big tier, [5] kernel launches of the call, [6] big->worst-case slab,
 * [7] polygon records evaluated, [8] queries that failed even with the worst-case slab, [9] lattice samples handled by the CTA tiers
 * (the warp tier handled stats[0] - stats[9]). */
int xs3d_volume_stats(xs3d_volume* vol, uint64_t stats[10]);
/* SM-clock cycles thread 0 of each query group spent per phase in the last xs3d_area_batch, summed over
 * queries: prof[0..5] warp tier, prof[6..11] CTA tiers (mid + big); phases: [0] lattice flood (tiles + walk), [1] claims,
 * [2] ownership + record emission, [3] [4] unused, [5] walk only. */
int xs3d_volume_profile(xs3d_volume* vol, uint64_t prof[12]);
/* CUDA-event durations (ms) of the kernels of the last call on this volume:
 * ms[0] warp-tier, ms[1] mid-tier (on its own stream, concurrent with the warp tier; tiny batches: mid tier + its second
 * half), ms[2] second round (wide + big tiers and everything behind them, only when a section left the mid tier's window),
 * ms[3] worst-case query kernels, ms[4] ingest kernel (last xs3d_volume_load), ms[5] polygon kernel, ms[6] combine kernel, ms[7] the CTA
 * tiers' second half (claims / ownership / record kernels over the parked sections, second stream),
 * ms[8] time the launching stream waits for the second stream (mid / big tiers + their polygon and combine kernels)
 * after its own combine kernel, ms[9] pre-pass (size estimate + ordering), ms[10] second launch of the mid tier
 * (late hand-overs; includes waiting for the warp tier), ms[11] the second stream's polygon + combine kernels. */
int xs3d_volume_timing(xs3d_volume* vol, float ms[12]);
/* Launch on the caller's stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream) instead of the
 * volume's own, so the caller's CUDA events bracket the work.  The stream must outlive the volume. */
int xs3d_volume_set_stream(xs3d_volume* vol, void* cuda_stream);
/* Pipelining hook for callers that stream batches: `fn(ctx)` is called on the calling thread by xs3d_area_batch /
 * xs3d_area_sweep_host once a batch (kernels + result copies) is enqueued, before the host waits for it.  Whatever the
 * callback enqueues on the volume's stream (the uploads of the caller's next step, ...) runs right behind the batch and does
 * not delay this batch's results: the host waits for an event recorded before the callback.  fn == NULL switches it off. */
int xs3d_volume_set_after_enqueue(xs3d_volume* vol, void (*fn)(void*), void* ctx);

/* ---- drop-in single-call entry points: one per fastxs3d export ------------------------------------- */
/* fastxs3d.area (fastxs3d.cpp:82-117).  `binimg` is always the image that is evaluated (it is uploaded on every
 * call).  use_persistent_data has the reference's meaning (xs3d.hpp:963-971): only scratch is reused — here the
 * process-wide default volume and its workspaces, which xs3d_set_shape() creates ahead of time — so the flag never
 * changes a result.  Callers that want to skip the upload keep a resident volume (xs3d_volume_* + xs3d_area_batch). */
int xs3d_area(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
              const float pos[3], const float normal[3], const float anisotropy[3],
              int slow_method, int use_persistent_data, float* area, uint8_t* contact);
/* xs3d.cross_sectional_area on a 2-D image: the reference's pure-Python cross_sectional_area_2d (xs3d/twod.py:73-174,
 * called at xs3d/__init__.py:79-80; its third-party dependency cc3d is only used to find the 8-connected component of
 * the seed pixel among the marked pixels).  `binimg`: HOST, x fastest ([sx][sy] Fortran order), 1 byte per pixel.
 * `vec` is the 2-D "normal" argument.  area is float64 like the reference's; contact bits: 0 x == 0, 1 x == sx-1,
 * 2 y == 0, 3 y == sy-1; a vertex outside the image or on background gives (0.0, 0b00111111) (twod.py:83-90).
 * XS3D_ERR_ARG with the message "more than two distinct intersection points..." where the reference raises ValueError
 * (twod.py:163-164). */
int xs3d_area_2d(const uint8_t* binimg, uint64_t sx, uint64_t sy, const float pos[2], const float vec[2],
                 const float anisotropy[2], double* area, uint8_t* contact);
/* fastxs3d.section (fastxs3d.cpp:13-80).  `out` is a HOST float32 F-order buffer of
 * sx*sy*sz elements (method 0/1) or ceil(sx/2)*ceil(sy/2)*ceil(sz/2) (method 2) that the caller has
 * zero-filled (the reference's std::fill, fastxs3d.cpp:42-44); non-zero contributions are scattered into it. */
int xs3d_section(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz,
                 const float pos[3], const float normal[3], const float anisotropy[3],
                 int method, float* out, uint8_t* contact);
/* Sparse form of the same result on a resident volume: up to `cap` (linear index, value) pairs, HOST buffers. */
int xs3d_section_sparse(xs3d_volume* vol, const float pos[3], const float normal[3], const float anisotropy[3],
                        int method, uint64_t* idx, float* val, uint64_t cap, uint64_t* nnz, uint8_t* contact);
/* fastxs3d.projection (fastxs3d.cpp:119-216).  Two-call protocol: the first call (out == NULL) returns the
 * cut-out shape; the second fills `out` (HOST, F-order [shape0][shape1], itemsize bytes per element). */
int xs3d_projection(const void* labels, uint64_t sx, uint64_t sy, uint64_t sz, int itemsize, int c_order,
                    const float pos[3], const float normal[3], const float anisotropy[3],
                    int standardize_basis, float crop_distance, void* out, int64_t shape[2]);
/* Batched slices on a resident label volume (fastxs3d.cpp:119-216 for n planes at once; cropped or not — an
 * un-cropped plane through a 1024^3 volume is a ~1700 x 1700 cut-out).  Ragged output in two steps:
 *   xs3d_projection_plan : sizes the cut-outs of n planes (fastxs3d.cpp:137-154 lattice, xs3d.hpp:1554-1621 fill as
 *        row summaries, :1585-1588 bounding box).  pos / normal: float32[n][3] in `mem`; shapes (int64[2n]) and offsets
 *        (uint64[n+1], in elements, offsets[n] = total) are HOST arrays.  crop_distance == 0 gives 0 x 0 cut-outs.
 *   xs3d_projection_fill : gathers planes [first, first + count) of the LAST plan on this volume into `out` (host or
 *        device memory per `mem`): plane i's cut-out, F-order [shapes[2i]][shapes[2i+1]], starts at element
 *        offsets[i] - offsets[first].  Any sub-range, any number of times — a caller bounds its buffer by filling the
 *        plan in pieces.
 * xs3d_projection_batch is the fixed-pitch form (plan + fill in one call): plane i at out + i * pitch_elems * itemsize;
 * a cut-out larger than pitch_elems is reported in shapes, flagged overflow[i] = 1 and not written (status XS3D_OK).
 * shapes / overflow live in `mem` like out.  The labels' dtype is only an item size (fastxs3d.cpp:199-215). */
int xs3d_projection_plan(xs3d_volume* vol, uint64_t n, const float* pos, const float* normal,
                         const float anisotropy[3], int standardize_basis, float crop_distance, int mem,
                         int64_t* shapes, uint64_t* offsets);
int xs3d_projection_fill(xs3d_volume* vol, uint64_t first, uint64_t count, void* out, int mem);
int xs3d_projection_batch(xs3d_volume* vol, uint64_t n, const float* pos, const float* normal,
                          const float anisotropy[3], int standardize_basis, float crop_distance,
                          void* out, uint64_t pitch_elems, int64_t* shapes, uint8_t* overflow, int mem);
/* fastxs3d.set_shape / clear_shape (fastxs3d.cpp:224-225).  set_shape creates the default volume of that shape and
 * its workspaces (the GPU counterpart of the reference's pre-sized visited buffer); xs3d_set_shape_image also uploads
 * an image into it (a warm start only: later calls still upload the image they are given); clear_shape frees it. */
int xs3d_set_shape(uint64_t sx, uint64_t sy, uint64_t sz);
int xs3d_set_shape_image(const uint8_t* binimg, uint64_t sx, uint64_t sy, uint64_t sz, int c_order);
int xs3d_clear_shape(void);

#ifdef __cplusplus
}
#endif
#endif /* XS3D_B200_H_ */
