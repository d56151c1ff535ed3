"""xs3d (B200 build): drop-in host layer for seung-lab/cross-section's Python API.

Same names, argument meaning, validation and error behaviour as the reference package
(/root/reference/xs3d/__init__.py): `cross_sectional_area` (:11), `cross_section` (:95), `slice` (:166),
`set_shape` (:234), `clear_shape` (:245).  The pybind11 extension `fastxs3d` is replaced by
libxs3d_b200.so (hand-written sm_100a CUDA behind a C ABI, include/xs3d_b200.h) reached through ctypes.

Additions for skeleton sweeps: `Volume` (a volume resident on one GPU) and
`cross_sectional_area_batch` (many (vertex, normal) queries per call); `xs3d.dist` shards a query
list over ranks.

There is no CPU fallback: every function raises if the CUDA extension is missing or fails.
The 2-D path of the reference (xs3d/twod.py, needs the third-party cc3d) is out of scope.
"""
import ctypes as C
from typing import Optional, Union

import numpy as np
import builtins
from collections.abc import Sequence
import numpy.typing as npt

from . import _abi
from ._abi import XS3DError, HOST, DEVICE

POINT_T = Union[tuple, npt.NDArray[np.integer]]
VECTOR_T = Union[tuple, npt.NDArray[np.float32]]

__all__ = [
  "cross_sectional_area", "cross_section", "slice", "set_shape", "clear_shape",
  "Volume", "SliceBatch", "cross_sectional_area_batch", "skeleton_tangents", "sweep_skeleton", "pack_bits", "XS3DError",
]


def _is_torch_tensor(x):
  return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


def _validate_vectors(pos, normal, anisotropy, ndim):
  """Coercion + validation of xs3d/__init__.py:59-75 (same exceptions, same messages)."""
  if anisotropy is None:
    anisotropy = (1.0, 1.0, 1.0)
    if ndim == 2:
      anisotropy = (1.0, 1.0)
  pos = np.asarray(pos, dtype=np.float32)
  normal = np.asarray(normal, dtype=np.float32)
  anisotropy = np.asarray(anisotropy, dtype=np.float32)
  if np.any(anisotropy <= 0):
    raise ValueError(f"anisotropy values must be > 0. Got: {anisotropy}")
  if np.all(normal == 0):
    raise ValueError("normal vector must not be a null vector (all zeros).")
  return pos, normal, anisotropy


def _image_pointer(binimg):
  """(array kept alive, pointer, c_order flag) for a 3-D bool/uint8 ndarray without a host-side
  transpose copy: C-contiguous input is ingested as such on the device (replaces the
  np.asfortranarray of xs3d/__init__.py:77)."""
  if binimg.flags.f_contiguous:
    a = binimg
    c_order = 0
  elif binimg.flags.c_contiguous:
    a = binimg
    c_order = 1
  else:
    a = np.asfortranarray(binimg)
    c_order = 0
  if a.dtype == bool:
    a = a.view(np.uint8)
  return a, a.ctypes.data, c_order


def _check_out(out, n):
  """User-supplied result buffers are written through raw pointers by the C ABI: refuse anything that is not exactly
  what it writes (float32 areas, uint8 contacts, >= n elements, C-contiguous, writeable)."""
  try:
    area, contact = out
  except Exception:
    raise ValueError("out must be a pair (area float32 [n], contact uint8 [n])") from None
  for a, dt, name in ((area, np.float32, "area"), (contact, np.uint8, "contact")):
    if not isinstance(a, np.ndarray) or a.dtype != dt or a.ndim != 1 or a.size < n or not a.flags.c_contiguous or not a.flags.writeable:
      raise ValueError(f"out: {name} must be a writeable C-contiguous 1-D {np.dtype(dt).name} ndarray of at least {n} elements")
  return area, contact


def _queries(positions, normals):
  """float32 [n,3] positions and normals of equal length (ValueError otherwise, as the single-query API raises)."""
  pos = np.ascontiguousarray(positions, dtype=np.float32)
  nrm = np.ascontiguousarray(normals, dtype=np.float32)
  if pos.size % 3 or nrm.size % 3:
    raise ValueError("positions and normals must be [n,3] arrays")
  pos = pos.reshape(-1, 3)
  nrm = nrm.reshape(-1, 3)
  if pos.shape != nrm.shape:
    raise ValueError("positions and normals must have the same length")
  if np.any(np.all(nrm == 0, axis=1)):
    raise ValueError("normal vector must not be a null vector (all zeros).")
  return pos, nrm


class SliceBatch(Sequence):
  """The cut-outs of Volume.slice_batch: a sequence of Fortran-ordered 2-D arrays, built as views when an element is
  accessed (10^4 tiny slices cost no Python loop).  Ragged form: `chunks` = [(first plane, 1-D buffer)] + `offsets`
  (elements, per plane) + `shapes`; fixed-pitch form: `buffer` [n, pitch] + `shapes`."""

  def __init__(self, buffer, shapes, offsets=None, chunks=None):
    self.buffer, self.shapes, self.offsets, self.chunks = buffer, shapes, offsets, chunks
    if chunks is not None:
      self._firsts = np.array([c[0] for c in chunks], dtype=np.int64)

  def __len__(self):
    return len(self.shapes)

  def __getitem__(self, i):
    if isinstance(i, builtins.slice):      # (this module defines its own `slice`)
      return [self[j] for j in range(*i.indices(len(self)))]
    if i < 0:
      i += len(self)
    if not 0 <= i < len(self):
      raise IndexError(i)
    h, w = int(self.shapes[i, 0]), int(self.shapes[i, 1])
    if self.chunks is None:
      return self.buffer[i, : h * w].reshape((h, w), order="F")
    k = int(np.searchsorted(self._firsts, i, side="right")) - 1
    first, buf = self.chunks[k]
    lo = int(self.offsets[i] - self.offsets[first])
    return buf[lo: lo + h * w].reshape((h, w), order="F")


class Volume:
  """A 3-D image resident on one GPU (ingested once; any number of queries afterwards).

  binimg: bool ndarray (3-D, any memory order) or a CUDA torch tensor of dtype bool/uint8
          (C-contiguous, zero-copy through its data pointer).
  labels: optional label volume (any 1/2/4/8-byte dtype) for `slice`/`slice_batch`.
  """

  def __init__(self, binimg=None, labels=None, device: int = 0, shape=None):
    self._h = None
    self._lib = _abi.lib()
    src = binimg if binimg is not None else labels
    if shape is None:
      if src is None:
        raise ValueError("Volume needs an image, labels or a shape")
      shape = tuple(int(s) for s in src.shape)
    if len(shape) != 3:
      raise ValueError("dimensions not supported")
    self.shape = tuple(int(s) for s in shape)
    self.device = int(device)
    h = C.c_void_p()
    _abi.check(self._lib.xs3d_volume_create(self.device, *self.shape, C.byref(h)))
    self._h = h
    self._keep = [None, None]          # host buffers whose asynchronous upload may still be in flight, per staging slot
    if binimg is not None:
      self.load(binimg)
    if labels is not None:
      self.load_labels(labels)

  # -- lifetime
  def close(self):
    if self._h is not None:
      self._lib.xs3d_volume_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def __enter__(self):
    return self

  def __exit__(self, *a):
    self.close()

  def _order_after_producers(self, tensor):
    """The library launches on its own (non-blocking) stream unless set_stream() handed it torch's: wait for whatever
    produced a CUDA tensor on torch's current stream (e.g. the .contiguous() copy) before its pointer is read."""
    if tensor.is_cuda:
      import torch
      cur = torch.cuda.current_stream(tensor.device)
      if getattr(self, "_stream", None) != cur.cuda_stream:
        cur.synchronize()

  # -- loading
  def load(self, binimg):
    if tuple(int(s) for s in binimg.shape) != self.shape:
      raise ValueError("image shape does not match the volume")
    if _is_torch_tensor(binimg):
      import torch
      if binimg.dtype not in (torch.bool, torch.uint8):
        raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
      c_order = 1
      if not binimg.is_contiguous():
        if binimg.permute(2, 1, 0).is_contiguous():
          c_order = 0                    # Fortran-ordered tensor: ingested as it is, no transposing copy
        else:
          binimg = binimg.contiguous()
      mem = DEVICE if binimg.is_cuda else HOST
      self._order_after_producers(binimg)
      _abi.check(self._lib.xs3d_volume_load(self._h, binimg.data_ptr(), mem, c_order))
      return
    if binimg.dtype != bool and binimg.dtype != np.uint8:
      raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
    a, ptr, c_order = _image_pointer(binimg)
    _abi.check(self._lib.xs3d_volume_load(self._h, ptr, HOST, c_order))

  def load_packed(self, bits, c_order=False):
    """Ingest a bit-packed image as xs3d.pack_bits() writes it (uint8 ndarray, or a torch uint8 tensor in host or
    device memory): 1/8 of the bytes of the boolean image cross PCIe.  c_order: the bits follow the C-contiguous
    voxel order instead of the Fortran one."""
    nbytes = (int(np.prod(self.shape)) + 7) // 8
    if _is_torch_tensor(bits):
      if bits.numel() * bits.element_size() < nbytes or not bits.is_contiguous():
        raise ValueError("bit stream too short for the volume (or not contiguous)")
      self._order_after_producers(bits)
      _abi.check(self._lib.xs3d_volume_load_packed(self._h, bits.data_ptr(), DEVICE if bits.is_cuda else HOST, int(bool(c_order))))
      return
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    if bits.size < nbytes:
      raise ValueError("bit stream too short for the volume")
    _abi.check(self._lib.xs3d_volume_load_packed(self._h, bits.ctypes.data, HOST, int(bool(c_order))))

  def stage_packed(self, bits, slot=0, c_order=False):
    """Start the upload of a bit-packed HOST image (pinned memory for an asynchronous copy) into the spare buffers of
    `slot` (0 or 1) and its ingest, both on the volume's upload stream; returns at once.  commit_staged(slot) swaps
    the result in."""
    nbytes = (int(np.prod(self.shape)) + 7) // 8
    if _is_torch_tensor(bits):
      if bits.is_cuda or bits.numel() * bits.element_size() < nbytes or not bits.is_contiguous():
        raise ValueError("stage_packed needs a contiguous HOST buffer of (voxels + 7) // 8 bytes")
      ptr = bits.data_ptr()
    else:
      if not isinstance(bits, np.ndarray) or bits.nbytes < nbytes or not bits.flags.c_contiguous:
        raise ValueError("stage_packed needs a contiguous HOST buffer of (voxels + 7) // 8 bytes")
      ptr = bits.ctypes.data
    if int(slot) not in (0, 1):
      raise ValueError("slot must be 0 or 1")
    # the call first waits for the copies of this slot's previous image, so the buffer kept for it can go afterwards
    _abi.check(self._lib.xs3d_volume_stage_packed(self._h, ptr, int(slot), int(bool(c_order))))
    self._keep[int(slot)] = bits

  def pack_and_stage(self, binimg, slot=0):
    """stage_packed for a boolean image that is still bytes (either memory order): bit-packed by the host thread pool,
    each part of the bit stream copied as soon as it is packed, occupancy bytes built behind the last copy."""
    if tuple(int(s) for s in binimg.shape) != self.shape:
      raise ValueError("image shape does not match the volume")
    if binimg.dtype != bool and binimg.dtype != np.uint8:
      raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
    a, ptr, c_order = _image_pointer(binimg)
    _abi.check(self._lib.xs3d_volume_pack_and_stage(self._h, ptr, int(slot), int(c_order)))

  def commit_staged(self, slot=0):
    """Make the image staged in `slot` the resident one: the launching stream waits (on the device) for the staged
    upload + ingest, the occupancy bytes change places.  No kernel, no host synchronisation."""
    _abi.check(self._lib.xs3d_volume_commit_staged(self._h, int(slot)))

  def staged_cubes(self, slot=0):
    """(device pointer, bytes) of the spare occupancy bytes of `slot` (for an NCCL broadcast into / out of them)."""
    p = C.c_void_p()
    n = C.c_uint64()
    _abi.check(self._lib.xs3d_volume_staged_cubes(self._h, int(slot), C.byref(p), C.byref(n)))
    return p.value, n.value

  def wait_staged(self, slot, cuda_stream):
    """Make `cuda_stream` wait (on the device) for what stage_packed / pack_and_stage put into `slot`."""
    _abi.check(self._lib.xs3d_volume_wait_staged(self._h, int(slot), C.c_void_p(int(cuda_stream))))

  def mark_staged(self, slot, cuda_stream):
    """Declare the spare occupancy bytes of `slot` written by everything enqueued on `cuda_stream` so far."""
    _abi.check(self._lib.xs3d_volume_mark_staged(self._h, int(slot), C.c_void_p(int(cuda_stream))))

  # -- one host image that every rank can read: each rank packs and uploads its part (xs3d.dist.stream_shared_volumes)
  def set_after_enqueue(self, fn):
    """`fn()` (or None) is called by every batched area call once the batch — kernels and result copies — is enqueued and
    before the host waits for it: what it enqueues on the volume's stream (the next step's uploads, ...) runs right
    behind the batch without delaying this batch's results."""
    if fn is None:
      self._after_cb = None
      _abi.check(self._lib.xs3d_volume_set_after_enqueue(self._h, None, None))
      return
    self._after_cb = C.CFUNCTYPE(None, C.c_void_p)(lambda _ctx: fn())      # kept alive with the volume
    _abi.check(self._lib.xs3d_volume_set_after_enqueue(self._h, C.cast(self._after_cb, C.c_void_p), None))

  def staged_bits(self, slot=0, min_bytes=0):
    """(device pointer, bytes of the image's bit stream) of the bit buffer of `slot`, at least `min_bytes` long."""
    p = C.c_void_p()
    n = C.c_uint64()
    _abi.check(self._lib.xs3d_volume_staged_bits(self._h, int(slot), int(min_bytes), C.byref(p), C.byref(n)))
    return p.value, n.value

  def pack_and_stage_part(self, binimg, slot, lo_voxel, hi_voxel):
    """Bit-pack voxels [lo, hi) of the boolean image (in memory order; lo a multiple of 4096) and copy them to their place
    in the bit buffer of `slot` on the upload stream.  Returns (c_order of the image) once the part has been read."""
    if tuple(int(s) for s in binimg.shape) != self.shape:
      raise ValueError("image shape does not match the volume")
    if binimg.dtype != bool and binimg.dtype != np.uint8:
      raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
    a, ptr, c_order = _image_pointer(binimg)
    _abi.check(self._lib.xs3d_volume_pack_and_stage_part(self._h, ptr, int(slot), int(lo_voxel), int(hi_voxel)))
    return bool(c_order)

  def wait_staged_bits(self, slot, cuda_stream):
    _abi.check(self._lib.xs3d_volume_wait_staged_bits(self._h, int(slot), C.c_void_p(int(cuda_stream))))

  def ingest_staged(self, slot, c_order, cuda_stream):
    """Bit buffer of `slot` -> its spare occupancy bytes on `cuda_stream`; commit_staged(slot) then swaps them in."""
    _abi.check(self._lib.xs3d_volume_ingest_staged(self._h, int(slot), int(bool(c_order)), C.c_void_p(int(cuda_stream))))

  def load_labels(self, labels):
    if tuple(int(s) for s in labels.shape) != self.shape:
      raise ValueError("labels shape does not match the volume")
    if _is_torch_tensor(labels):
      if not labels.is_contiguous():
        labels = labels.contiguous()
      mem = DEVICE if labels.is_cuda else HOST
      self._order_after_producers(labels)
      _abi.check(self._lib.xs3d_volume_load_labels(self._h, labels.data_ptr(), labels.element_size(), mem, 1))
      self._label_dtype = None
      return
    if not labels.flags.f_contiguous:
      labels = np.ascontiguousarray(labels)           # xs3d/__init__.py:225-226
    c_order = 1 if labels.flags.c_contiguous else 0   # fastxs3d.cpp:135
    if labels.dtype.itemsize not in (1, 2, 4, 8):
      raise ValueError("labels must have a 1, 2, 4 or 8 byte dtype")
    self._label_dtype = labels.dtype
    _abi.check(self._lib.xs3d_volume_load_labels(self._h, labels.ctypes.data, labels.dtype.itemsize, HOST, c_order))

  def labels_buffer(self, dtype, c_order):
    """(device pointer, nbytes) of the label buffer, allocated for `dtype` / memory order without an upload (the caller
    fills it: xs3d.dist.broadcast_labels)."""
    dt = np.dtype(dtype)
    p = C.c_void_p()
    n = C.c_uint64()
    _abi.check(self._lib.xs3d_volume_labels_buffer(self._h, dt.itemsize, int(bool(c_order)), C.byref(p), C.byref(n)))
    self._label_dtype = dt
    return p.value, int(n.value)

  def select_label(self, label):
    """Make the resident binary image `labels == label` (labels loaded with load_labels / Volume(labels=...)):
    rebuilt on the device, nothing is uploaded.  One resident segmentation then serves every object in it."""
    _abi.check(self._lib.xs3d_volume_select_label(self._h, C.c_uint64(int(label) & 0xFFFFFFFFFFFFFFFF)))

  def cubes(self):
    """(device pointer, nbytes) of the ingested 2x2x2 occupancy bytes (used by xs3d.dist.broadcast)."""
    p = C.c_void_p()
    n = C.c_uint64()
    _abi.check(self._lib.xs3d_volume_cubes(self._h, C.byref(p), C.byref(n)))
    return p.value, int(n.value)

  def mark_loaded(self):
    _abi.check(self._lib.xs3d_volume_mark_loaded(self._h))

  def stats(self):
    s = (C.c_uint64 * 10)()
    _abi.check(self._lib.xs3d_volume_stats(self._h, s))
    return {"samples": s[0], "blocks": s[1], "cells_tested": s[2], "escalated_mid": s[3], "escalated_big": s[4],
            "launches": s[5], "escalated_worst_case": s[6], "polygons": s[7], "failed": s[8], "samples_cta_tiers": s[9]}

  def timing(self):
    """CUDA-event durations (ms) of the last call (include/xs3d_b200.h, xs3d_volume_timing).  The mid / big tiers
    run on a second stream beside the warp tier; mid_exposed_ms is what they add to the launching stream."""
    t = (C.c_float * 12)()
    _abi.check(self._lib.xs3d_volume_timing(self._h, t))
    return {"warp_tier_ms": t[0], "mid_tier_ms": t[1], "second_round_ms": t[2], "big_tier_ms": t[2], "worst_case_ms": t[3], "ingest_ms": t[4],
            "polygon_ms": t[5], "combine_ms": t[6], "p2_ms": t[7], "mid_exposed_ms": t[8], "prepass_ms": t[9],
            "mid_second_launch_ms": t[10], "second_stream_poly_combine_ms": t[11]}

  def set_stream(self, cuda_stream):
    """Launch on the given cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)."""
    _abi.check(self._lib.xs3d_volume_set_stream(self._h, C.c_void_p(int(cuda_stream))))
    self._stream = int(cuda_stream)

  def profile(self):
    """Per-phase SM cycles of the last batch (thread 0 of each query group, summed over queries)."""
    p = (C.c_uint64 * 12)()
    _abi.check(self._lib.xs3d_volume_profile(self._h, p))
    names = ["flood", "claims", "emit", "tile_sampling", "emit_pass1", "walk_only"]
    return {"warp_tier": dict(zip(names, p[0:6])), "cta_tier": dict(zip(names, p[6:12]))}

  # -- queries
  def cross_sectional_area_batch(self, positions, normals, anisotropy=None, return_contact=False, out=None):
    """Areas (float32 [n]) of n (vertex, normal) queries; optionally the contact bitfields (uint8 [n]).

    positions / normals: float32 [n,3] ndarrays, or CUDA torch tensors (then results are CUDA tensors,
    nothing crosses PCIe).  Per-query semantics are exactly xs3d.cross_sectional_area's.
    """
    if anisotropy is None:
      anisotropy = (1.0, 1.0, 1.0)
    w = _abi.f3(anisotropy)
    if np.any(w <= 0):
      raise ValueError(f"anisotropy values must be > 0. Got: {w}")
    if _is_torch_tensor(positions):
      import torch
      if not _is_torch_tensor(normals):
        normals = torch.as_tensor(np.asarray(normals, dtype=np.float32))
      if positions.numel() % 3 or normals.numel() % 3:
        raise ValueError("positions and normals must be [n,3] arrays")
      pos = positions.to(dtype=torch.float32).contiguous().reshape(-1, 3)
      nrm = normals.to(dtype=torch.float32, device=pos.device).contiguous().reshape(-1, 3)
      if pos.shape != nrm.shape:
        raise ValueError("positions and normals must have the same length")
      if not pos.is_cuda:
        raise ValueError("torch inputs must be CUDA tensors (pass ndarrays for host buffers)")
      n = pos.shape[0]
      area = torch.empty(n, dtype=torch.float32, device=pos.device)
      contact = torch.empty(n, dtype=torch.uint8, device=pos.device)
      cur = torch.cuda.current_stream(pos.device)
      if getattr(self, "_stream", None) != cur.cuda_stream:
        cur.synchronize()              # the library launches on another stream: order it after the producers of pos / nrm
      try:
        _abi.check(self._lib.xs3d_area_batch(self._h, n, pos.data_ptr(), nrm.data_ptr(), _abi.fptr(w), DEVICE,
                                             area.data_ptr(), contact.data_ptr()))
      except XS3DError as e:           # null normals are detected on the device (no host round trip before the batch)
        if "null vector" in str(e):
          raise ValueError("normal vector must not be a null vector (all zeros).") from None
        raise
      return (area, contact) if return_contact else area
    pos, nrm = _queries(positions, normals)
    n = pos.shape[0]
    if out is not None:
      area, contact = _check_out(out, n)
    else:
      area = np.empty(n, dtype=np.float32)
      contact = np.empty(n, dtype=np.uint8)
    _abi.check(self._lib.xs3d_area_batch(self._h, n, pos.ctypes.data, nrm.ctypes.data, _abi.fptr(w), HOST,
                                         area.ctypes.data, contact.ctypes.data))
    return (area, contact) if return_contact else area

  def sweep(self, binimg, positions, normals, anisotropy=None, return_contact=False, out=None):
    """(Re)load `binimg` into this volume and run all queries, in ONE call from host buffers: the natural unit
    when sweeping many objects (one volume + its skeleton vertices per call).  Several Volumes driven from
    different host threads overlap their uploads with each other's kernels."""
    if binimg.dtype != bool and binimg.dtype != np.uint8:
      raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
    if tuple(int(s) for s in binimg.shape) != self.shape:
      raise ValueError("image shape does not match the volume")
    if anisotropy is None:
      anisotropy = (1.0, 1.0, 1.0)
    w = _abi.f3(anisotropy)
    if np.any(w <= 0):
      raise ValueError(f"anisotropy values must be > 0. Got: {w}")
    a, ptr, c_order = _image_pointer(binimg)
    pos, nrm = _queries(positions, normals)
    n = pos.shape[0]
    area, contact = _check_out(out, n) if out is not None else (np.empty(n, dtype=np.float32), np.empty(n, dtype=np.uint8))
    _abi.check(self._lib.xs3d_area_sweep_host(self._h, ptr, c_order, n, pos.ctypes.data, nrm.ctypes.data, _abi.fptr(w),
                                              area.ctypes.data, contact.ctypes.data))
    return (area, contact) if return_contact else area

  def skeleton_tangents(self, vertices, edges):
    """Unit tangent per skeleton vertex, computed on this volume's GPU (same definition as xs3d.skeleton_tangents).
    vertices [M,3] (voxel coordinates) and edges [E,2]: ndarrays -> float32 ndarray; CUDA torch tensors -> CUDA tensor,
    nothing crosses PCIe."""
    if _is_torch_tensor(vertices):
      import torch
      if not vertices.is_cuda:
        raise ValueError("torch inputs must be CUDA tensors (pass ndarrays for host buffers)")
      vv = vertices.to(dtype=torch.float32).contiguous().reshape(-1, 3)
      ee = (edges if _is_torch_tensor(edges) else torch.as_tensor(np.asarray(edges, dtype=np.int64))).to(device=vv.device, dtype=torch.int64).contiguous().reshape(-1, 2)
      out = torch.empty_like(vv)
      self._order_after_producers(vv)
      _abi.check(self._lib.xs3d_skeleton_tangents(self._h, vv.shape[0], vv.data_ptr(), ee.shape[0], ee.data_ptr(), DEVICE, out.data_ptr()))
      return out
    vv = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    ee = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
    out = np.empty_like(vv)
    _abi.check(self._lib.xs3d_skeleton_tangents(self._h, vv.shape[0], vv.ctypes.data, ee.shape[0], ee.ctypes.data, HOST, out.ctypes.data))
    return out

  def sweep_skeleton(self, vertices, edges, anisotropy=None, return_contact=False):
    """Cross-sectional area at every skeleton vertex, planes normal to the skeleton: tangents and sections both on the
    device.  With CUDA tensors for vertices / edges the whole sweep stays in HBM."""
    normals = self.skeleton_tangents(vertices, edges)
    if _is_torch_tensor(vertices):
      import torch
      pos = vertices.to(dtype=torch.float32).contiguous().reshape(-1, 3)
    else:
      pos = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    return self.cross_sectional_area_batch(pos, normals, anisotropy, return_contact)

  def cross_sectional_area(self, pos, normal, anisotropy=None, return_contact=False):
    pos, normal, anisotropy = _validate_vectors(pos, normal, anisotropy, 3)
    a, c = self.cross_sectional_area_batch(pos[None], normal[None], anisotropy, return_contact=True)
    return (float(a[0]), int(c[0])) if return_contact else float(a[0])

  def slice_plan(self, positions, normals, anisotropy=None, standardize_basis=True, crop=float("inf")):
    """Sizes of the cut-outs of n planes through the resident label volume (no labels are read yet):
    (shapes int64 [n,2], offsets uint64 [n+1] in elements).  slice_fill() then gathers any sub-range of the plan."""
    assert crop >= 0.0
    if anisotropy is None:
      anisotropy = (1.0, 1.0, 1.0)
    w = _abi.f3(anisotropy)
    if np.any(w <= 0):
      raise ValueError(f"anisotropy values must be > 0. Got: {w}")
    shapes_of = lambda n: (np.zeros((n, 2), dtype=np.int64), np.zeros(n + 1, dtype=np.uint64))
    if _is_torch_tensor(positions):
      import torch
      if not _is_torch_tensor(normals) or not positions.is_cuda or not normals.is_cuda:
        raise ValueError("positions and normals must both be ndarrays or both CUDA tensors")
      pos = positions.to(dtype=torch.float32).contiguous().reshape(-1, 3)
      nrm = normals.to(dtype=torch.float32).contiguous().reshape(-1, 3)
      if pos.shape != nrm.shape:
        raise ValueError("positions and normals must have the same length")
      if bool((nrm == 0).all(dim=1).any()):
        raise ValueError("normal vector must not be a null vector (all zeros).")
      self._order_after_producers(pos)
      shapes, offsets = shapes_of(pos.shape[0])
      _abi.check(self._lib.xs3d_projection_plan(self._h, pos.shape[0], pos.data_ptr(), nrm.data_ptr(), _abi.fptr(w), int(standardize_basis),
                                                C.c_float(crop), DEVICE, shapes.ctypes.data, offsets.ctypes.data))
      return shapes, offsets
    pos, nrm = _queries(positions, normals)
    shapes, offsets = shapes_of(pos.shape[0])
    _abi.check(self._lib.xs3d_projection_plan(self._h, pos.shape[0], pos.ctypes.data, nrm.ctypes.data, _abi.fptr(w), int(standardize_basis),
                                              C.c_float(crop), HOST, shapes.ctypes.data, offsets.ctypes.data))
    return shapes, offsets

  def slice_fill(self, first, count, out):
    """Gather planes [first, first + count) of the last slice_plan() into `out`: a 1-D ndarray (host) or CUDA torch
    tensor of the labels' dtype holding at least offsets[first + count] - offsets[first] elements."""
    if _is_torch_tensor(out):
      if not out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA tensor (or an ndarray)")
      self._order_after_producers(out)
      _abi.check(self._lib.xs3d_projection_fill(self._h, int(first), int(count), out.data_ptr(), DEVICE))
    else:
      if not isinstance(out, np.ndarray) or not out.flags.c_contiguous or not out.flags.writeable:
        raise ValueError("out must be a writeable contiguous ndarray")
      _abi.check(self._lib.xs3d_projection_fill(self._h, int(first), int(count), out.ctypes.data, HOST))
    return out

  def slice_batch(self, positions, normals, anisotropy=None, standardize_basis=True, crop=float("inf"), pitch=None, max_bytes=1 << 31):
    """xs3d.slice for n planes of the resident label volume, cropped or not.  Returns a `SliceBatch`: a sequence of n
    Fortran-ordered 2-D arrays (views, made on access) over ragged buffers of at most `max_bytes` each (an un-cropped
    plane through a 1024^3 uint64 volume is ~25 MB, so a long list is gathered in pieces).  `pitch` (elements per
    plane) selects the fixed-pitch form instead: one [n, pitch] buffer, XS3DError if a cut-out does not fit."""
    assert crop >= 0.0
    if getattr(self, "_label_dtype", None) is None:
      raise ValueError("load_labels() with an ndarray first")
    dt = self._label_dtype
    if pitch is not None:
      if anisotropy is None:
        anisotropy = (1.0, 1.0, 1.0)
      w = _abi.f3(anisotropy)
      pos, nrm = _queries(positions, normals)
      n = pos.shape[0]
      if crop == 0:
        return [np.zeros((0, 0), dtype=dt, order="F") for _ in range(n)]
      out = np.zeros((n, pitch), dtype=dt)
      shapes = np.zeros((n, 2), dtype=np.int64)
      over = np.zeros(n, dtype=np.uint8)
      _abi.check(self._lib.xs3d_projection_batch(self._h, n, pos.ctypes.data, nrm.ctypes.data, _abi.fptr(w),
                                                 int(standardize_basis), C.c_float(crop), out.ctypes.data, pitch,
                                                 shapes.ctypes.data, over.ctypes.data, HOST))
      if np.any(over):
        raise XS3DError(f"slice_batch: {int(over.sum())} cut-outs exceed the pitch of {pitch} elements; pass a larger pitch")
      return SliceBatch(out, shapes)
    shapes, offsets = self.slice_plan(positions, normals, anisotropy, standardize_basis, crop)
    n = len(shapes)
    budget = max(1, int(max_bytes) // dt.itemsize)
    chunks, first = [], 0
    while first < n:
      last = int(np.searchsorted(offsets, offsets[first] + np.uint64(budget), side="right")) - 1   # planes [first, last) fit the budget
      last = min(n, max(last, first + 1))
      buf = np.empty(int(offsets[last] - offsets[first]), dtype=dt)
      if buf.size:
        self.slice_fill(first, last - first, buf)
      chunks.append((first, buf))
      first = last
    return SliceBatch(None, shapes, offsets, chunks)

  def slice(self, pos, normal, anisotropy=None, standardize_basis=True, crop=float("inf")):
    """xs3d.slice of the RESIDENT label volume: no upload, one plane."""
    pos, normal, anisotropy = _validate_vectors(pos, normal, anisotropy, 3)
    r = self.slice_batch(pos[None], normal[None], anisotropy, standardize_basis, crop)
    return np.asfortranarray(r[0])

  def cross_section_sparse(self, pos, normal, anisotropy=None, method=0):
    """(linear F-order voxel indices uint64 [nnz], contributions float32 [nnz], contact)."""
    pos, normal, anisotropy = _validate_vectors(pos, normal, anisotropy, 3)
    cap = 1 << 16
    while True:
      idx = np.empty(cap, dtype=np.uint64)
      val = np.empty(cap, dtype=np.float32)
      nnz = C.c_uint64(0)
      ct = C.c_uint8(0)
      rc = self._lib.xs3d_section_sparse(self._h, _abi.fptr(_abi.f3(pos)), _abi.fptr(_abi.f3(normal)),
                                         _abi.fptr(_abi.f3(anisotropy)), int(method), idx.ctypes.data,
                                         val.ctypes.data, cap, C.byref(nnz), C.byref(ct))
      if rc == 2 and nnz.value > cap:
        cap = int(nnz.value)
        continue
      _abi.check(rc)
      return idx[:nnz.value], val[:nnz.value], int(ct.value)


def cross_sectional_area_batch(binimg, positions, normals, anisotropy=None, return_contact=False, device=0):
  """Batched xs3d.cross_sectional_area: one upload + ingest of `binimg` (or a resident `Volume`),
  then all queries in one pass on the GPU."""
  if isinstance(binimg, Volume):
    return binimg.cross_sectional_area_batch(positions, normals, anisotropy, return_contact)
  if not _is_torch_tensor(binimg) and binimg.dtype != bool:
    raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
  if binimg.ndim != 3:
    raise ValueError("dimensions not supported")
  if _is_torch_tensor(binimg) or _is_torch_tensor(positions):
    with Volume(binimg, device=device) as vol:
      return vol.cross_sectional_area_batch(positions, normals, anisotropy, return_contact)
  with Volume(shape=binimg.shape, device=device) as vol:
    return vol.sweep(binimg, positions, normals, anisotropy, return_contact)


def pack_bits(binimg, out=None):
  """Boolean image -> one bit per voxel (bit i & 7 of byte i >> 3 <-> voxel i of the image in its own memory order,
  i.e. np.packbits(binimg.ravel(order="K"), bitorder="little")), packed by the library's host thread pool at memory
  speed.  Returns (bits, c_order) for Volume.load_packed.  `out`: optional uint8 buffer (e.g. pinned) of
  (binimg.size + 7) // 8 bytes, ndarray or CPU torch tensor."""
  if binimg.dtype != bool and binimg.dtype != np.uint8:
    raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
  a, ptr, c_order = _image_pointer(binimg)
  nbytes = (a.size + 7) // 8
  if out is None:
    out = np.empty(nbytes, dtype=np.uint8)
  if _is_torch_tensor(out):
    if out.is_cuda or out.numel() * out.element_size() < nbytes or not out.is_contiguous():
      raise ValueError("out must be a contiguous CPU buffer of (binimg.size + 7) // 8 bytes")
    optr = out.data_ptr()
  else:
    if out.dtype != np.uint8 or out.size < nbytes or not out.flags.c_contiguous:
      raise ValueError("out must be a contiguous uint8 buffer of (binimg.size + 7) // 8 bytes")
    optr = out.ctypes.data
  _abi.check(_abi.lib().xs3d_pack_bits(ptr, a.size, optr))
  return out, bool(c_order)


def skeleton_tangents(vertices, edges):
  """Unit tangent per skeleton vertex, in the coordinates of `vertices` (pass voxel coordinates: the reference
  takes plane normals in voxel space, README.md:20-24).  The tangent of a vertex is the normalised sum of its
  incident edge directions, each flipped to agree with the vertex's first edge (so the two edges of a path vertex
  reinforce instead of cancelling).  Vertices without edges get (0, 0, 1)."""
  v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
  e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
  m = len(v)
  src = np.concatenate([e[:, 0], e[:, 1]])
  dst = np.concatenate([e[:, 1], e[:, 0]])
  d = v[dst] - v[src]
  nrm = np.linalg.norm(d, axis=1)
  ok = nrm > 0
  src, d = src[ok], d[ok] / nrm[ok, None]
  first = np.full(m, -1, dtype=np.int64)
  order = np.arange(len(src))[::-1]
  first[src[order]] = order                     # first incident half-edge of every vertex
  ref = np.zeros((m, 3))
  has = first >= 0
  ref[has] = d[first[has]]
  sign = np.where(np.einsum("ij,ij->i", d, ref[src]) < 0, -1.0, 1.0)
  t = np.zeros((m, 3))
  np.add.at(t, src, d * sign[:, None])
  n = np.linalg.norm(t, axis=1)
  out = np.tile(np.array([0.0, 0.0, 1.0]), (m, 1))
  good = n > 0
  out[good] = t[good] / n[good, None]
  return out.astype(np.float32)


def sweep_skeleton(binimg, vertices, edges, anisotropy=None, return_contact=False, device=0):
  """Cross-sectional area at every vertex of a skeleton (vertices [M,3] in voxel coordinates, edges [E,2]),
  section planes normal to the skeleton tangent — the use the reference is written for (README.md:59-71),
  as one batched call: tangents (Volume.skeleton_tangents) and sections are both computed on the device.
  `binimg` may be a bool ndarray or a resident Volume."""
  if isinstance(binimg, Volume):
    return binimg.sweep_skeleton(vertices, edges, anisotropy, return_contact)
  with Volume(binimg, device=device) as vol:
    return vol.sweep_skeleton(vertices, edges, anisotropy, return_contact)


def cross_sectional_area(
  binimg: npt.NDArray[np.bool_],
  pos: POINT_T,
  normal: VECTOR_T,
  anisotropy: Optional[VECTOR_T] = None,
  return_contact: bool = False,
  slow_method: bool = False,
  use_persistent_data: bool = False,
) -> Union[float, tuple]:
  """
  Find the cross sectional area for a given binary image, point, and normal vector
  (signature and semantics of the reference, xs3d/__init__.py:11-93).

  binimg: a binary 3d numpy image (bool).  A resident `Volume` is also accepted.
  pos: the point in the image from which to extract the section, e.g. [5,10,2]
  normal: a vector normal to the section plane (need not be a unit vector)
  anisotropy: resolution of the x, y, and z axis, e.g. [4,4,40]
  return_contact: if true, return (area, contact); contact is the bitfield
    0: 0 X  1: Max X  2: 0 Y  3: Max Y  4: 0 Z  5: Max Z
  slow_method: evaluate every foreground voxel (no connected-component restriction).
  use_persistent_data: reuse the scratch that xs3d.set_shape allocated (as in the reference, where it is
    a pre-allocated visited buffer: xs3d.hpp:963-971).  `binimg` is always the image that is evaluated;
    keep a `Volume` to skip the upload of an unchanged image.

  Returns: physical area covered by the section plane
  """
  if isinstance(binimg, Volume):
    return binimg.cross_sectional_area(pos, normal, anisotropy, return_contact)
  pos, normal, anisotropy = _validate_vectors(pos, normal, anisotropy, binimg.ndim)
  if binimg.dtype != bool:
    raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
  if binimg.ndim == 2:
    # the reference's pure-Python 2-D path (xs3d/twod.py:73, via xs3d/__init__.py:79-80) on the device: xs_twod.cu
    if pos.size != 2 or normal.size != 2 or anisotropy.size != 2:
      raise ValueError("a 2-D image needs 2-D pos, normal and anisotropy")
    a2 = np.asfortranarray(binimg).view(np.uint8)
    pos, normal, anisotropy = (np.ascontiguousarray(v.reshape(-1)) for v in (pos, normal, anisotropy))
    area2 = C.c_double(0.0)
    contact2 = C.c_uint8(0)
    rc = _abi.lib().xs3d_area_2d(a2.ctypes.data, a2.shape[0], a2.shape[1], _abi.fptr(pos), _abi.fptr(normal), _abi.fptr(anisotropy),
                                 C.byref(area2), C.byref(contact2))
    if rc == 2 and b"more than two distinct" in (_abi.lib().xs3d_last_error() or b""):
      raise ValueError(_abi.lib().xs3d_last_error().decode())
    _abi.check(rc)
    return (area2.value, int(contact2.value)) if return_contact else area2.value
  elif binimg.ndim != 3:
    raise ValueError("dimensions not supported")
  L = _abi.lib()
  area = C.c_float(0)
  contact = C.c_uint8(0)
  sx, sy, sz = binimg.shape
  if slow_method or binimg.flags.f_contiguous or not binimg.flags.c_contiguous:
    a = np.asfortranarray(binimg).view(np.uint8)   # no copy when already Fortran-ordered
    _abi.check(L.xs3d_area(a.ctypes.data, sx, sy, sz, _abi.fptr(_abi.f3(pos)), _abi.fptr(_abi.f3(normal)),
                           _abi.fptr(_abi.f3(anisotropy)), int(slow_method), int(use_persistent_data),
                           C.byref(area), C.byref(contact)))
  else:
    # C-contiguous: ingested as such on the device, no host-side transpose copy
    with Volume(binimg) as vol:
      r = vol.cross_sectional_area(pos, normal, anisotropy, return_contact=True)
    area, contact = C.c_float(r[0]), C.c_uint8(r[1])
  if return_contact:
    return (float(area.value), int(contact.value))
  return float(area.value)


def cross_section(
  binimg: npt.NDArray[np.bool_],
  pos: POINT_T,
  normal: VECTOR_T,
  anisotropy: Optional[VECTOR_T] = None,
  return_contact: bool = False,
  method: int = 0,
) -> Union[npt.NDArray[np.float32], tuple]:
  """
  Compute which voxels are intercepted by a section plane (reference: xs3d/__init__.py:95-164).

  Returns: float32 Fortran-order volume where each voxel's value is its contribution to the
  cross sectional area (method 0: the component of `pos`; 1: every foreground voxel;
  2: 2x2x2 blocks at half resolution), optionally with the contact bitfield.
  """
  if anisotropy is None:
    anisotropy = (1.0, 1.0, 1.0)
    if binimg.ndim == 2:
      anisotropy = (1.0, 1.0)
  pos = np.array(pos, dtype=np.float32)
  normal = np.array(normal, dtype=np.float32)
  anisotropy = np.array(anisotropy, dtype=np.float32)
  if binimg.dtype != bool:
    raise ValueError(f"A boolean image is required. Got: {binimg.dtype}")
  if np.any(anisotropy <= 0):
    raise ValueError(f"anisotropy values must be > 0. Got: {anisotropy}")
  if np.all(normal == 0):
    raise ValueError("normal vector must not be a null vector (all zeros).")
  binimg = np.asfortranarray(binimg)
  if binimg.ndim != 3:
    raise ValueError("dimensions not supported")
  sx, sy, sz = binimg.shape
  shape = (sx, sy, sz) if method < 2 else ((sx + 1) // 2, (sy + 1) // 2, (sz + 1) // 2)
  section = np.zeros(shape, dtype=np.float32, order="F")   # fastxs3d.cpp:42-44
  contact = C.c_uint8(0)
  _abi.check(_abi.lib().xs3d_section(binimg.view(np.uint8).ctypes.data, sx, sy, sz, _abi.fptr(_abi.f3(pos)),
                                     _abi.fptr(_abi.f3(normal)), _abi.fptr(_abi.f3(anisotropy)), int(method),
                                     section.ctypes.data, C.byref(contact)))
  if return_contact:
    return (section, int(contact.value))
  return section


def slice(
  labels: npt.NDArray[np.integer],
  pos: POINT_T,
  normal: VECTOR_T,
  anisotropy: Optional[VECTOR_T] = None,
  standardize_basis: bool = True,
  crop: float = float("inf"),
) -> npt.NDArray[np.integer]:
  """
  Oblique 2-D resampling of a label volume through `pos` with plane normal `normal`
  (reference: xs3d/__init__.py:166-232).  The orientation of the projection is not guaranteed.
  `labels` may also be a `Volume` holding the labels (Volume(labels=...)): the ndarray form uploads the label volume
  on every call, which is what a single call costs; a resident Volume serves any number of planes without it.

  crop: distance in physical units to limit the slice to.
  """
  assert crop >= 0.0
  if isinstance(labels, Volume):          # resident labels: nothing is uploaded
    return labels.slice(pos, normal, anisotropy, standardize_basis, crop)
  if anisotropy is None:
    anisotropy = (1.0, 1.0, 1.0)
    if labels.ndim == 2:
      anisotropy = (1.0, 1.0)
  pos = np.asarray(pos, dtype=np.float32)
  normal = np.asarray(normal, dtype=np.float32)
  anisotropy = np.asarray(anisotropy, dtype=np.float32)
  if np.all(normal == 0):
    raise ValueError("normal vector must not be a null vector (all zeros).")
  if labels.ndim != 3:
    raise ValueError(f"{labels.ndim} dimensions not supported")
  if not labels.flags.f_contiguous:
    labels = np.ascontiguousarray(labels)
  itemsize = labels.dtype.itemsize
  if itemsize not in (1, 2, 4, 8):
    raise ValueError("labels must have a 1, 2, 4 or 8 byte dtype")
  c_order = 1 if labels.flags.c_contiguous else 0
  L = _abi.lib()
  shape = (C.c_int64 * 2)()
  args = (labels.ctypes.data, labels.shape[0], labels.shape[1], labels.shape[2], itemsize, c_order,
          _abi.fptr(_abi.f3(pos)), _abi.fptr(_abi.f3(normal)), _abi.fptr(_abi.f3(anisotropy)),
          int(standardize_basis), C.c_float(crop))
  _abi.check(L.xs3d_projection(*args, None, shape))
  out = np.zeros((shape[0], shape[1]), dtype=labels.dtype, order="F")
  if out.size:
    _abi.check(L.xs3d_projection(*args, out.ctypes.data, shape))
  return out


def set_shape(image):
  """
  Pre-allocate the GPU-side scratch (default volume + workspaces) for images of `image.shape`, the GPU meaning of
  the reference's pre-allocated visited buffer (xs3d/__init__.py:234-243).  Later calls still evaluate — and
  upload — the image they are given; only the shape has to match for the scratch to be reused.
  """
  sx, sy, sz = (int(s) for s in image.shape)
  _abi.check(_abi.lib().xs3d_set_shape(sx, sy, sz))


def clear_shape():
  """Free the resident image of set_shape."""
  _abi.check(_abi.lib().xs3d_clear_shape())
