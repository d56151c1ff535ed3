"""ctypes bindings for the CHECKERS under oracle/ — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (cross-section_b200/) never does.

  Ref      -> oracle/_ref/libxs3d_ref.so : the unmodified reference behind oracle/ref_shim.cpp
  Oracle   -> oracle/_build/libxs3d_oracle.so : the plain-C restatement oracle/xs3d_oracle.c
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libxs3d_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libxs3d_oracle.so")

_f3 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(3))
_fp = C.POINTER(C.c_float)
_u8p = C.POINTER(C.c_uint8)


def _ptr(a, t):
  return a.ctypes.data_as(t)


def _as_u8_f(binimg):
  """F-order uint8 view of a bool/uint8 3-D array (what xs3d/__init__.py:77-83 hands to C++)."""
  a = np.asfortranarray(binimg)
  if a.dtype == bool:
    a = a.view(np.uint8)
  assert a.dtype == np.uint8 and a.ndim == 3
  return a


def normalize_f32(n):
  """normal /= normal.norm() exactly as xs3d.hpp:1347-1349 (float32, sqrt of left-to-right sum)."""
  n = np.asarray(n, dtype=np.float32)
  s = np.float32(np.float32(n[0] * n[0]) + np.float32(n[1] * n[1]))
  s = np.float32(s + np.float32(n[2] * n[2]))
  return (n / np.sqrt(s, dtype=np.float32)).astype(np.float32)


class _AreaAPI:
  """Common surface of the two CPU checkers: same names/arguments for Ref and Oracle."""
  prefix = ""

  def __init__(self, path):
    if not os.path.exists(path):
      raise FileNotFoundError(f"{path} missing — run `make -C oracle` (or __graft_entry__.build())")
    self.lib = lib = C.CDLL(path)
    p = self.prefix
    u64 = C.c_uint64
    self._area = getattr(lib, p + "area")
    self._area.argtypes = [_u8p, u64, u64, u64, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _u8p]
    self._area.restype = None
    self._batch = getattr(lib, p + "area_batch")
    self._batch.argtypes = [_u8p, u64, u64, u64, u64, _fp, _fp, _fp, _fp, _u8p, C.c_int]
    self._batch.restype = None
    self._section = getattr(lib, p + "section")
    self._section.argtypes = [_u8p, u64, u64, u64, _fp, _fp, _fp, C.c_int, _fp, _u8p]
    self._section.restype = None
    self._projection = getattr(lib, p + "projection")
    self._projection.argtypes = [C.c_void_p, u64, u64, u64, C.c_int, C.c_int, _fp, _fp, _fp,
                                 C.c_int, C.c_float, C.c_void_p, C.POINTER(C.c_int64)]
    self._projection.restype = C.c_int

  # -- fastxs3d.area (fastxs3d.cpp:82)
  def area(self, binimg, pos, normal, anisotropy=(1, 1, 1), slow_method=False, use_persistent_data=False):
    a = _as_u8_f(binimg)
    out = C.c_float(0)
    ct = C.c_uint8(0)
    p, n, w = _f3(pos), _f3(normal), _f3(anisotropy)
    sx, sy, sz = a.shape
    self._area(_ptr(a, _u8p), sx, sy, sz, _ptr(p, _fp), _ptr(n, _fp), _ptr(w, _fp),
               int(slow_method), int(use_persistent_data), C.byref(out), C.byref(ct))
    return np.float32(out.value), int(ct.value)

  def area_batch(self, binimg, pos, normal, anisotropy=(1, 1, 1), nthreads=1):
    a = _as_u8_f(binimg)
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normal, dtype=np.float32).reshape(-1, 3)
    w = _f3(anisotropy)
    nq = pos.shape[0]
    area = np.zeros(nq, dtype=np.float32)
    contact = np.zeros(nq, dtype=np.uint8)
    sx, sy, sz = a.shape
    self._batch(_ptr(a, _u8p), sx, sy, sz, nq, _ptr(pos, _fp), _ptr(nrm, _fp), _ptr(w, _fp),
                _ptr(area, _fp), _ptr(contact, _u8p), int(nthreads))
    return area, contact

  # -- fastxs3d.section (fastxs3d.cpp:13)
  def section(self, binimg, pos, normal, anisotropy=(1, 1, 1), method=0):
    a = _as_u8_f(binimg)
    sx, sy, sz = a.shape
    shape = (sx, sy, sz) if method < 2 else ((sx + 1) // 2, (sy + 1) // 2, (sz + 1) // 2)
    out = np.zeros(shape, dtype=np.float32, order="F")
    ct = C.c_uint8(0)
    p, n, w = _f3(pos), _f3(normal), _f3(anisotropy)
    self._section(_ptr(a, _u8p), sx, sy, sz, _ptr(p, _fp), _ptr(n, _fp), _ptr(w, _fp),
                  int(method), _ptr(out, _fp), C.byref(ct))
    return out, int(ct.value)

  # -- fastxs3d.projection (fastxs3d.cpp:119): plane sizing + bbox(+1) cut-out restated from :137-196
  @staticmethod
  def plane_size(shape, anisotropy, crop):
    w = np.asarray(anisotropy, dtype=np.float32)
    w = (w / w.min()).astype(np.float32)
    distortion = int(np.ceil(np.abs(w).max()))
    largest = int(max(shape))
    if np.float32(largest) > np.float32(crop) and crop >= 0:
      largest = int(np.ceil(np.float32(crop)))
    return (distortion * 2 * 97 * largest) // 56 + 1

  def projection(self, labels, pos, normal, anisotropy=(1, 1, 1), standardize_basis=True, crop=float("inf")):
    if not labels.flags.f_contiguous:
      labels = np.ascontiguousarray(labels)
    c_order = bool(labels.flags.c_contiguous)   # fastxs3d.cpp:135 (c_style flag wins for 1-wide arrays)
    itemsize = labels.dtype.itemsize
    udt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[itemsize]
    if crop == 0:
      return np.zeros((0, 0), dtype=labels.dtype, order="F")
    psx = self.plane_size(labels.shape, anisotropy, crop)
    out = np.zeros((psx, psx), dtype=udt, order="F")
    bbox = (C.c_int64 * 4)()
    p, n, w = _f3(pos), _f3(normal), _f3(anisotropy)
    sx, sy, sz = labels.shape
    rc = self._projection(labels.ctypes.data, sx, sy, sz, itemsize, int(c_order),
                          _ptr(p, _fp), _ptr(n, _fp), _ptr(w, _fp), int(standardize_basis),
                          C.c_float(crop), out.ctypes.data, bbox)
    assert rc == 0
    x0, x1, y0, y1 = bbox[0], bbox[1] + 1, bbox[2], bbox[3] + 1
    return np.asfortranarray(out[x0:x1, y0:y1]).view(labels.dtype)


class Ref(_AreaAPI):
  """The unmodified reference (oracle/_ref)."""
  prefix = "ref_"

  def __init__(self, path=REF_SO):
    super().__init__(path)
    lib = self.lib
    u64 = C.c_uint64
    lib.ref_set_shape.argtypes = [u64, u64, u64]
    lib.ref_check_1x1x1.argtypes = [u64, u64, u64, _fp, _fp, _fp]
    lib.ref_check_1x1x1.restype = C.c_int
    lib.ref_check_2x2x2.argtypes = [u64, u64, u64, _fp, _fp, _fp, C.POINTER(C.c_uint16)]
    lib.ref_check_2x2x2.restype = C.c_int
    lib.ref_points_to_area.argtypes = [_fp, C.c_int, _fp, _fp]
    lib.ref_points_to_area.restype = C.c_float
    lib.ref_voxel_area.argtypes = [u64, u64, u64, _fp, _fp, _fp]
    lib.ref_voxel_area.restype = C.c_float
    lib.ref_block_area.argtypes = [C.c_uint8, _fp, _fp, _fp, _fp]
    lib.ref_block_area.restype = C.c_float
    lib.ref_is_26_connected.argtypes = [C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.c_int]
    lib.ref_is_26_connected.restype = C.c_int
    lib.ref_compute_cube.argtypes = [_u8p, u64, u64, u64, u64, u64, u64]
    lib.ref_compute_cube.restype = C.c_uint8
    lib.ref_slice_basis.argtypes = [_fp, C.c_int, _fp, _fp]

  def set_shape(self, shape):
    self.lib.ref_set_shape(*[int(s) for s in shape])

  def clear_shape(self):
    self.lib.ref_clear_shape()

  def voxel_points(self, xyz, pos, normal, block=False):
    p, n = _f3(pos), _f3(normal)
    pts = np.zeros((12, 3), dtype=np.float32)
    if block:
      e = C.c_uint16(0)
      k = self.lib.ref_check_2x2x2(*[int(v) for v in xyz], _ptr(p, _fp), _ptr(n, _fp), _ptr(pts, _fp), C.byref(e))
    else:
      k = self.lib.ref_check_1x1x1(*[int(v) for v in xyz], _ptr(p, _fp), _ptr(n, _fp), _ptr(pts, _fp))
    return pts[:k].copy()

  def points_to_area(self, pts, anisotropy, normal):
    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
    w, n = _f3(anisotropy), _f3(normal)
    return np.float32(self.lib.ref_points_to_area(_ptr(pts, _fp), pts.shape[0], _ptr(w, _fp), _ptr(n, _fp)))

  def voxel_area(self, xyz, pos, normal, anisotropy):
    p, n, w = _f3(pos), _f3(normal), _f3(anisotropy)
    return np.float32(self.lib.ref_voxel_area(*[int(v) for v in xyz], _ptr(p, _fp), _ptr(n, _fp), _ptr(w, _fp)))

  def block_area(self, cube, cur, pos, normal, anisotropy):
    c, p, n, w = _f3(cur), _f3(pos), _f3(normal), _f3(anisotropy)
    return np.float32(self.lib.ref_block_area(int(cube), _ptr(c, _fp), _ptr(p, _fp), _ptr(n, _fp), _ptr(w, _fp)))

  def is_26_connected(self, center, candidate, o):
    return bool(self.lib.ref_is_26_connected(int(center), int(candidate), int(o[0]), int(o[1]), int(o[2])))

  def slice_basis(self, normal, positive=True):
    n = _f3(normal)
    b1 = np.zeros(3, np.float32)
    b2 = np.zeros(3, np.float32)
    self.lib.ref_slice_basis(_ptr(n, _fp), int(positive), _ptr(b1, _fp), _ptr(b2, _fp))
    return b1, b2


class Oracle(_AreaAPI):
  """The plain-C restatement (oracle/xs3d_oracle.c)."""
  prefix = "oracle_"

  def __init__(self, path=ORACLE_SO):
    super().__init__(path)


def have_ref():
  return os.path.exists(REF_SO)


def have_oracle():
  return os.path.exists(ORACLE_SO)
