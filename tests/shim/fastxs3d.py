"""fastxs3d.py — ctypes stand-in for the reference's pybind11 module `fastxs3d` (/root/reference/src/fastxs3d.cpp:218-226)
over libxs3d_b200.so.  Same five names and call signatures, so the reference's own Python package
(/root/reference/xs3d/__init__.py:5 `import fastxs3d`) and its own test file run unchanged on the CUDA path.
This is the binding INTEGRATION.md §2 describes, as a file; tests/test_reference_suite.py puts it on the path.
Test infrastructure: the product's host layer is cross-section_b200/xs3d."""
import ctypes as C
import os

import numpy as np

_LIB = os.environ.get("XS3D_B200_LIB") or os.path.join(
  os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "cross-section_b200", "lib", "libxs3d_b200.so")
_L = C.CDLL(_LIB)
_L.xs3d_last_error.restype = C.c_char_p
_u64 = C.c_uint64


def _f3(point, normal, anisotropy):
  arrs = [np.ascontiguousarray(x, np.float32) for x in (point, normal, anisotropy)]
  return arrs, [a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs]


def _chk(rc):
  if rc:
    raise RuntimeError((_L.xs3d_last_error() or b"?").decode())


def area(binimg, point, normal, anisotropy, slow_method, use_persistent_data):
  a, c = C.c_float(), C.c_uint8()
  sx, sy, sz = binimg.shape
  keep, (p, n, w) = _f3(point, normal, anisotropy)
  _chk(_L.xs3d_area(C.c_void_p(binimg.ctypes.data), _u64(sx), _u64(sy), _u64(sz), p, n, w,
                    int(slow_method), int(use_persistent_data), C.byref(a), C.byref(c)))
  return a.value, c.value


def section(binimg, point, normal, anisotropy, method=0):
  sx, sy, sz = binimg.shape
  shape = (sx, sy, sz) if method < 2 else tuple((s + 1) // 2 for s in (sx, sy, sz))
  out, c = np.zeros(shape, np.float32, order="F"), C.c_uint8()
  keep, (p, n, w) = _f3(point, normal, anisotropy)
  _chk(_L.xs3d_section(C.c_void_p(binimg.ctypes.data), _u64(sx), _u64(sy), _u64(sz), p, n, w, int(method),
                       C.c_void_p(out.ctypes.data), C.byref(c)))
  return out, c.value


def projection(labels, point, normal, anisotropy, standardize_basis, crop_distance):
  sx, sy, sz = labels.shape
  shp = (C.c_int64 * 2)()
  keep, (p, n, w) = _f3(point, normal, anisotropy)
  args = (C.c_void_p(labels.ctypes.data), _u64(sx), _u64(sy), _u64(sz), labels.dtype.itemsize,
          int(labels.flags.c_contiguous), p, n, w, int(standardize_basis), C.c_float(crop_distance))
  _chk(_L.xs3d_projection(*args, None, shp))
  out = np.zeros((shp[0], shp[1]), labels.dtype, order="F")
  if out.size:
    _chk(_L.xs3d_projection(*args, C.c_void_p(out.ctypes.data), shp))
  return out


def set_shape(sx, sy, sz):
  _chk(_L.xs3d_set_shape(_u64(sx), _u64(sy), _u64(sz)))


def clear_shape():
  _chk(_L.xs3d_clear_shape())
