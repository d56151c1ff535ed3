"""ctypes binding of libxs3d_b200.so (include/xs3d_b200.h) — the only way the host layer reaches the GPU.

There is deliberately no CPU fallback: if the shared library is missing or a call fails, an exception
is raised (XS3DError).  This module plays the role of `import fastxs3d` in the reference
(/root/reference/xs3d/__init__.py:5).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XS3D_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libxs3d_b200.so")   # XS3D_LIB: dev aid (A/B builds)

HOST, DEVICE = 0, 1

# every symbol include/xs3d_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
  "xs3d_last_error", "xs3d_version", "xs3d_device_count",
  "xs3d_volume_create", "xs3d_volume_destroy", "xs3d_volume_load", "xs3d_pack_bits", "xs3d_volume_load_packed", "xs3d_volume_stage_packed", "xs3d_volume_pack_and_stage", "xs3d_volume_commit_staged", "xs3d_volume_staged_cubes", "xs3d_volume_wait_staged", "xs3d_volume_mark_staged", "xs3d_volume_staged_bits", "xs3d_volume_pack_and_stage_part", "xs3d_volume_wait_staged_bits", "xs3d_volume_ingest_staged", "xs3d_volume_load_labels",
  "xs3d_volume_select_label", "xs3d_volume_cubes", "xs3d_volume_mark_loaded", "xs3d_volume_shape",
  "xs3d_area_batch", "xs3d_area_sweep_host", "xs3d_volume_stats", "xs3d_volume_profile", "xs3d_volume_timing", "xs3d_volume_set_stream", "xs3d_volume_set_after_enqueue",
  "xs3d_area", "xs3d_area_2d", "xs3d_section", "xs3d_section_sparse", "xs3d_projection", "xs3d_projection_batch", "xs3d_projection_plan", "xs3d_projection_fill",
  "xs3d_volume_labels_buffer", "xs3d_skeleton_tangents",
  "xs3d_set_shape", "xs3d_set_shape_image", "xs3d_clear_shape",
]


class XS3DError(RuntimeError):
  pass


_lib = None


def lib():
  """Load libxs3d_b200.so (once).  Fails loudly when the CUDA extension has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise XS3DError(
      f"{LIB_PATH} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
      "There is no CPU fallback."
    )
  L = C.CDLL(LIB_PATH)
  u64, fp, u8p, vp = C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.c_void_p
  L.xs3d_last_error.restype = C.c_char_p
  L.xs3d_version.restype = C.c_int
  L.xs3d_device_count.argtypes = [C.POINTER(C.c_int)]
  L.xs3d_volume_create.argtypes = [C.c_int, u64, u64, u64, C.POINTER(vp)]
  L.xs3d_volume_destroy.argtypes = [vp]
  L.xs3d_volume_load.argtypes = [vp, vp, C.c_int, C.c_int]
  L.xs3d_pack_bits.argtypes = [vp, u64, vp]
  L.xs3d_volume_load_packed.argtypes = [vp, vp, C.c_int, C.c_int]
  L.xs3d_volume_stage_packed.argtypes = [vp, vp, C.c_int, C.c_int]
  L.xs3d_volume_pack_and_stage.argtypes = [vp, vp, C.c_int, C.c_int]
  L.xs3d_volume_commit_staged.argtypes = [vp, C.c_int]
  L.xs3d_volume_staged_cubes.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(u64)]
  L.xs3d_volume_wait_staged.argtypes = [vp, C.c_int, vp]
  L.xs3d_volume_mark_staged.argtypes = [vp, C.c_int, vp]
  L.xs3d_area_2d.argtypes = [vp, u64, u64, fp, fp, fp, C.POINTER(C.c_double), u8p]
  L.xs3d_volume_staged_bits.argtypes = [vp, C.c_int, u64, C.POINTER(vp), C.POINTER(u64)]
  L.xs3d_volume_pack_and_stage_part.argtypes = [vp, vp, C.c_int, u64, u64]
  L.xs3d_volume_wait_staged_bits.argtypes = [vp, C.c_int, vp]
  L.xs3d_volume_ingest_staged.argtypes = [vp, C.c_int, C.c_int, vp]
  L.xs3d_volume_load_labels.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
  L.xs3d_volume_select_label.argtypes = [vp, u64]
  L.xs3d_volume_cubes.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
  L.xs3d_volume_mark_loaded.argtypes = [vp]
  L.xs3d_volume_shape.argtypes = [vp, C.POINTER(u64)]
  L.xs3d_area_batch.argtypes = [vp, u64, vp, vp, fp, C.c_int, vp, vp]
  L.xs3d_area_sweep_host.argtypes = [vp, vp, C.c_int, u64, vp, vp, fp, vp, vp]
  L.xs3d_volume_stats.argtypes = [vp, C.POINTER(u64)]
  L.xs3d_volume_profile.argtypes = [vp, C.POINTER(u64)]
  L.xs3d_volume_timing.argtypes = [vp, fp]
  L.xs3d_volume_set_stream.argtypes = [vp, vp]
  L.xs3d_volume_set_after_enqueue.argtypes = [vp, vp, vp]
  L.xs3d_area.argtypes = [vp, u64, u64, u64, fp, fp, fp, C.c_int, C.c_int, fp, u8p]
  L.xs3d_section.argtypes = [vp, u64, u64, u64, fp, fp, fp, C.c_int, vp, u8p]
  L.xs3d_section_sparse.argtypes = [vp, fp, fp, fp, C.c_int, vp, vp, u64, C.POINTER(u64), u8p]
  L.xs3d_projection.argtypes = [vp, u64, u64, u64, C.c_int, C.c_int, fp, fp, fp, C.c_int, C.c_float, vp, C.POINTER(C.c_int64)]
  L.xs3d_projection_batch.argtypes = [vp, u64, vp, vp, fp, C.c_int, C.c_float, vp, u64, vp, vp, C.c_int]
  L.xs3d_projection_plan.argtypes = [vp, u64, vp, vp, fp, C.c_int, C.c_float, C.c_int, vp, vp]
  L.xs3d_projection_fill.argtypes = [vp, u64, u64, vp, C.c_int]
  L.xs3d_volume_labels_buffer.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(u64)]
  L.xs3d_skeleton_tangents.argtypes = [vp, u64, vp, u64, vp, C.c_int, vp]
  L.xs3d_set_shape.argtypes = [u64, u64, u64]
  L.xs3d_set_shape_image.argtypes = [vp, u64, u64, u64, C.c_int]
  for name in SYMBOLS:
    if name not in ("xs3d_last_error",):
      getattr(L, name).restype = C.c_char_p if name == "xs3d_last_error" else C.c_int
  _lib = L
  return L


def check(rc):
  if rc != 0:
    msg = lib().xs3d_last_error()
    raise XS3DError(f"libxs3d_b200 error {rc}: {msg.decode() if msg else '?'}")


def f3(a):
  a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
  if a.size != 3:
    raise ValueError("expected a 3-vector")
  return a


def fptr(a):
  return a.ctypes.data_as(C.POINTER(C.c_float))


def device_count():
  n = C.c_int(0)
  check(lib().xs3d_device_count(C.byref(n)))
  return n.value
