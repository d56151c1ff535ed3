"""Helpers to iterate the committed golden vectors (tests/golden/golden_v1.npz)."""
import numpy as np

import cases


def small_cases(g):
  """Yield (i, vol, pos, normal, anisotropy) for the stored small volumes."""
  n = len(g["small_shapes"])
  for i in range(n):
    shape = tuple(int(s) for s in g["small_shapes"][i])
    packed = g["small_packed"][g["small_off"][i]:g["small_off"][i + 1]]
    vol = np.unpackbits(packed)[: int(np.prod(shape))].astype(bool).reshape(shape, order="F")
    yield i, np.asfortranarray(vol), g["small_pos"][i], g["small_nrm"][i], g["small_ani"][i]


def section_golden(g, m, i, shape):
  idx = g[f"sec{m}_idx"][g[f"sec{m}_off"][i]:g[f"sec{m}_off"][i + 1]]
  val = g[f"sec{m}_val"][g[f"sec{m}_off"][i]:g[f"sec{m}_off"][i + 1]]
  if m == 2:
    shape = tuple((s + 1) // 2 for s in shape)
  img = np.zeros(int(np.prod(shape)), np.float32)
  img[idx] = val
  return img.reshape(shape, order="F"), int(g[f"sec{m}_contact"][i])


def slice_cases(g):
  """Yield (i, labels, pos, normal, anisotropy, standardize_basis, crop, expected array, ub flag)."""
  rng = np.random.default_rng(77)
  n = len(g["slice_shapes"])
  for i in range(n):
    lab, p, nrm, w, sb, crop = cases.rand_labels_case(rng, max_side=16)
    shape = tuple(int(s) for s in g["slice_shapes"][i])
    raw = g["slice_bytes"][g["slice_off"][i]:g["slice_off"][i + 1]]
    exp = np.frombuffer(raw.tobytes(), dtype=lab.dtype).reshape(shape, order="F") if raw.size else np.zeros(shape, lab.dtype, order="F")
    yield i, lab, p, nrm, w, sb, crop, exp, bool(g["slice_ub"][i])


def same_bits(a, b):
  a = np.asarray(a)
  b = np.asarray(b)
  return a.shape == b.shape and a.dtype == b.dtype and a.tobytes(order="F") == b.tobytes(order="F")


def twod_cases(path=None):
  """Yield (i, image, pos, vec, anisotropy, area float64, contact, raises) from tests/golden/golden_2d_v1.npz
  (outputs of the reference's own xs3d/twod.py, tests/golden/make_golden_2d.py)."""
  import os
  path = path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_2d_v1.npz")
  g = np.load(path)
  for i in range(len(g["area"])):
    sx, sy = (int(v) for v in g["shapes"][i])
    img = np.unpackbits(g["packed"][g["off"][i]:g["off"][i + 1]])[: sx * sy].astype(bool).reshape((sx, sy), order="F")
    yield i, np.asfortranarray(img), g["pos"][i], g["vec"][i], g["ani"][i], float(g["area"][i]), int(g["contact"][i]), bool(g["err"][i])
