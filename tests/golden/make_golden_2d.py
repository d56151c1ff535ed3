"""Generate tests/golden/golden_2d_v1.npz from the UNMODIFIED reference 2-D path (dev container only).

  python tests/golden/make_golden_2d.py

Imports /root/reference/xs3d/twod.py as it is.  Its one dependency that is absent here, the third-party `cc3d`, is stood
in for by scipy.ndimage.label with the 8-connected structure: twod.py only ever compares labels with the label of the
seed pixel (twod.py:93-94, 112), so any labelling with background 0 gives the reference's result.  Inputs are stored
(bit-packed images) with the outputs: area (float64), contact, and a flag for the cases where the reference raises
ValueError (more than two distinct intersection points in a pixel).  numpy >= 2 (NEP 50 scalar promotion) is assumed —
under numpy 1.x the reference itself computes some of these values in float64 instead of float32.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import scipy.ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
assert int(np.__version__.split(".")[0]) >= 2, np.__version__

cc3d = types.ModuleType("cc3d")
cc3d.connected_components = lambda img, connectivity=8: ndi.label(np.asarray(img), structure=np.ones((3, 3), int))[0]
sys.modules["cc3d"] = cc3d
pkg = types.ModuleType("xs3d")
pkg.__path__ = ["/root/reference/xs3d"]
sys.modules["xs3d"] = pkg
for name in ("typing", "twod"):
  spec = importlib.util.spec_from_file_location("xs3d." + name, "/root/reference/xs3d/%s.py" % name)
  mod = importlib.util.module_from_spec(spec)
  sys.modules["xs3d." + name] = mod
  spec.loader.exec_module(mod)
twod = sys.modules["xs3d.twod"]
assert twod.__file__.startswith("/root/reference"), twod.__file__


def rand_case_2d(rng, i):
  sx, sy = int(rng.integers(1, 28)), int(rng.integers(1, 28))
  kind = i % 6
  if kind == 0:
    img = np.ones((sx, sy), bool)
  elif kind == 1:
    img = rng.random((sx, sy)) < 0.8
  elif kind == 2:
    img = rng.random((sx, sy)) < 0.45
  else:
    g = np.stack(np.meshgrid(np.arange(sx), np.arange(sy), indexing="ij"), -1).astype(np.float32)
    c = rng.uniform(0, [sx, sy]).astype(np.float32)
    img = ((g - c) ** 2).sum(-1) <= rng.uniform(1, max(sx, sy)) ** 2
    img ^= rng.random((sx, sy)) < 0.05
  fg = np.argwhere(img)
  if len(fg) and rng.random() < 0.9:
    p = fg[rng.integers(len(fg))].astype(np.float32)
    if rng.random() < 0.5:
      p = p + rng.uniform(0, 0.999, 2).astype(np.float32)      # int(pos) is the seed pixel: fractions keep it
  else:
    p = rng.uniform(-2, [sx + 2, sy + 2]).astype(np.float32)
  v = rng.normal(size=2).astype(np.float32)
  s = int(rng.integers(8))
  if s == 0: v[0] = 0
  if s == 1: v[1] = 0
  if s == 2: v = np.array([1, 1], np.float32) * np.float32(rng.choice([-1, 1]))
  if s == 3: v = np.array([1, -1], np.float32)
  if s == 4: v = np.round(v * 3)
  if not v.any(): v = np.array([0, 1], np.float32)
  w = [(1, 1), (3, 3), (4, 40), (0.5, 2.0)][int(rng.integers(4))]
  return np.asfortranarray(img), p, v, np.array(w, np.float32)


if __name__ == "__main__":
  rng = np.random.default_rng(4242)
  N = 600
  shapes, packed, offs, pos, vec, ani, area, contact, err = [], [], [0], [], [], [], [], [], []
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for i in range(N):
      img, p, v, w = rand_case_2d(rng, i)
      try:
        a, c = twod.cross_sectional_area_2d(img, p, v, w)
        e = 0
      except ValueError:
        a, c, e = 0.0, 0, 1
      shapes.append(img.shape)
      bits = np.packbits(img.ravel(order="F"))
      packed.append(bits); offs.append(offs[-1] + len(bits))
      pos.append(p); vec.append(v); ani.append(w); area.append(float(a)); contact.append(int(c)); err.append(e)
  out = os.path.join(ROOT, "tests", "golden", "golden_2d_v1.npz")
  np.savez_compressed(out, shapes=np.array(shapes, np.int32), packed=np.concatenate(packed), off=np.array(offs, np.int64),
                      pos=np.array(pos, np.float32), vec=np.array(vec, np.float32), ani=np.array(ani, np.float32),
                      area=np.array(area, np.float64), contact=np.array(contact, np.uint8), err=np.array(err, np.uint8))
  print("wrote", out, "cases", N, "errors", int(np.sum(err)), "nonzero areas", int(np.sum(np.array(area) > 0)), "early-outs", int(np.sum(np.array(contact) == 63)))
