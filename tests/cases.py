"""Seeded input generators shared by the golden-vector script and the tests."""
import numpy as np

ANISOTROPIES = [[1, 1, 1], [32, 32, 40], [4, 4, 40], [0.5, 2, 1.25]]


def rand_normal(rng):
  k = rng.integers(6)
  if k == 0:
    n = np.eye(3)[rng.integers(3)] * rng.choice([-1, 1])
  elif k == 1:
    n = rng.normal(size=3)
    n[rng.integers(3)] = 0
  elif k == 2:
    n = rng.choice([-1, 0, 1], size=3).astype(float)
  else:
    n = rng.normal(size=3)
  if not np.any(n):
    n = np.array([0, 0, 1.0])
  return n.astype(np.float32)


def rand_volume(rng, max_side=16):
  shape = rng.integers(1, max_side + 1, size=3)
  kind = rng.integers(4)
  if kind == 0:
    vol = rng.random(shape) < rng.uniform(0.2, 0.95)
  elif kind == 1:
    vol = np.ones(shape, bool)
  else:
    g = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).astype(np.float32)
    c = rng.uniform(0, shape, size=3)
    r = rng.uniform(1, max(shape))
    vol = ((g - c) ** 2).sum(-1) <= r * r
    if kind == 3:
      vol ^= (rng.random(shape) < 0.1)
  return np.asfortranarray(vol)


def rand_case(rng, max_side=16):
  """(vol, pos, normal, anisotropy) — volume with at least one foreground voxel."""
  while True:
    vol = rand_volume(rng, max_side)
    fg = np.argwhere(vol)
    if len(fg):
      break
  p = fg[rng.integers(len(fg))].astype(np.float32)
  u = rng.random()
  if u < 0.25:
    p = p + rng.uniform(-0.45, 0.45, size=3).astype(np.float32)
  elif u < 0.3:
    p = rng.uniform(-2, np.array(vol.shape) + 2).astype(np.float32)
  w = ANISOTROPIES[rng.integers(len(ANISOTROPIES))]
  return vol, p, rand_normal(rng), np.asarray(w, np.float32)


def rand_labels_case(rng, max_side=24):
  shape = tuple(int(s) for s in rng.integers(1, max_side + 1, size=3))
  dt = [np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.int64][rng.integers(6)]
  lab = rng.integers(1, 250, size=shape).astype(dt)
  if rng.random() < 0.5:
    lab = np.asfortranarray(lab)
  p = rng.uniform(0, shape).astype(np.float32)
  if rng.random() < 0.5:
    p = np.floor(p)
  if rng.random() < 0.05:
    p = rng.uniform(-2, np.array(shape) + 2).astype(np.float32)
  n = rand_normal(rng)
  w = np.asarray(ANISOTROPIES[rng.integers(len(ANISOTROPIES))], np.float32)
  crop = [float("inf"), 0.0, 3.0, 10.0, 100.0, 17.5][rng.integers(6)]
  sb = bool(rng.integers(2))
  return lab, p, n, w, sb, crop
