"""Deterministic synthetic inputs for the BASELINE.json configurations.

These are measurement/test inputs, not part of the reference: sphere (config 0), oblique cylinder
(config 1), branching-tube "neurite" with its skeleton vertices and tangents (configs 2-4) and a
label volume derived from it (config 4, slice).
"""
import numpy as np


def make_sphere(n=500, center=250.0, radius=200.0):
  """bool F-order n^3 ball: (x-c)^2+(y-c)^2+(z-c)^2 <= r^2 in float32."""
  ax = (np.arange(n, dtype=np.float32) - np.float32(center)) ** 2
  d2 = ax[:, None, None] + ax[None, :, None] + ax[None, None, :]
  return np.asfortranarray(d2 <= np.float32(radius * radius))


def make_cylinder(n=512, axis=(1.0, 0.4, 0.25), radius=40.0):
  """bool F-order n^3 oblique cylinder through the volume centre; returns (vol, unit axis float32)."""
  a = np.asarray(axis, dtype=np.float64)
  a = (a / np.linalg.norm(a)).astype(np.float32)
  c = np.float32(n // 2)
  vol = np.zeros((n, n, n), dtype=bool, order="F")
  x = (np.arange(n, dtype=np.float32) - c)
  for z in range(n):
    dx = x[:, None]
    dy = x[None, :]
    dz = np.float32(z) - c
    t = dx * a[0] + dy * a[1] + dz * a[2]
    d2 = (dx - t * a[0]) ** 2 + (dy - t * a[1]) ** 2 + (dz - t * a[2]) ** 2
    vol[:, :, z] = d2 <= np.float32(radius * radius)
  return vol, a


def _raster_capsule(vol, p, q, r):
  n = np.array(vol.shape)
  lo = np.maximum(np.floor(np.minimum(p, q) - r - 1).astype(int), 0)
  hi = np.minimum(np.ceil(np.maximum(p, q) + r + 1).astype(int) + 1, n)
  if np.any(lo >= hi):
    return
  gx = np.arange(lo[0], hi[0], dtype=np.float32)[:, None, None]
  gy = np.arange(lo[1], hi[1], dtype=np.float32)[None, :, None]
  gz = np.arange(lo[2], hi[2], dtype=np.float32)[None, None, :]
  p32 = p.astype(np.float32)
  d = (q - p).astype(np.float32)
  dd = np.float32(d @ d)
  t = ((gx - p32[0]) * d[0] + (gy - p32[1]) * d[1] + (gz - p32[2]) * d[2]) / dd
  t = np.clip(t, np.float32(0), np.float32(1))
  ex = gx - (p32[0] + t * d[0])
  ey = gy - (p32[1] + t * d[1])
  ez = gz - (p32[2] + t * d[2])
  inside = (ex * ex + ey * ey + ez * ez) <= np.float32(r * r)
  vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] |= inside


def make_neurite(n=512, seed=0, n_branches=None, seg_len=24, r_lo=3.0, r_hi=9.0):
  """Branching random-walk tubes.

  Returns (vol bool F-order [n,n,n], verts int32 [M,3], tangents float32 [M,3]).
  Skeleton vertices are the rounded points along every segment that fall on foreground voxels
  inside the volume; the tangent of a vertex is its segment's direction.
  """
  if n_branches is None:
    n_branches = 10 if n <= 512 else 16
  rng = np.random.default_rng(seed)
  vol = np.zeros((n, n, n), dtype=bool, order="F")
  verts, tans = [], []
  starts = [(np.array([n * .1, n * .5, n * .5]), np.array([1., .15, .1]))]
  for b in range(n_branches):
    p, d = starts[rng.integers(len(starts))] if b > 0 else starts[0]
    p = p.copy()
    d = d / np.linalg.norm(d)
    if b > 0:
      d = d + rng.normal(0, .8, 3)
      d /= np.linalg.norm(d)
    r = rng.uniform(r_lo, r_hi)
    for _ in range(int(2.2 * n / seg_len)):
      d = d + rng.normal(0, .25, 3)
      d /= np.linalg.norm(d)
      q = p + d * seg_len
      if np.any(q < -4) or np.any(q > n + 3):
        break
      r = float(np.clip(r + rng.normal(0, .4), r_lo, r_hi))
      _raster_capsule(vol, p, q, r)
      for u in np.linspace(0, 1, seg_len, endpoint=False):
        v = np.rint(p + u * (q - p)).astype(np.int64)
        if np.all(v >= 0) and np.all(v < n):
          verts.append(v)
          tans.append(d.astype(np.float32))
      p = q
      if rng.random() < .15:
        starts.append((p.copy(), d.copy()))
  verts = np.asarray(verts, dtype=np.int32).reshape(-1, 3)
  tans = np.asarray(tans, dtype=np.float32).reshape(-1, 3)
  keep = vol[verts[:, 0], verts[:, 1], verts[:, 2]]
  return vol, verts[keep], tans[keep]


def draw_queries(verts, tans, count=10000, seed=1):
  """(pos float32 [count,3], normal float32 [count,3]) drawn with replacement."""
  idx = np.random.default_rng(seed).choice(len(verts), size=count, replace=True)
  return verts[idx].astype(np.float32), tans[idx].astype(np.float32)


def make_labels(vol, dtype=np.uint64, hashed=False):
  """Label volume for slice(): 2**40+7 inside the tubes (itemsize permitting), 0 outside; the hashed
  variant stores a per-voxel hash of the linear index so gathers can be checked value by value."""
  dt = np.dtype(dtype)
  if hashed:
    idx = np.arange(vol.size, dtype=np.uint64).reshape(vol.shape, order="F")
    h = (idx * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(64 - 8 * dt.itemsize)
    return np.asfortranarray(h.astype(dt))
  val = (2 ** 40 + 7) if dt.itemsize == 8 else (2 ** (8 * dt.itemsize - 1) + 7)
  out = np.zeros(vol.shape, dtype=dt, order="F")
  out[vol] = val
  return out
