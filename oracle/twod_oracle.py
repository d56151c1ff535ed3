"""TEST INFRASTRUCTURE — CPU restatement of the reference's 2-D path (xs3d/twod.py:73-174, reached from
xs3d/__init__.py:79-80 when the image is 2-D).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this; the product path (libxs3d_b200.so: xs_twod.cu) never does.

The reference's algorithm, restated with every dtype decision numpy (>= 2: NEP 50, Python scalars are "weak") makes
spelled out — pos / vec / anisotropy arrive as float32 arrays, so everything derived from them stays float32 until a
Python float list goes through np.unique (float64 from there on):

  1. mark (twod.py:25-71): the line through `pos` perpendicular... rather: with slope = -vec[0] / vec[1] (float32; inf when
     vec[1] == 0) and intercept b = pos[1] - slope * pos[0]; every FOREGROUND pixel (x, y) whose nearest point (p_x, p_y) on
     that line lies inside the pixel ([x-.5, x+.5] x [y-.5, y+.5], bounds included) marks pixel
     (min(rint(p_x), sx-1), min(rint(p_y), sy-1)) — with that pixel's own image value;
  2. component (twod.py:71, 93-94): the reference labels the marks with the third-party cc3d (8-connected) and keeps the
     label of the pixel (int(pos[0]), int(pos[1])).  cc3d is absent here; only "same label as the seed" is ever used,
     so a flood fill from the seed gives the same set.  When the seed pixel itself is unmarked its label is 0 — the
     background label — and the reference then walks ALL unmarked pixels (reproduced);
  3. lengths (twod.py:96-172): nhat = normalize((-vec[1], vec[0])) in float32, slope = nhat[1] / nhat[0] (inf when
     nhat[0] == 0), y-intercept in float32; per selected pixel the intersections of that line with the pixel's four
     sides (float32 arithmetic, bounds included), exact duplicates removed (np.unique on float64 rows); fewer than two
     distinct points: nothing; more than two: ValueError; else the anisotropy-scaled distance, summed in float64 in
     raster order (y outer, x inner).  contact: bit 0 x == 0, 1 x == sx-1, 2 y == 0, 3 y == sy-1 over the selected pixels.

Pinned against the reference's own twod.py (run here with a scipy.ndimage.label stand-in for cc3d) by
tests/golden/make_golden_2d.py -> tests/golden/golden_2d_v1.npz: areas equal bit for bit (float64), contacts equal,
ValueError cases equal.
"""
import numpy as np

f32 = np.float32
EARLY = (0.0, 0b00111111)       # twod.py:83-90


def _marks(binimg, pos, vec):
  """twod.py:25-71 without the labelling: the boolean mark image."""
  sx, sy = binimg.shape
  inf = f32(np.inf)
  with np.errstate(all="ignore"):
    slope = inf if vec[1] == 0 else f32(f32(-vec[0]) / vec[1])
    b = f32(pos[1] - f32(slope * pos[0]))
    marks = np.zeros((sx, sy), bool, order="F")
    for y in range(sy):
      for x in range(sx):
        if not binimg[x, y]:
          continue
        if slope == 0:
          p_x, p_y = f32(x), pos[1]
        elif slope == inf:
          p_x, p_y = pos[0], f32(y)
        else:
          m_n = f32(f32(-1.0) / slope)                          # (m != 0 here)
          b_n = f32(f32(y) - f32(m_n * f32(x)))
          p_x = f32(f32(b - b_n) / f32(m_n - slope))
          p_y = f32(f32(slope * p_x) + b)
        if p_x < x - 0.5 or p_x > x + 0.5 or p_y < y - 0.5 or p_y > y + 0.5:
          continue
        px = min(int(np.rint(p_x)), sx - 1)
        py = min(int(np.rint(p_y)), sy - 1)
        marks[px, py] = binimg[px, py]
  return marks


def _component(marks, seed):
  """8-connected component of `seed` in the marks; all UNMARKED pixels when the seed is unmarked (label 0)."""
  if not marks[seed]:
    return ~marks
  sx, sy = marks.shape
  sel = np.zeros_like(marks)
  sel[seed] = True
  stack = [seed]
  while stack:
    x, y = stack.pop()
    for dy in (-1, 0, 1):
      for dx in (-1, 0, 1):
        u, v = x + dx, y + dy
        if 0 <= u < sx and 0 <= v < sy and marks[u, v] and not sel[u, v]:
          sel[u, v] = True
          stack.append((u, v))
  return sel


def cross_sectional_area_2d(binimg, pos, vec, anisotropy=(1.0, 1.0)):
  """(area float64, contact) as twod.py:73-174 returns them; raises ValueError where the reference does."""
  binimg = np.asarray(binimg)
  pos = np.asarray(pos, f32)
  vec = np.asarray(vec, f32)
  anisotropy = np.asarray(anisotropy, f32)
  sx, sy = binimg.shape
  if pos[0] >= sx or pos[0] < 0 or pos[1] >= sy or pos[1] < 0:
    return EARLY
  seed = (int(pos[0]), int(pos[1]))
  if not binimg[seed]:
    return EARLY
  sel = _component(_marks(binimg, pos, vec), seed)
  inf = f32(np.inf)
  with np.errstate(all="ignore"):
    n0, n1 = f32(-vec[1]), vec[0]
    norm = f32(np.sqrt(f32(f32(n0 * n0) + f32(n1 * n1))))
    n0, n1 = f32(n0 / norm), f32(n1 / norm)
    slope = inf if n0 == 0 else f32(n1 / n0)
    wx, wy = float(anisotropy[0]), float(anisotropy[1])
    yi = f32(pos[1] - f32(slope * pos[0]))
    total, contact = 0.0, 0
    for y in range(sy):
      for x in range(sx):
        if not sel[x, y]:
          continue
        contact |= int(x == 0) | (int(x == sx - 1) << 1) | (int(y == 0) << 2) | (int(y == sy - 1) << 3)
        h0, h1, v0, v1 = f32(y - 0.5), f32(y + 0.5), f32(x - 0.5), f32(x + 0.5)
        pts = []
        if slope == 0:
          if y == pos[1]:
            pts = [(float(v0), float(y)), (float(v1), float(y))]
        elif slope == inf or slope == -inf:
          if x == pos[0]:
            pts = [(float(x), float(h0)), (float(x), float(h1))]
        else:
          x0, x1 = f32(f32(h0 - yi) / slope), f32(f32(h1 - yi) / slope)
          y0, y1 = f32(f32(v0 * slope) + yi), f32(f32(v1 * slope) + yi)
          if h0 <= y0 <= h1:
            pts.append((float(v0), float(y0)))
          if h0 <= y1 <= h1:
            pts.append((float(v1), float(y1)))
          if v0 <= x0 <= v1:
            pts.append((float(x0), float(h0)))
          if v0 <= x1 <= v1:
            pts.append((float(x1), float(h1)))
        uniq = sorted(set(pts))
        if len(uniq) < 2:
          continue
        if len(uniq) > 2:
          raise ValueError(np.array(uniq))
        (ax, ay), (bx, by) = uniq
        px, py = wx * (ax - bx), wy * (ay - by)
        total += float(np.sqrt(px * px + py * py))
  return total, contact
