"""CPU: the CUDA kernel headers compiled as plain C++ (tests/hostsim) against the oracle.  This covers the
query logic that runs on the GPU — lattice tiles, bitmap walk, first-touch hash, ownership compaction,
ordered sums, slice row-scan — on a machine without a GPU.  Device-only concerns (float contraction,
intrinsics, atomics under real concurrency) are covered by the -m gpu tests."""
import numpy as np
import pytest

import cases
import golden_util as gu
from conftest import bits


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_area_random(hostsim, oracle, mode):
  rng = np.random.default_rng(1234 + mode)
  for it in range(250):
    vol, p, n, w = cases.rand_case(rng, max_side=18)
    a, c = oracle.area(vol, p, n, w)
    a2, c2, _ = hostsim.area_batch(vol, p, n, w, mode)
    assert bits(a) == bits(a2[0]) and c == c2[0], (it, vol.shape, p, n, w)


def test_area_golden(hostsim, golden):
  for i, vol, p, n, w in gu.small_cases(golden):
    for mode in (1, 5):
      a2, c2, _ = hostsim.area_batch(vol, p, n, w, mode)
      assert bits(a2[0]) == bits(golden["small_area"][i]) and c2[0] == golden["small_contact"][i], (i, mode)


def test_area_neurite_tiers(hostsim, oracle):
  """Warp-tier-sized arena with overflow to the big arena (mode 1) and ring-spill walk (mode 2)."""
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(256, seed=0)
  qp, qn = synthetic.draw_queries(verts, tans, 400, seed=3)
  a, c = oracle.area_batch(vol, qp, qn, [32, 32, 40])
  a1, c1, st = hostsim.area_batch(vol, qp, qn, [32, 32, 40], 1)
  assert np.array_equal(bits(a), bits(a1)) and np.array_equal(c, c1)
  assert st[0] > 0, "expected some queries to overflow the small arena"
  a2, c2, _ = hostsim.area_batch(vol, qp[:60], qn[:60], [32, 32, 40], 2)
  assert np.array_equal(bits(a[:60]), bits(a2)) and np.array_equal(c[:60], c2)
  a3, c3, st3 = hostsim.area_batch(vol, qp[:200], qn[:200], [32, 32, 40], 3)   # neighbour-byte (mid tier) flood
  assert np.array_equal(bits(a[:200]), bits(a3)) and np.array_equal(c[:200], c3)
  a4, c4, st4 = hostsim.area_batch(vol, qp[:200], qn[:200], [32, 32, 40], 4)   # ... with 32-bit cells and a 160-cell window (wide tier)
  assert np.array_equal(bits(a[:200]), bits(a4)) and np.array_equal(c[:200], c4)
  a5, c5, st5 = hostsim.area_batch(vol, qp, qn, [32, 32, 40], 5)               # the warp tier's neighbour-byte arena, big arena on overflow
  assert np.array_equal(bits(a), bits(a5)) and np.array_equal(c, c5)
  assert 0 < st5[0] < len(qp) // 2, "some, not most, queries should overflow the warp tier's arena"


def test_area_long_sections_mid_tier(hostsim, oracle):
  """Neighbour-mask flood (mode 3): sections long enough to grow the sampled tile set beyond the initial 3x3
  tiles, bent tubes with side branches touching tile borders, and sections that leave the 256-cell window."""
  rng = np.random.default_rng(77)
  N = 160
  g = np.stack(np.meshgrid(*[np.arange(N, dtype=np.float32)] * 3, indexing="ij"), -1)
  grown = 0
  for it in range(6):
    c0 = rng.uniform(60, 100, size=3).astype(np.float32)
    d = rng.normal(size=3).astype(np.float32); d /= np.linalg.norm(d)
    r = rng.uniform(3, 7)
    rel = g - c0
    t = rel @ d
    dist2 = (rel ** 2).sum(-1) - t ** 2
    vol = dist2 <= r * r
    vol ^= (rng.random(vol.shape) < 0.02) & (dist2 <= (r + 2) ** 2)       # ragged surface, small holes
    vol[int(c0[0]), int(c0[1]), int(c0[2])] = True
    e = np.cross(d, rng.normal(size=3)).astype(np.float32); e /= np.linalg.norm(e)
    qp, qn = [], []
    for k in range(4):
      qp.append(np.floor(c0)); qn.append(e + rng.uniform(0.0, 0.12) * d)   # plane (almost) contains the tube axis
    qp = np.array(qp, np.float32); qn = np.array(qn, np.float32)
    vol = np.asfortranarray(vol)
    a, c = oracle.area_batch(vol, qp, qn, [32, 32, 40])
    a3, c3, st = hostsim.area_batch(vol, qp, qn, [32, 32, 40], 3)
    assert np.array_equal(bits(a), bits(a3)) and np.array_equal(c, c3), it
    a4, c4, st4 = hostsim.area_batch(vol, qp, qn, [32, 32, 40], 4)       # wide-tier-like: 160-cell window, 32-bit cells
    assert np.array_equal(bits(a), bits(a4)) and np.array_equal(c, c4), it
    grown += int(st[1] > 4 * 600)
  assert grown > 0, "expected sections larger than the initial 96x96-cell tile set"


@pytest.mark.parametrize("mode", [0, 1])
def test_slice_golden(hostsim, golden, mode):
  for i, lab, p, n, w, sb, crop, exp, ub in gu.slice_cases(golden):
    got = hostsim.projection(lab, p, n, w, sb, crop, mode)
    assert got.shape == exp.shape and got.dtype == exp.dtype, i
    if not ub:
      assert gu.same_bits(got, exp), i


def test_slice_random_vs_oracle(hostsim, oracle):
  rng = np.random.default_rng(8)
  for it in range(300):
    lab, p, n, w, sb, crop = cases.rand_labels_case(rng)
    o = oracle.projection(lab, p, n, w, sb, crop)
    for mode in (0, 1):
      h = hostsim.projection(lab, p, n, w, sb, crop, mode)
      assert gu.same_bits(o, h), (it, mode)


def test_round_half_away(hostsim):
  xs = np.concatenate([np.array([0.5, -0.5, 1.5, -1.5, 2.5, 0.49999997, -0.49999997, 8388607.5, -8388607.5, 1e9, -0.0, 0.0], np.float32),
                       np.random.default_rng(0).normal(scale=100, size=2000).astype(np.float32)])
  for x in xs:
    exp = np.float32(np.sign(x) * np.floor(np.abs(np.float64(x)) + 0.5))
    assert hostsim.lib.hostsim_round(float(x)) == exp, x


def test_polygon_prefilter_never_drops_a_cut_cube(hostsim):
  """plane_cube_miss (the polygon kernel's pre-filter) may only drop cubes for which plane_cube_points finds no point:
  planes through / next to corners, edges and faces, zero and tiny normal components, coordinates up to 60 000."""
  for span, seed in ((64.0, 1), (1100.0, 2), (60000.0, 3)):
    r = hostsim.miss_check(2_000_000, seed, span)
    assert r["violations"] == 0, r
    assert r["miss"] > r["far"] > 0 and r["polygons"] > 0, r   # it filters more than the sphere test, and cut cubes are present
