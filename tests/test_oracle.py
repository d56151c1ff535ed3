"""CPU: the plain-C oracle against (1) the committed golden vectors produced by the unmodified reference,
(2) the reference's own known-answer tests (automated_test.py), (3) the compiled reference itself on
fresh randomised inputs when oracle/_ref is present."""
import numpy as np
import pytest

import cases
import golden_util as gu
from conftest import bits


def test_golden_small_area(oracle, golden):
  for i, vol, p, n, w in gu.small_cases(golden):
    a, c = oracle.area(vol, p, n, w)
    assert bits(a) == bits(golden["small_area"][i]) and c == golden["small_contact"][i], i
    a, c = oracle.area(vol, p, n, w, slow_method=True)
    assert bits(a) == bits(golden["small_slow_area"][i]) and c == golden["small_slow_contact"][i], i


def test_golden_small_sections(oracle, golden):
  for i, vol, p, n, w in gu.small_cases(golden):
    for m in (0, 1, 2):
      img, c = oracle.section(vol, p, n, w, method=m)
      exp, ec = gu.section_golden(golden, m, i, vol.shape)
      assert gu.same_bits(img, np.asfortranarray(exp)) and c == ec, (i, m)


def test_golden_slices(oracle, golden):
  for i, lab, p, n, w, sb, crop, exp, ub in gu.slice_cases(golden):
    got = oracle.projection(lab, p, n, w, sb, crop)
    assert got.shape == exp.shape and got.dtype == exp.dtype, i
    if not ub:   # the reference's own output is garbage where it casts -1 to uint64
      assert gu.same_bits(got, exp), i


def test_golden_neurite(oracle, golden):
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(256, seed=0)
  assert vol.sum() == golden["neurite256_fg"] and len(verts) == golden["neurite256_nverts"]
  qp, qn = synthetic.draw_queries(verts, tans, 1000, seed=1)
  a, c = oracle.area_batch(vol, qp, qn, [32, 32, 40])
  assert np.array_equal(bits(a), bits(golden["neurite256_area"]))
  assert np.array_equal(c, golden["neurite256_contact"])


def test_golden_cylinder(oracle, golden):
  from xs3d import synthetic
  cyl, axis = synthetic.make_cylinder()
  assert np.array_equal(axis, golden["cyl_axis"])
  for k in range(3):
    a, c = oracle.area(cyl, golden["cyl_verts"][k], axis, [32, 32, 40])
    assert bits(a) == bits(golden["cyl_area"][k]) and c == golden["cyl_contact"][k]


# ---- the reference's own known-answer tests, restated on the oracle (automated_test.py) ------------------------
def area(oracle, img, p, n, w=(1, 1, 1)):
  return float(oracle.area(np.asfortranarray(img), p, n, w)[0])


@pytest.mark.parametrize("aniso", [[1, 1, 1], [2, 2, 2], [1000, 1000, 1000], [0.0001, 0.0001, 0.0001], [1, 1, 0.001], [1, 0.001, 1], [0.001, 1, 1]])
def test_single_voxel(oracle, aniso):   # automated_test.py:7-98
  voxel = np.ones([1, 1, 1], dtype=bool, order="F")
  a, c = oracle.area(voxel, [0, 0, 0], [0, 0, 1], aniso)
  assert np.isclose(a, aniso[0] * aniso[1]) and c > 0
  a, c = oracle.area(voxel, [0, 0, 0], [0, 1, 0], aniso)
  assert np.isclose(a, aniso[0] * aniso[2]) and c > 0
  a, c = oracle.area(voxel, [0, 0, 0], [1, 0, 0], aniso)
  assert np.isclose(a, aniso[1] * aniso[2]) and c > 0
  a, c = oracle.area(voxel, [0, 0, 0], [1, 1, 0], aniso)
  assert np.isclose(a, np.sqrt(aniso[0] ** 2 + aniso[1] ** 2) * aniso[2])
  a, c = oracle.area(voxel, [0, 0, 0], [1, 1, 1], aniso)
  tri = lambda s: np.sqrt(3) / 8 * (s ** 2)
  hexagon = 0.75 if 0.001 in aniso else 2 * sum(tri(x) for x in aniso)
  assert np.isclose(a, hexagon) and c > 0
  a, c = oracle.area(voxel, [1, 0, 0], [1, 0, 0], aniso)
  assert a == 0 and c == 0
  assert oracle.area(voxel, [-1, 0, 0], [1, 0, 0], aniso)[0] == 0
  for incr in range(11):
    a = oracle.area(voxel, [0, 0, 0], [incr * 0.1, 0, 1])[0]
    assert 1 <= a <= np.sqrt(2)


def test_ccl_and_connectivity(oracle):   # automated_test.py:101-160
  img = np.zeros([10, 10, 10], dtype=bool, order="F")
  img[:3, :3, :3] = True
  img[6:, 6:, :3] = True
  a, c = oracle.area(img, [1, 1, 1], [0, 0, 1])
  assert a == 9 and c > 0
  assert area(oracle, img, [7, 7, 1], [0, 0, 1]) == 16
  assert area(oracle, img, [7, 7, 5], [0, 0, 1]) == 0
  img[:4, :4, :4] = True
  img[1, 1, 1] = False
  assert area(oracle, img, [0, 1, 1], [0, 0, 1]) == 15
  img[2:4, 2:4, 6:8] = True
  a, c = oracle.area(img, [2, 2, 6], [0, 0, 1])
  assert a == 4 and c == 0
  img = np.zeros([4, 4, 3], dtype=bool, order="F")
  for k in range(4):
    img[k, k] = True
  assert area(oracle, img, [0, 0, 0], [0, 0, 1]) == 4
  img = np.zeros([4, 4, 3], dtype=bool, order="F")
  img[3, 0] = img[2, 1] = img[1, 2] = img[3, 3] = True
  assert area(oracle, img, [3, 0, 0], [0, 0, 1]) == 3
  img = np.zeros([4, 4, 3], dtype=bool, order="F")
  img[3, 0] = img[1, 1] = img[0, 2] = img[3, 3] = True
  assert area(oracle, img, [3, 0, 0], [0, 0, 1]) == 1


def test_contact(oracle):   # automated_test.py:433-465
  labels = np.zeros([11, 11, 11], dtype=bool)
  labels[1:-1, 1:-1, 1:-1] = 1
  assert oracle.area(labels, [5, 5, 5], [0, 0, 1])[1] == 0
  labels = np.ones([11, 11, 11], dtype=bool)
  assert oracle.area(labels, [5, 5, 5], [0, 0, 1])[1] == 0b00001111
  assert oracle.area(labels, [5, 5, 5], [0, 1, 0])[1] == 0b00110011
  assert oracle.area(labels, [5, 5, 5], [1, 0, 0])[1] == 0b00111100
  assert oracle.area(labels, [5, 5, 5], [1, 1, 1])[1] == 0b00111111
  assert oracle.area(labels, [5, 5, 2], [1, 1, 1])[1] == 0b00111111
  assert oracle.area(labels, [100, 5, 2], [1, 1, 1])[1] == 0
  labels = np.zeros([11, 11, 11], dtype=bool)
  labels[1:-1, :, :] = 1
  assert oracle.area(labels, [5, 5, 5], [1, 1, 1])[1] == 0b00111100


# ---- against the compiled reference on fresh inputs (dev container / any box that has oracle/_ref) -------------
def test_vs_reference_random(oracle, ref):
  rng = np.random.default_rng(99)
  for it in range(400):
    vol, p, n, w = cases.rand_case(rng, max_side=20)
    a, c = ref.area(vol, p, n, w)
    a2, c2 = oracle.area(vol, p, n, w)
    assert bits(a) == bits(a2) and c == c2, it
    if it % 4 == 0:
      for m in (0, 1, 2):
        s, c = ref.section(vol, p, n, w, method=m)
        s2, c2 = oracle.section(vol, p, n, w, method=m)
        assert gu.same_bits(s, s2) and c == c2, (it, m)
      a, c = ref.area(vol, p, n, w, slow_method=True)
      a2, c2 = oracle.area(vol, p, n, w, slow_method=True)
      assert bits(a) == bits(a2) and c == c2, it


def test_vs_reference_slices(oracle, ref):
  rng = np.random.default_rng(5)
  for it in range(300):
    lab, p, n, w, sb, crop = cases.rand_labels_case(rng)
    r = ref.projection(lab, p, n, w, sb, crop)
    oracle.lib.oracle_projection_clamped()
    o = oracle.projection(lab, p, n, w, sb, crop)
    ub = oracle.lib.oracle_projection_clamped() > 0
    assert r.shape == o.shape and r.dtype == o.dtype, it
    if not ub:
      assert gu.same_bits(r, o), it


def test_is_26_connected_exhaustive(oracle, ref):
  orc = oracle.lib
  orc.oracle_is_26_connected.restype = int
  import ctypes as C
  orc.oracle_is_26_connected.argtypes = [C.c_uint8, C.c_uint8, C.c_int, C.c_int, C.c_int]
  for ox in (-2, 0, 2):
    for oy in (-2, 0, 2):
      for oz in (-2, 0, 2):
        for centre in range(0, 256, 3):
          for cand in range(0, 256, 5):
            assert bool(orc.oracle_is_26_connected(centre, cand, ox, oy, oz)) == ref.is_26_connected(centre, cand, (ox, oy, oz))


def test_twod_oracle_matches_reference_golden():
  """oracle/twod_oracle.py (restatement of xs3d/twod.py:73-174) against the outputs of the reference's own twod.py:
  float64 areas equal bit for bit, contacts equal."""
  import os
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
  import twod_oracle
  n = 0
  for i, img, p, v, w, area, contact, raises in gu.twod_cases():
    if raises:
      with pytest.raises(ValueError):
        twod_oracle.cross_sectional_area_2d(img, p, v, w)
    else:
      a, c = twod_oracle.cross_sectional_area_2d(img, p, v, w)
      assert a == area and c == contact, (i, a, area, c, contact)
    n += 1
  assert n == 600


def test_twod_oracle_reference_known_answers():
  """automated_test.py:393-418 (test_2d) on the restatement."""
  import os
  import sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
  from twod_oracle import cross_sectional_area_2d as f
  labels = np.ones([3, 3], dtype=bool)
  assert f(labels, [1, 1], [0, 1])[0] == 3
  assert f(labels, [1, 1], [0, 1], [3, 3])[0] == 9
  assert f(labels, [1, 1], [1, 0], [1, 1])[0] == 3
  assert f(labels, [1, 1], [1, 0], [5, 5])[0] == 15
  assert np.isclose(f(labels, [0, 0], [-1, 1], [1, 1])[0], 3 * np.sqrt(2))
  assert f(labels, [-1, 0], [-1, 1], [1, 1])[0] == 0
  assert f(labels, [1, -1], [-1, 1], [1, 1])[0] == 0
  labels[1, 1] = 0
  assert f(labels, [1, 1], [0, 1], [1, 1])[0] == 0
