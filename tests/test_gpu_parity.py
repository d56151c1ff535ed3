"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes, xs3d/_abi.py), against the oracle on the
same seeded inputs, against the committed golden vectors produced by the unmodified reference, against
the compiled reference when oracle/_ref travelled with the repo, and through size-independent
properties at BASELINE.json's full sizes.  Integer / mask / contact results must be bit-exact; areas and
voxel contributions are compared as float32 BIT PATTERNS (stricter than north_star's 1e-5 relative)."""
import numpy as np
import pytest

import cases
import golden_util as gu
from conftest import bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xs(xs3d_gpu):
  assert xs3d_gpu._abi.device_count() >= 1
  return xs3d_gpu


@pytest.fixture(scope="module")
def neurite512():
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(512, seed=0)
  assert vol.sum() == 377673 and len(verts) == 2724
  return vol, verts, tans


def _cubes_of(xs, v):
  import torch
  with xs.Volume(v) as V:
    ptr, nbytes = V.cubes()
    holder = type("H", (), {})()
    holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(holder, device="cuda").cpu().numpy()


def _cubes_numpy(vol):
  shape = vol.shape
  h = [(s + 1) // 2 for s in shape]
  pad = np.zeros([2 * x for x in h], bool)
  pad[: shape[0], : shape[1], : shape[2]] = vol
  exp = np.zeros(h, np.uint8)
  for b in range(8):
    exp |= (pad[(b & 1)::2, ((b >> 1) & 1)::2, (b >> 2)::2].astype(np.uint8) << b)
  return exp.ravel(order="F")


def test_ingest_matches_numpy(xs):
  """The occupancy-byte volume equals compute_cube() of every block, for aligned/unaligned sizes and both memory orders."""
  rng = np.random.default_rng(0)
  for shape in [(64, 48, 32), (33, 17, 9), (1, 1, 1), (16, 1, 5), (128, 64, 2), (50, 50, 50)]:
    vol = rng.random(shape) < 0.4
    for order in ("F", "C"):
      v = np.asfortranarray(vol) if order == "F" else np.ascontiguousarray(vol)
      assert np.array_equal(_cubes_of(xs, v), _cubes_numpy(vol)), (shape, order)


def test_bit_ingest_kernels_any_size(xs):
  """pack_bits + load_packed on small volumes of awkward sizes: rows of the bit stream that start at any bit (Fortran
  order with sx % 16 != 0, C order with sz % 8 != 0), partial tiles, single voxels."""
  import torch
  rng = np.random.default_rng(11)
  shapes = [(1, 1, 1), (3, 5, 7), (17, 9, 33), (33, 17, 9), (50, 50, 50), (64, 48, 32), (65, 3, 260), (130, 7, 66), (31, 2, 300), (16, 16, 40)]
  for shape in shapes:
    vol = rng.random(shape) < 0.45
    exp = _cubes_numpy(vol)
    for order in ("F", "C"):
      v = np.asfortranarray(vol) if order == "F" else np.ascontiguousarray(vol)
      pb, c_order = xs.pack_bits(v)
      with xs.Volume(shape=shape) as V:
        V.load_packed(pb, c_order=c_order)
        ptr, nbytes = V.cubes()
        holder = type("H", (), {})()
        holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        got = torch.as_tensor(holder, device="cuda").cpu().numpy()
      assert np.array_equal(got, exp), (shape, order)


def test_bit_packed_upload_matches_numpy(xs):
  """Host images of >= 8 Mi voxels go up bit-packed (xs_pack_bits on the host cores, ingest_bits_* kernels on the
  device): same occupancy bytes as the plain upload, for the aligned Fortran fast path, odd sizes and C order, from
  pageable and from page-locked memory, and as the device-resident plain path."""
  import torch
  rng = np.random.default_rng(7)
  for shape in [(256, 256, 128), (272, 190, 171), (203, 211, 197), (2048, 64, 65), (131, 201, 336)]:
    vol = rng.random(shape) < 0.35
    assert vol.size >= (8 << 20)
    exp = _cubes_numpy(vol)
    for order in ("F", "C"):
      v = np.asfortranarray(vol) if order == "F" else np.ascontiguousarray(vol)
      assert np.array_equal(_cubes_of(xs, v), exp), (shape, order)
      # page-locked source
      rt = torch.cuda.cudart()
      assert int(rt.cudaHostRegister(v.ctypes.data, v.nbytes, 0)) == 0
      try:
        assert np.array_equal(_cubes_of(xs, v), exp), (shape, order, "pinned")
      finally:
        rt.cudaHostUnregister(v.ctypes.data)
      # the two halves of the packed upload as separate calls: bits from host memory and from device memory
      pb, c_order = xs.pack_bits(v)
      assert c_order == (order == "C")
      for src in (pb, torch.from_numpy(pb).cuda()):
        with xs.Volume(shape=shape) as V:
          V.load_packed(src, c_order=c_order)
          ptr, nbytes = V.cubes()
          holder = type("H", (), {})()
          holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
          assert np.array_equal(torch.as_tensor(holder, device="cuda").cpu().numpy(), exp), (shape, order, "load_packed")
      # upload + ingest ahead of use: stage into either slot, commit (a buffer swap) later
      with xs.Volume(shape=shape) as V:
        pinned = torch.from_numpy(pb).pin_memory()
        for slot in (0, 1):
          V.load(np.zeros(shape, bool, order="F"))
          if slot == 0:
            V.stage_packed(pinned, slot, c_order=c_order)
          else:
            V.pack_and_stage(v, slot)                        # packing, upload and ingest in one call, from the byte image
          V.commit_staged(slot)
          torch.cuda.synchronize()
          ptr, nbytes = V.cubes()
          holder = type("H", (), {})()
          holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
          assert np.array_equal(torch.as_tensor(holder, device="cuda").cpu().numpy(), exp), (shape, order, "staged", slot)
      # the plain (unpacked) path: the same image handed over as a device tensor
      assert np.array_equal(_cubes_of(xs, torch.from_numpy(v.view(np.uint8)).cuda().view(torch.bool)), exp), (shape, order)


def test_golden_small_area_and_sections(xs, golden):
  for i, vol, p, n, w in gu.small_cases(golden):
    a, c = xs.cross_sectional_area(vol, p, n, w, return_contact=True)
    assert bits(a) == bits(golden["small_area"][i]) and c == golden["small_contact"][i], i
    if i % 2 == 0:
      a, c = xs.cross_sectional_area(vol, p, n, w, return_contact=True, slow_method=True)
      assert bits(a) == bits(golden["small_slow_area"][i]) and c == golden["small_slow_contact"][i], i
      for m in (0, 1, 2):
        img, c = xs.cross_section(vol, p, n, w, return_contact=True, method=m)
        exp, ec = gu.section_golden(golden, m, i, vol.shape)
        assert img.dtype == np.float32 and img.flags.f_contiguous
        assert gu.same_bits(img, np.asfortranarray(exp)) and c == ec, (i, m)


def test_golden_slices(xs, golden):
  for i, lab, p, n, w, sb, crop, exp, ub in gu.slice_cases(golden):
    got = xs.slice(lab, p, n, w, sb, crop)
    assert got.shape == exp.shape and got.dtype == exp.dtype, i
    if not ub:
      assert gu.same_bits(got, exp), i


def test_random_vs_oracle(xs, oracle):
  rng = np.random.default_rng(31337)
  for it in range(300):
    vol, p, n, w = cases.rand_case(rng, max_side=24)
    a, c = oracle.area(vol, p, n, w)
    ga, gc = xs.cross_sectional_area(vol, p, n, w, return_contact=True)
    assert bits(a) == bits(ga) and c == gc, (it, vol.shape, p, n, w)
    if it % 6 == 0:
      s, c = oracle.section(vol, p, n, w, method=0)
      gs, gc = xs.cross_section(vol, p, n, w, return_contact=True)
      assert gu.same_bits(s, gs) and c == gc, it
    if it % 10 == 0:
      cv = np.ascontiguousarray(vol)   # C-order input: ingested on the device without a host transpose
      ga2 = xs.cross_sectional_area(cv, p, n, w)
      assert bits(ga2) == bits(a), it


def test_random_slices_vs_oracle(xs, oracle):
  rng = np.random.default_rng(99)
  for it in range(200):
    lab, p, n, w, sb, crop = cases.rand_labels_case(rng, max_side=40)
    o = oracle.projection(lab, p, n, w, sb, crop)
    g = xs.slice(lab, p, n, w, sb, crop)
    assert gu.same_bits(o, g), (it, lab.shape, lab.dtype, p, n, w, sb, crop)


def test_reference_known_answers(xs):
  """automated_test.py:101-160, 252-286, 288-292, 433-465 through the drop-in API."""
  img = np.zeros([10, 10, 10], dtype=bool, order="F")
  img[:3, :3, :3] = True
  img[6:, 6:, :3] = True
  a, c = xs.cross_sectional_area(img, [1, 1, 1], [0, 0, 1], return_contact=True)
  assert a == 9 and c > 0
  assert xs.cross_sectional_area(img, [7, 7, 1], [0, 0, 1]) == 16
  assert xs.cross_sectional_area(img, [7, 7, 5], [0, 0, 1]) == 0
  img[:4, :4, :4] = True
  img[1, 1, 1] = False
  assert xs.cross_sectional_area(img, [0, 1, 1], [0, 0, 1]) == 15
  assert xs.cross_sectional_area(np.ones([5, 5, 1], dtype=bool), [0, 0, 0], [0, 0, 1]) == 25
  labels = np.ones((5, 5, 5), dtype=bool, order="F")
  for nrm in ([1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]):
    assert xs.cross_sectional_area(labels, [2, 2, 2], nrm) == 25
  for nrm in ([1, 1, 0], [1, 0, 1], [0, 1, 1], [-1, 1, 0], [0, -1, 1]):
    assert np.isclose(xs.cross_sectional_area(labels, [2, 2, 2], nrm), 25 * np.sqrt(2))
  assert xs.cross_sectional_area(np.zeros([0, 0, 0], dtype=bool), [0, 0, 0], [1, 1, 1]) == 0
  labels = np.ones([11, 11, 11], dtype=bool)
  for nrm, exp in (([0, 0, 1], 0b001111), ([0, 1, 0], 0b110011), ([1, 0, 0], 0b111100), ([1, 1, 1], 0b111111)):
    assert xs.cross_sectional_area(labels, [5, 5, 5], nrm, return_contact=True)[1] == exp
  assert xs.cross_sectional_area(labels, [100, 5, 2], [1, 1, 1], return_contact=True) == (0.0, 0)
  lab = np.arange(9, dtype=np.uint8).reshape([3, 3, 1], order="F")
  assert np.all(xs.slice(lab, [0, 0, 0], [0, 0, 1], standardize_basis=True) == lab[:, :, 0])


def test_cross_section_sums_to_area(xs):   # automated_test.py:317-330
  labels = np.ones((5, 5, 5), dtype=bool, order="F")
  for theta in range(0, 25, 3):
    nrm = [0, np.cos(theta / 25 * 2 * np.pi), np.sin(theta / 25 * 2 * np.pi)]
    a = xs.cross_sectional_area(labels, (2, 2, 2), nrm)
    for method in (0, 1):
      img = xs.cross_section(labels, (2, 2, 2), nrm, method=method)
      assert img.dtype == np.float32 and np.isclose(img.sum(), a)


def test_neurite_batch_golden_and_tiers(xs, golden):
  """1000 skeleton queries on the 256^3 neurite: float32-bit-identical to the reference's outputs."""
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(256, seed=0)
  qp, qn = synthetic.draw_queries(verts, tans, 1000, seed=1)
  with xs.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, [32, 32, 40], return_contact=True)
    st = V.stats()
  assert np.array_equal(bits(a), bits(golden["neurite256_area"]))
  assert np.array_equal(c, golden["neurite256_contact"])
  assert st["samples"] > 0 and st["launches"] >= 2


def test_host_sweep_entry_point_and_threads(xs, golden):
  """xs3d_area_sweep_host (image + queries from host in one call) equals load + batch, also when two volumes
  are swept concurrently from two host threads (the pipelined end-to-end mode of bench.py)."""
  import threading
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(256, seed=0)
  qp, qn = synthetic.draw_queries(verts, tans, 1000, seed=1)
  w = [32, 32, 40]
  a, c = xs.cross_sectional_area_batch(vol, qp, qn, w, return_contact=True)
  assert np.array_equal(bits(a), bits(golden["neurite256_area"])) and np.array_equal(c, golden["neurite256_contact"])
  cvol = np.ascontiguousarray(vol)
  res = {}

  def worker(k, image):
    with xs.Volume(shape=vol.shape) as V:
      for _ in range(3):
        res[k] = V.sweep(image, qp, qn, w, return_contact=True)

  ts = [threading.Thread(target=worker, args=(0, vol)), threading.Thread(target=worker, args=(1, cvol))]
  [t.start() for t in ts]
  [t.join() for t in ts]
  for k in (0, 1):
    assert np.array_equal(bits(res[k][0]), bits(a)) and np.array_equal(res[k][1], c)


def test_stream_volumes_single_rank(xs, golden):
  """xs3d.dist.stream_volumes on one GPU: a sequence of different host volumes (both memory orders) goes through one
  resident Volume — packed + uploaded + ingested ahead on a helper thread, swapped in per step — and every step's
  sweep sees exactly its own volume."""
  from xs3d import synthetic, dist as xdist
  w = [32, 32, 40]
  vols, queries, expected = [], [], []
  for seed in (0, 5, 9, 0):
    vol, verts, tans = synthetic.make_neurite(256, seed=seed)
    qp, qn = synthetic.draw_queries(verts, tans, 400, seed=seed + 1)
    vols.append(vol if seed != 5 else np.ascontiguousarray(vol))
    queries.append((qp, qn))
    expected.append(xs.cross_sectional_area_batch(vol, qp, qn, w, return_contact=True))
  assert np.array_equal(bits(expected[0][0]), bits(xs.cross_sectional_area_batch(vols[0], *queries[0], w)))
  with xs.Volume(shape=vols[0].shape) as V:
    res = xdist.stream_volumes(V, vols, len(vols), lambda k: V.cross_sectional_area_batch(*queries[k], w, return_contact=True))
    for k in range(len(vols)):
      assert np.array_equal(bits(res[k][0]), bits(expected[k][0])) and np.array_equal(res[k][1], expected[k][1]), k
    # a second pass over the same handle (the staging slots are reused), fewer steps than slots
    res = xdist.stream_volumes(V, vols[1:2], 1, lambda k: V.cross_sectional_area_batch(*queries[1], w, return_contact=True))
    assert np.array_equal(bits(res[0][0]), bits(expected[1][0]))


def test_neurite512_sweep_parity_and_properties(xs, oracle, neurite512):
  """BASELINE config 3: the 10k-vertex sweep.  Oracle parity on a 1500-query sample, then
  size-independent properties on all 10k: determinism, order independence (permuting the batch
  permutes the result), batch == single-call, C-order ingest == F-order ingest."""
  from xs3d import synthetic
  vol, verts, tans = neurite512
  qp, qn = synthetic.draw_queries(verts, tans, 10000, seed=1)
  w = [32, 32, 40]
  with xs.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    st = V.stats()
    a2, c2 = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    perm = np.random.default_rng(0).permutation(len(qp))
    a3, c3 = V.cross_sectional_area_batch(qp[perm], qn[perm], w, return_contact=True)
    singles = [V.cross_sectional_area(qp[i], qn[i], w, return_contact=True) for i in range(0, 10000, 500)]
  oa, oc = oracle.area_batch(vol, qp[:1500], qn[:1500], w)
  assert np.array_equal(bits(a[:1500]), bits(oa)) and np.array_equal(c[:1500], oc)
  assert np.array_equal(bits(a), bits(a2)) and np.array_equal(c, c2)
  assert np.array_equal(bits(a[perm]), bits(a3)) and np.array_equal(c[perm], c3)
  for k, i in enumerate(range(0, 10000, 500)):
    assert bits(singles[k][0]) == bits(a[i]) and singles[k][1] == c[i]
  assert (a > 0).all() and st["escalated_mid"] > 0
  with xs.Volume(np.ascontiguousarray(vol)) as VC:
    a4 = VC.cross_sectional_area_batch(qp[:2000], qn[:2000], w)
  assert np.array_equal(bits(a4), bits(a[:2000]))


def test_neurite512_vs_compiled_reference(xs, ref, neurite512):
  """When oracle/_ref (the unmodified reference, compiled in the dev container) travelled with the repo:
  all 10k queries of the sweep, bit for bit."""
  from xs3d import synthetic
  vol, verts, tans = neurite512
  qp, qn = synthetic.draw_queries(verts, tans, 10000, seed=1)
  with xs.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, [32, 32, 40], return_contact=True)
  ra, rc = ref.area_batch(vol, qp, qn, [32, 32, 40])
  assert np.array_equal(bits(a), bits(ra)) and np.array_equal(c, rc)


def test_config0_sphere_and_config1_cylinder(xs, golden):
  """BASELINE configs 0/1 at full size against the reference's stored outputs (large single sections:
  CTA tier with the bitmap in shared memory, ring-spilled walk, global hash)."""
  from xs3d import synthetic
  sph = synthetic.make_sphere()
  a, c = xs.cross_sectional_area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], return_contact=True)
  assert bits(a) == bits(golden["sphere_area"]) and c == golden["sphere_contact"]
  img = xs.cross_section(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40])
  assert np.count_nonzero(img) == golden["sphere_section_nnz"]
  assert img.sum(dtype=np.float64) == golden["sphere_section_sum"]
  xs.set_shape(sph)
  a2 = xs.cross_sectional_area(sph, [250, 250, 250], [0.01, 0.033, 0.9], [32, 32, 40], use_persistent_data=True)
  xs.clear_shape()
  assert bits(a2) == bits(a)
  del img, sph
  cyl, axis = synthetic.make_cylinder()
  with xs.Volume(cyl) as V:
    for k in range(3):
      a, c = V.cross_sectional_area(golden["cyl_verts"][k], axis, [32, 32, 40], return_contact=True)
      assert bits(a) == bits(golden["cyl_area"][k]) and c == golden["cyl_contact"][k], k
    idx, val, ct = V.cross_section_sparse(golden["cyl_verts"][0], axis, [32, 32, 40])
  assert len(idx) == golden["cyl_section_nnz"] and len(np.unique(idx)) == len(idx)
  order = np.argsort(idx)
  assert val[order].sum(dtype=np.float64) == golden["cyl_section_sum"]


def test_slice_batch_matches_single(xs, oracle, neurite512):
  """BASELINE config 4 proxy at 512^3 / uint64: batched cropped slices equal single-call slices and the oracle."""
  from xs3d import synthetic
  vol, verts, tans = neurite512
  lab = synthetic.make_labels(vol, np.uint64, hashed=True)
  qp, qn = synthetic.draw_queries(verts, tans, 300, seed=2)
  w = [32, 32, 40]
  with xs.Volume(labels=lab) as V:
    res = V.slice_batch(qp, qn, w, crop=100)
    res_big = V.slice_batch(qp[:20], qn[:20], w, crop=1000)
  for i in range(0, 300, 7):
    o = oracle.projection(lab, qp[i], qn[i], w, True, 100.0)
    assert gu.same_bits(np.asfortranarray(res[i]), o), i
  for i in range(0, 20, 4):
    o = oracle.projection(lab, qp[i], qn[i], w, True, 1000.0)
    assert gu.same_bits(np.asfortranarray(res_big[i]), o), i
  o = oracle.projection(lab, qp[0], qn[0], w, True, float("inf"))
  g = xs.slice(lab, qp[0], qn[0], w)
  assert gu.same_bits(g, o)


def test_config3_neurite1024_sweep(xs, oracle):
  """BASELINE config 3: 10k skeleton queries on the 1024^3 neurite (1 GiB image, 128 MiB occupancy bytes).
  Oracle parity on a 300-query sample, size-independent properties on all 10k."""
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(1024, seed=0)
  assert vol.sum() > 500000 and len(verts) > 3000
  qp, qn = synthetic.draw_queries(verts, tans, 10000, seed=1)
  w = [32, 32, 40]
  with xs.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    perm = np.random.default_rng(1).permutation(len(qp))
    a2, c2 = V.cross_sectional_area_batch(qp[perm], qn[perm], w, return_contact=True)
  oa, oc = oracle.area_batch(vol, qp[:300], qn[:300], w)
  assert np.array_equal(bits(a[:300]), bits(oa)) and np.array_equal(c[:300], oc)
  assert np.array_equal(bits(a[perm]), bits(a2)) and np.array_equal(c[perm], c2)
  assert (a > 0).all()


def test_config4_slice_1024_uint64(xs, oracle):
  """BASELINE config 4: cropped slices of a 1024^3 uint64 label volume (8 GiB), 10k planes in one batch.
  Skipped when the host cannot hold the label volume."""
  import psutil
  if psutil.virtual_memory().available < 40 * 2 ** 30:
    pytest.skip("needs ~20 GiB of free host memory")
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(1024, seed=0)
  lab = synthetic.make_labels(vol, np.uint64, hashed=True)
  del vol
  qp, qn = synthetic.draw_queries(verts, tans, 10000, seed=2)
  w = [32, 32, 40]
  with xs.Volume(labels=lab) as V:
    res = V.slice_batch(qp, qn, w, crop=100)
  assert len(res) == 10000
  for i in range(0, 10000, 997):
    o = oracle.projection(lab, qp[i], qn[i], w, True, 100.0)
    assert gu.same_bits(np.asfortranarray(res[i]), o), i


def test_select_label_and_skeleton_sweep(xs, oracle):
  """a resident label volume serves each object through xs3d_volume_select_label (mask built on the
  device), and sweep_skeleton runs a whole skeleton in one call.  Both equal the per-object reference path."""
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(192, seed=3, n_branches=5)
  rng = np.random.default_rng(0)
  other = np.asfortranarray(rng.random(vol.shape) < 0.02) & ~vol
  for dt in (np.uint8, np.uint32, np.uint64):
    lab = np.zeros(vol.shape, dtype=dt, order="F")
    lab[vol] = 7
    lab[other] = 9
    for arr in (lab, np.ascontiguousarray(lab)):
      with xs.Volume(labels=arr) as V:
        for label, mask in ((7, vol), (9, other)):
          V.select_label(label)
          idx = np.argwhere(mask)[:: max(1, int(mask.sum()) // 40)][:40].astype(np.float32)
          nr = np.tile(np.array([0.3, -0.2, 0.9], np.float32), (len(idx), 1))
          a, c = V.cross_sectional_area_batch(idx, nr, [32, 32, 40], return_contact=True)
          oa, oc = oracle.area_batch(mask, idx, nr, [32, 32, 40])
          assert np.array_equal(bits(a), bits(oa)) and np.array_equal(c, oc), (dt, label)
  # skeleton sweep: tangents on the device (deterministic fixed-point sums), sections of the planes normal to them
  import torch
  edges = np.stack([np.arange(len(verts) - 1), np.arange(1, len(verts))], 1)
  edges = np.concatenate([edges, [[5, 5], [0, len(verts) - 1]], edges[::7][:, ::-1]])      # a zero-length edge, a long one, reversed duplicates
  with xs.Volume(vol) as V:
    nrm = V.skeleton_tangents(verts, edges)
    nrm2 = V.skeleton_tangents(verts, edges)
    dn = V.skeleton_tangents(torch.from_numpy(verts.astype(np.float32)).cuda(), torch.from_numpy(edges).cuda())
    a, c = V.sweep_skeleton(verts, edges, [32, 32, 40], return_contact=True)
    da, dc = V.sweep_skeleton(torch.from_numpy(verts.astype(np.float32)).cuda(), torch.from_numpy(edges).cuda(), [32, 32, 40], return_contact=True)
  host = xs.skeleton_tangents(verts, edges)
  assert nrm.dtype == np.float32 and np.array_equal(bits(nrm), bits(nrm2)) and np.array_equal(bits(nrm), bits(dn.cpu().numpy()))
  assert np.allclose(nrm, host, atol=2e-6) and np.allclose(np.linalg.norm(nrm, axis=1), 1, atol=1e-6)
  oa, oc = oracle.area_batch(vol, verts.astype(np.float32), nrm, [32, 32, 40])
  assert np.array_equal(bits(a), bits(oa)) and np.array_equal(c, oc)
  assert np.array_equal(bits(da.cpu().numpy()), bits(oa)) and np.array_equal(dc.cpu().numpy(), oc)
  a2 = xs.sweep_skeleton(vol, verts, edges, [32, 32, 40])
  assert np.array_equal(bits(a2), bits(oa))
  lone = xs.Volume(vol)
  assert np.array_equal(lone.skeleton_tangents(verts[:3], np.zeros((0, 2), np.int64)), np.tile(np.float32([0, 0, 1]), (3, 1)))
  lone.close()


def test_select_label_occupancy_all_dtypes(xs):
  """xs3d_volume_select_label builds exactly the occupancy bytes of (labels == L): every item size, both memory orders,
  row lengths that take the 16-byte fast path and ones that do not, labels with set high bits / equal low halves."""
  import torch
  rng = np.random.default_rng(21)
  for shape in [(64, 23, 17), (48, 9, 31), (50, 20, 10), (7, 5, 3), (20, 9, 48), (70, 5, 272), (33, 4, 24)]:
    for dt in (np.uint8, np.uint16, np.uint32, np.uint64, np.int16, np.int64):
      top = np.iinfo(dt).max
      palette = np.array([0, 1, top, top - 1, top // 2 + 1, 258 % (int(top) + 1)], dtype=dt)
      if np.dtype(dt).itemsize == 8:
        palette = np.concatenate([palette, np.array([(5 << 32) | 7, (6 << 32) | 7, (5 << 32) | 8], dtype=dt)])   # halves that collide
      lab = palette[rng.integers(0, len(palette), shape)]
      for arr in (np.asfortranarray(lab), np.ascontiguousarray(lab)):
        with xs.Volume(labels=arr) as V:
          for L in palette[1:]:
            V.select_label(int(L))
            ptr, nbytes = V.cubes()
            holder = type("H", (), {})()
            holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            got = torch.as_tensor(holder, device="cuda").cpu().numpy()
            assert np.array_equal(got, _cubes_numpy(lab == L)), (shape, dt, int(L), arr.flags.f_contiguous)


def test_null_normal_is_rejected_on_every_path(xs):
  """An all-zero normal is a ValueError in the reference (xs3d/__init__.py:72-75): host arrays are checked before
  the batch, CUDA tensors by the pre-pass kernel (no host round trip before the launch)."""
  import torch
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(128, seed=0, n_branches=4)
  qp, qn = synthetic.draw_queries(verts, tans, 64, seed=1)
  qn = qn.copy(); qn[17] = 0
  with xs.Volume(vol) as V:
    with pytest.raises(ValueError):
      V.cross_sectional_area_batch(qp, qn, [32, 32, 40])
    with pytest.raises(ValueError):
      V.cross_sectional_area_batch(torch.from_numpy(qp).cuda(), torch.from_numpy(qn).cuda(), [32, 32, 40])
    with pytest.raises(ValueError):
      V.cross_sectional_area_batch(torch.from_numpy(qp[16:18]).cuda(), torch.from_numpy(qn[16:18]).cuda(), [32, 32, 40])
    a = V.cross_sectional_area_batch(torch.from_numpy(qp[:16]).cuda(), torch.from_numpy(qn[:16]).cuda(), [32, 32, 40])
    b = V.cross_sectional_area_batch(qp[:16], qn[:16], [32, 32, 40])
    assert np.array_equal(bits(a.cpu().numpy()), bits(b))


def test_long_and_wide_sections_all_cta_tiers(xs, oracle):
  """Sections of every size class through the device tiers: tubes cut (almost) along their axis give sections
  hundreds of cells long — beyond the warp tier, the mid tier's 256-cell window (wide tier) and the wide tier's
  416-cell window (big tier).  Batched (tier hand-overs while the tiers run side by side) and one by one."""
  rng = np.random.default_rng(5)
  N = 352
  g = np.stack(np.meshgrid(*[np.arange(N, dtype=np.float32)] * 3, indexing="ij"), -1)
  vol = np.zeros((N, N, N), bool)
  qp, qn = [], []
  for it in range(5):
    c0 = rng.uniform(120, 230, size=3).astype(np.float32)
    d = rng.normal(size=3).astype(np.float32); d /= np.linalg.norm(d)
    r = rng.uniform(3, 8)
    rel = g - c0
    t = rel @ d
    dist2 = (rel ** 2).sum(-1) - t ** 2
    half = [40, 90, 150, 230, 400][it]                       # half-length of the tube in voxels
    tube = (dist2 <= r * r) & (np.abs(t) <= half)
    tube ^= (rng.random(tube.shape) < 0.02) & (dist2 <= (r + 2) ** 2) & (np.abs(t) <= half)
    vol |= tube
    vol[int(c0[0]), int(c0[1]), int(c0[2])] = True
    e = np.cross(d, rng.normal(size=3)).astype(np.float32); e /= np.linalg.norm(e)
    for k in range(3):
      qp.append(np.floor(c0)); qn.append(e + rng.uniform(0.0, 0.1) * d)   # the plane (almost) contains the tube axis
    qp.append(np.floor(c0)); qn.append(d)                                  # and one cut across the tube
  del g
  vol = np.asfortranarray(vol)
  qp = np.array(qp, np.float32); qn = np.array(qn, np.float32)
  w = [32, 32, 40]
  oa, oc = oracle.area_batch(vol, qp, qn, w)
  with xs.Volume(vol) as V:
    a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    st = V.stats()
    # a batch that follows one with wide sections launches the wide / big tiers in its first round (no host round trip)
    a2, c2 = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    st2 = V.stats()
    # ... and one without wide sections switches that off again for the batch after it
    small = [i for i in range(len(qp)) if i % 4 == 3]
    a3, c3 = V.cross_sectional_area_batch(np.repeat(qp[small], 2, 0), np.repeat(qn[small], 2, 0), w, return_contact=True)
    a4, c4 = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
    singles = [V.cross_sectional_area(qp[i], qn[i], w, return_contact=True) for i in range(len(qp))]
  assert np.array_equal(bits(a), bits(oa)) and np.array_equal(c, oc)
  assert np.array_equal(bits(a2), bits(oa)) and np.array_equal(c2, oc)
  assert np.array_equal(bits(a4), bits(oa)) and np.array_equal(c4, oc)
  assert np.array_equal(bits(a3), bits(np.repeat(oa[small], 2))) and np.array_equal(c3, np.repeat(oc[small], 2))
  assert st2["escalated_big"] == st["escalated_big"], (st, st2)
  for i, (sa, sc) in enumerate(singles):
    assert bits(np.float32(sa)) == bits(oa[i]) and sc == oc[i], i
  assert st["escalated_mid"] > 0 and st["escalated_big"] > 0, st     # the mid tier handed sections on to the wide / big tiers


def _sparse_of(img):
  flat = img.ravel(order="F")
  idx = np.flatnonzero(flat)
  return idx.astype(np.uint64), flat[idx]


def test_config1_config2_sections_bitwise(xs, oracle):
  """BASELINE configs 1 and 2 as written: the cross_section contribution image AND the contact bitfield, at full size,
  voxel by voxel as float32 bit patterns against the oracle — the 500^3 sphere and all three vertices of the 512^3
  cylinder, two of which sit where the tube leaves the volume (contact != 0)."""
  from xs3d import synthetic
  w = [32, 32, 40]
  sph = synthetic.make_sphere()
  p, n = [250, 250, 250], [0.01, 0.033, 0.9]
  oimg, oct_ = oracle.section(sph, p, n, w)
  oidx, oval = _sparse_of(oimg)
  del oimg
  with xs.Volume(sph) as V:
    idx, val, ct = V.cross_section_sparse(p, n, w)
  order = np.argsort(idx)
  assert ct == oct_ and np.array_equal(idx[order], oidx) and np.array_equal(bits(val[order]), bits(oval))
  del sph
  cyl, axis = synthetic.make_cylinder()
  verts = [(256, 256, 256), (496, 352, 316), (12, 158, 195)]          # SURVEY.md §8d; tests/golden: cyl_verts
  contacts = []
  with xs.Volume(cyl) as V:
    for v in verts:
      oimg, oct_ = oracle.section(cyl, v, axis, w)
      oidx, oval = _sparse_of(oimg)
      del oimg
      idx, val, ct = V.cross_section_sparse(v, axis, w)
      order = np.argsort(idx)
      assert ct == oct_, (v, ct, oct_)
      assert len(oidx) > 0 and np.array_equal(idx[order], oidx) and np.array_equal(bits(val[order]), bits(oval)), v
      a, c = V.cross_sectional_area(v, axis, w, return_contact=True)
      oa, oc = oracle.area(cyl, v, axis, w)
      assert bits(a) == bits(oa) and c == oc, v
      contacts.append(ct)
  assert contacts[0] == 0 and contacts[1] != 0 and contacts[2] != 0     # the border vertices do touch the volume faces
  # the drop-in call (dense image + return_contact) on a border vertex
  img, ct = xs.cross_section(cyl, verts[1], axis, w, return_contact=True)
  oimg, oct_ = oracle.section(cyl, verts[1], axis, w)
  assert ct == oct_ and img.flags.f_contiguous and np.array_equal(img.view(np.uint32), oimg.view(np.uint32))


def test_use_persistent_data_evaluates_the_image_it_is_given(xs, oracle):
  """The reference's persistent data is scratch only (xs3d.hpp:963-971): after set_shape(img_a), a call with img_b of
  the same shape must answer for img_b."""
  rng = np.random.default_rng(4)
  shape = (40, 36, 28)
  img_a = np.asfortranarray(rng.random(shape) < 0.6)
  img_b = np.asfortranarray(rng.random(shape) < 0.6)
  img_a[20, 18, 14] = img_b[20, 18, 14] = True
  p, n, w = [20, 18, 14], [0.2, -0.5, 0.8], [4, 4, 40]
  xs.set_shape(img_a)
  try:
    got_a = xs.cross_sectional_area(img_a, p, n, w, return_contact=True, use_persistent_data=True)
    got_b = xs.cross_sectional_area(img_b, p, n, w, return_contact=True, use_persistent_data=True)
    got_bc = xs.cross_sectional_area(np.ascontiguousarray(img_b), p, n, w, return_contact=True, use_persistent_data=True)
  finally:
    xs.clear_shape()
  oa, ob = oracle.area(img_a, p, n, w), oracle.area(img_b, p, n, w)
  assert bits(oa[0]) != bits(ob[0])
  assert bits(got_a[0]) == bits(oa[0]) and got_a[1] == oa[1]
  assert bits(got_b[0]) == bits(ob[0]) and got_b[1] == ob[1]
  assert bits(got_bc[0]) == bits(ob[0]) and got_bc[1] == ob[1]


def test_batch_argument_validation(xs):
  """Lengths, dtypes and sizes are checked before raw pointers cross the C ABI."""
  import torch
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(128, seed=0, n_branches=4)
  qp, qn = synthetic.draw_queries(verts, tans, 64, seed=1)
  w = [32, 32, 40]
  with xs.Volume(vol) as V:
    with pytest.raises(ValueError):
      V.cross_sectional_area_batch(qp, qn[:32], w)
    with pytest.raises(ValueError):
      V.cross_sectional_area_batch(torch.from_numpy(qp).cuda(), torch.from_numpy(qn[:32]).cuda(), w)
    with pytest.raises(ValueError):
      V.sweep(vol, qp, qn[:10], w)
    for bad in [(np.empty(63, np.float32), np.empty(64, np.uint8)), (np.empty(64, np.float64), np.empty(64, np.uint8)),
                (np.empty(64, np.float32), np.empty(64, np.int32)), (np.empty(128, np.float32)[::2], np.empty(64, np.uint8)),
                (np.empty(64, np.float32),)]:
      with pytest.raises(ValueError):
        V.cross_sectional_area_batch(qp, qn, w, out=bad)
      with pytest.raises(ValueError):
        V.sweep(vol, qp, qn, w, out=bad)
    out = (np.empty(80, np.float32), np.empty(64, np.uint8))
    a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True, out=out)
    b = V.cross_sectional_area_batch(qp, qn, w)
    assert a is out[0] and np.array_equal(bits(a[:64]), bits(b))
    with pytest.raises(ValueError):
      V.stage_packed(np.zeros(100, np.uint8), 0)
    with pytest.raises(ValueError):
      V.stage_packed(np.zeros(vol.size // 8, np.uint8), 2)


def test_multi_gpu_every_rank_against_the_oracle():
  """N > 1 (skipped on a one-GPU box): torchrun, one process per GPU — the occupancy bytes NCCL-broadcast from rank 0,
  the query list sharded, results all-gathered; EVERY rank compares the gathered answers, its own shard computed by its
  own GPU included, with the oracle (tools/dist_check.py), and the same for the sharded slices."""
  import subprocess, sys, os
  import torch
  n = torch.cuda.device_count()
  if n < 2:
    pytest.skip("needs >= 2 GPUs")
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 8)}", "--master-addr", "127.0.0.1",
                      "--master-port", "29631", os.path.join(root, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
  assert r.stdout.count("bit-identical to the oracle: True") >= min(n, 8), r.stdout[-2000:]


def test_slice_batch_random_volumes_vs_oracle(xs, oracle):
  """The batched slice path (plan + fill, ragged output) on random small label volumes: every dtype, both memory orders,
  vertices outside the volume, axis-aligned and degenerate normals, every crop — 16 planes per volume in one batch,
  each cut-out bit-identical to the oracle's (and so are the shapes the plan reports before any label is read)."""
  rng = np.random.default_rng(2024)
  for it in range(60):
    lab, p, n, w, sb, crop = cases.rand_labels_case(rng, max_side=40)
    if crop == 0.0:
      crop = 5.0
    ps = [p] + [rng.uniform(-1, np.array(lab.shape) + 1).astype(np.float32) for _ in range(15)]
    ns = [n] + [cases.rand_normal(rng) for _ in range(15)]
    with xs.Volume(labels=lab) as V:
      shapes, offsets = V.slice_plan(ps, ns, w, sb, crop)
      res = V.slice_batch(ps, ns, w, sb, crop)
      one = V.slice(ps[0], ns[0], w, sb, crop)
    assert len(res) == 16
    for k in range(16):
      o = oracle.projection(lab, ps[k], ns[k], w, sb, crop)
      assert tuple(shapes[k]) == o.shape and int(offsets[k + 1] - offsets[k]) == o.size, (it, k)
      assert gu.same_bits(np.asfortranarray(res[k]), o), (it, k, lab.shape, lab.dtype, ps[k], ns[k], w, sb, crop)
    assert gu.same_bits(one, np.asfortranarray(res[0]))
  lab = np.arange(24, dtype=np.uint16).reshape(2, 3, 4)
  with xs.Volume(labels=lab) as V:
    assert all(r.shape == (0, 0) for r in V.slice_batch([[0, 0, 0]] * 3, [[0, 0, 1]] * 3, crop=0)[:])
    with pytest.raises(ValueError):
      V.slice_batch([[0, 0, 0]], [[0, 0, 0]])


def test_slice_batch_uncropped_planes_and_fixed_pitch(xs, oracle, neurite512):
  """un-cropped planes through a 512^3 uint64 label volume (cut-outs of up to ~870 x 870 cells) as ONE
  ragged batch, gathered in pieces (max_bytes smaller than the whole), against the oracle; the fixed-pitch entry point
  with device-resident inputs and outputs, where a cut-out larger than the pitch is flagged and left unwritten."""
  import ctypes as C
  import torch
  from xs3d import synthetic, _abi
  vol, verts, tans = neurite512
  lab = synthetic.make_labels(vol, np.uint64, hashed=True)
  qp, qn = synthetic.draw_queries(verts, tans, 12, seed=4)
  qn[3] = [0, 0, 1]; qn[4] = [1, 0, 0]; qn[5] = [1, 1, 0]
  w = [32, 32, 40]
  with xs.Volume(labels=lab) as V:
    res = V.slice_batch(qp, qn, w, crop=float("inf"), max_bytes=16 << 20)
    assert len(res.chunks) > 1                                     # gathered in several fills of one plan
    for i in range(12):
      o = oracle.projection(lab, qp[i], qn[i], w, True, float("inf"))
      assert o.shape[0] > 400 and gu.same_bits(np.asfortranarray(res[i]), o), i
    assert gu.same_bits(xs.slice(V, qp[0], qn[0], w), np.asfortranarray(res[0]))       # resident labels behind xs3d.slice
    # fixed pitch, everything on the device: crop = 300 gives ~24 x 24 cut-outs; the pitch holds only the smaller ones
    sizes = np.prod(V.slice_plan(qp, qn, w, True, 300.0)[0], axis=1)
    n, pitch = 12, int(np.sort(sizes)[n_small := 6])
    assert sizes.min() < sizes.max()
    dp, dn = torch.from_numpy(qp).cuda(), torch.from_numpy(qn).cuda()
    out = torch.full((n, pitch), 0xDEAD, dtype=torch.int64, device="cuda")
    shapes = torch.zeros((n, 2), dtype=torch.int64, device="cuda")
    over = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    _abi.check(V._lib.xs3d_projection_batch(V._h, n, dp.data_ptr(), dn.data_ptr(), _abi.fptr(_abi.f3(w)), 1, C.c_float(300.0),
                                            out.data_ptr(), pitch, shapes.data_ptr(), over.data_ptr(), _abi.DEVICE))
    out, shapes, over = out.cpu().numpy().view(np.uint64), shapes.cpu().numpy(), over.cpu().numpy()
    assert set(over.tolist()) == {0, 1}
    for i in range(n):
      o = oracle.projection(lab, qp[i], qn[i], w, True, 300.0)
      assert tuple(shapes[i]) == o.shape, i
      if over[i]:
        assert o.size > pitch and np.all(out[i] == np.uint64(0xDEAD))
      else:
        assert gu.same_bits(out[i, : o.size].reshape(o.shape, order="F"), o), i


def test_cross_section_component_topologies(xs, oracle):
  """Path B's union-find over lattice row runs on the shapes that stress it: dense random noise (thousands of short
  runs, components that merge late), a comb and a spiral (one component whose runs are linked through long detours,
  so roots change many times), rings around the vertex that must NOT be reached, planes along the axes and oblique."""
  rng = np.random.default_rng(77)
  N = 72
  g = np.stack(np.meshgrid(*[np.arange(N, dtype=np.float32)] * 3, indexing="ij"), -1)
  c0 = np.array([36, 36, 36], np.float32)
  rad = np.sqrt(((g[..., :2] - c0[:2]) ** 2).sum(-1))
  ang = np.arctan2(g[..., 1] - 36, g[..., 0] - 36)
  shapes = {
    "noise55": rng.random((N, N, N)) < 0.55,
    "noise62": rng.random((N, N, N)) < 0.62,
    "comb": ((g[..., 0].astype(int) % 4 < 2) | (g[..., 1] < 6)) & (g[..., 2] > 20) & (g[..., 2] < 50),
    "spiral": (np.abs(((rad - 2.2 * (ang + np.pi)) % 13.8) - 3) < 1.6) & (rad < 34),
    "rings": ((rad < 5) | ((rad > 9) & (rad < 13)) | ((rad > 20) & (rad < 25))),
  }
  for name, vol in shapes.items():
    vol = np.asfortranarray(vol)
    fg = np.argwhere(vol)
    for k in range(6):
      p = fg[rng.integers(len(fg))].astype(np.float32) if k else np.array([36, 36, 36], np.float32)
      if not vol[tuple(p.astype(int))]:
        vol[tuple(p.astype(int))] = True
      n = [[0, 0, 1], [1, 0, 0], [0.3, -0.2, 0.9], [1, 1, 0], [0.05, 0.02, 1], [0.7, 0.1, -0.4]][k]
      w = [[1, 1, 1], [32, 32, 40], [4, 4, 40]][k % 3]
      oimg, oct_ = oracle.section(vol, p, n, w)
      with xs.Volume(vol) as V:
        idx, val, ct = V.cross_section_sparse(p, n, w)
        idx2, val2, ct2 = V.cross_section_sparse(p, n, w)       # the visited set is all-zero again after a call
      oidx, oval = _sparse_of(oimg)
      order = np.argsort(idx)
      assert ct == oct_ and np.array_equal(idx[order], oidx) and np.array_equal(bits(val[order]), bits(oval)), (name, k)
      order2 = np.argsort(idx2)
      assert ct2 == ct and np.array_equal(idx2[order2], oidx) and np.array_equal(bits(val2[order2]), bits(oval)), (name, k, "second call")


def test_p2_pools_grow_and_rerun(xs, oracle):
  """The rarely taken paths of the CTA tiers' second half: a sample pool that is too small (doubles, the batch runs
  again) and first-touch tables that fill up (slots per sample x4, again) — forced through the dev knobs in a fresh
  process, results bit-identical to the oracle."""
  import subprocess, sys, os, textwrap
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  code = textwrap.dedent('''
    import sys, numpy as np
    sys.path.insert(0, r"%s/cross-section_b200"); sys.path.insert(0, r"%s/oracle")
    import xs3d, refapi
    from xs3d import synthetic
    vol, verts, tans = synthetic.make_neurite(256, seed=0)
    qp, qn = synthetic.draw_queries(verts, tans, 600, seed=3)
    cyl, axis = synthetic.make_cylinder()
    w = [32, 32, 40]
    orc = refapi.Oracle()
    with xs3d.Volume(vol) as V:
      a, c = V.cross_sectional_area_batch(qp, qn, w, return_contact=True)
      st = V.stats()
    oa, oc = orc.area_batch(vol, qp, qn, w)
    assert st["escalated_mid"] > 0
    assert np.array_equal(a.view(np.uint32), oa.view(np.uint32)) and np.array_equal(c, oc)
    with xs3d.Volume(cyl) as V:
      r = V.cross_sectional_area((256, 256, 256), axis, w, return_contact=True)
    o = orc.area(cyl, (256, 256, 256), axis, w)
    assert np.float32(r[0]).view(np.uint32) == np.float32(o[0]).view(np.uint32) and r[1] == o[1]
    print("ok")
  ''' % (root, root))
  for env in ({"XS3D_P2_ORDER_CAP": "300"}, {"XS3D_P2_FACTOR": "1"}, {"XS3D_P2_ORDER_CAP": "256", "XS3D_P2_FACTOR": "1"}):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (env, (r.stdout + r.stderr)[-2000:])


def test_twod_path_matches_reference(xs):
  """2-D images (xs3d/twod.py:73-174 via xs3d/__init__.py:79-80) on the device: the 600 golden cases of the reference's
  own twod.py — float64 areas bit for bit (the non-zero lengths are added in the reference's raster order), contacts —
  and the known answers of automated_test.py:393-418."""
  import golden_util as gu
  n = 0
  for i, img, p, v, w, area, contact, raises in gu.twod_cases():
    if raises:
      with pytest.raises(ValueError):
        xs.cross_sectional_area(img, p, v, w, return_contact=True)
    else:
      a, c = xs.cross_sectional_area(img, p, v, w, return_contact=True)
      assert a == area and c == contact, (i, img.shape, p, v, w, a, area, c, contact)
      if i % 7 == 0:    # C-ordered input and the contact-less return
        assert xs.cross_sectional_area(np.ascontiguousarray(img), p, v, w) == area
    n += 1
  assert n == 600
  labels = np.ones([3, 3], dtype=bool)
  assert xs.cross_sectional_area(labels, [1, 1], [0, 1]) == 3
  assert xs.cross_sectional_area(labels, [1, 1], [0, 1], [3, 3]) == 9
  assert xs.cross_sectional_area(labels, [1, 1], [1, 0], [5, 5]) == 15
  assert np.isclose(xs.cross_sectional_area(labels, [0, 0], [-1, 1], [1, 1]), 3 * np.sqrt(2))
  assert xs.cross_sectional_area(labels, [-1, 0], [-1, 1], [1, 1]) == 0
  big = np.ones([1500, 1100], dtype=bool)                      # a larger image: many CTAs, long union-find chains
  a, c = xs.cross_sectional_area(big, [700.25, 500.5], [0.3, 1.0], [4, 40], return_contact=True)
  import twod_oracle
  oa, oc = twod_oracle.cross_sectional_area_2d(big, [700.25, 500.5], [0.3, 1.0], [4, 40])
  assert a == oa and c == oc


def test_after_enqueue_hook(xs, oracle):
  """xs3d_volume_set_after_enqueue: the callback runs once per batch, behind the enqueue and before the host wait; work it
  puts on the volume's stream (here: a write over a scratch tensor and a marker) does not change or delay the results."""
  import torch
  from xs3d import synthetic
  vol, verts, tans = synthetic.make_neurite(96, seed=3, n_branches=3)
  pos, nrm = synthetic.draw_queries(verts, tans, 500, seed=2)
  w = [32, 32, 40]
  oa, oc = oracle.area_batch(vol, pos, nrm, w)
  st = torch.cuda.Stream()
  scratch = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
  calls = []
  with torch.cuda.stream(st):
    with xs.Volume(vol) as V:
      V.set_stream(st.cuda_stream)

      def hook():
        calls.append(len(calls))
        scratch.fill_(len(calls))            # enqueued on `st`, behind the batch
      V.set_after_enqueue(hook)
      a1, c1 = V.cross_sectional_area_batch(pos, nrm, w, return_contact=True)
      a2, c2 = V.cross_sectional_area_batch(pos[:7], nrm[:7], w, return_contact=True)
      V.set_after_enqueue(None)
      a3, c3 = V.cross_sectional_area_batch(pos, nrm, w, return_contact=True)
      st.synchronize()
  assert calls == [0, 1] and int(scratch[0]) == 2
  for a, c in ((a1, c1), (a3, c3)):
    assert np.array_equal(bits(a), bits(oa)) and np.array_equal(c, oc)
  assert np.array_equal(bits(a2), bits(oa[:7])) and np.array_equal(c2, oc[:7])
