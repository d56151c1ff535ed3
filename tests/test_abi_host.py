"""CPU: the C-ABI library loads, exports every symbol the header declares, the Python host layer mirrors
the reference's argument validation, and compute calls fail loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  src = open(os.path.join(ROOT, "include", "xs3d_b200.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(xs3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
  from xs3d import _abi
  lib = C.CDLL(_abi.LIB_PATH)
  syms = header_symbols()
  assert len(syms) >= 20
  for s in syms:
    assert hasattr(lib, s), f"{s} declared in include/xs3d_b200.h but not exported"
  assert sorted(_abi.SYMBOLS) == syms
  assert lib.xs3d_version() >= 100


def test_no_cpu_fallback_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  import xs3d
  vol = np.ones((4, 4, 4), dtype=bool, order="F")
  with pytest.raises(xs3d.XS3DError):
    xs3d.cross_sectional_area(vol, [1, 1, 1], [0, 0, 1])
  with pytest.raises(xs3d.XS3DError):
    xs3d.Volume(vol)
  with pytest.raises(xs3d.XS3DError):
    xs3d.cross_section(vol, [1, 1, 1], [0, 0, 1])
  with pytest.raises(xs3d.XS3DError):
    xs3d.slice(vol.astype(np.uint8), [1, 1, 1], [0, 0, 1])
  with pytest.raises(xs3d.XS3DError):
    xs3d.cross_sectional_area(np.ones((3, 3), dtype=bool), [1, 1], [0, 1])      # the 2-D path is device code too


def test_input_validation_matches_reference():
  """automated_test.py:344-390 and :332-342 (the ValueErrors are raised before the device is touched)."""
  import xs3d
  labels = np.arange(9, dtype=np.uint8).reshape([3, 3, 1], order="F")
  for fn in (xs3d.cross_sectional_area, xs3d.cross_section):
    with pytest.raises(ValueError):
      fn(labels, [0, 0, 0], [0, 0, 1])
    ok = np.ones([3, 3, 1], dtype=bool, order="F")
    with pytest.raises(ValueError):
      fn(ok, [0, 0, 0], [0, 0, 0])
    for bad in ([-1, 1, 1], [1, -1, 1], [1, 1, -1], [-1, -1, -1], [0, 1, 1]):
      with pytest.raises(ValueError):
        fn(ok, [0, 0, 0], [0, 0, 1], bad)
    with pytest.raises(ValueError):
      fn(np.ones([3, 3, 3, 3], dtype=bool, order="F"), [0, 0, 0], [0, 0, 1], [1, 1, 1])
  # 2-D images (xs3d/__init__.py:59-80): same checks, before the device is touched
  img2 = np.ones([3, 3], dtype=bool)
  with pytest.raises(ValueError):
    xs3d.cross_sectional_area(img2.astype(np.uint8), [1, 1], [0, 1])
  with pytest.raises(ValueError):
    xs3d.cross_sectional_area(img2, [1, 1], [0, 0])
  with pytest.raises(ValueError):
    xs3d.cross_sectional_area(img2, [1, 1], [0, 1], [1, -1])
  with pytest.raises(ValueError):
    xs3d.cross_sectional_area(img2, [1, 1, 1], [0, 1, 0], [1, 1, 1])
  with pytest.raises(ValueError):
    xs3d.slice(labels, [0, 0, 0], [0, 0, 0])
  with pytest.raises(ValueError):
    xs3d.slice(np.ones([3, 3, 3, 3], dtype=bool, order="F"), [0, 0, 0], [0, 0, 1])
  with pytest.raises(AssertionError):
    xs3d.slice(labels, [0, 0, 0], [0, 0, 1], crop=-1.0)


def test_signatures_match_reference():
  import inspect
  import xs3d
  assert list(inspect.signature(xs3d.cross_sectional_area).parameters) == ["binimg", "pos", "normal", "anisotropy", "return_contact", "slow_method", "use_persistent_data"]
  assert list(inspect.signature(xs3d.cross_section).parameters) == ["binimg", "pos", "normal", "anisotropy", "return_contact", "method"]
  assert list(inspect.signature(xs3d.slice).parameters) == ["labels", "pos", "normal", "anisotropy", "standardize_basis", "crop"]
  assert list(inspect.signature(xs3d.set_shape).parameters) == ["image"]
  assert list(inspect.signature(xs3d.clear_shape).parameters) == []


def test_shard_bounds():
  from xs3d.dist import shard_bounds
  for n in (0, 1, 7, 10000, 10001):
    for world in (1, 2, 3, 8):
      chunks = [shard_bounds(n, world, r) for r in range(world)]
      assert chunks[0][0] == 0 and chunks[-1][1] == n
      assert all(chunks[i][1] == chunks[i + 1][0] for i in range(world - 1))
      sizes = [hi - lo for lo, hi in chunks]
      assert max(sizes) - min(sizes) <= 1


def test_skeleton_tangents():
  import xs3d
  verts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [2, 1, 0], [9, 9, 9]], float)
  edges = np.array([[0, 1], [1, 2], [2, 3]])
  t = xs3d.skeleton_tangents(verts, edges)
  assert t.dtype == np.float32 and t.shape == (5, 3)
  assert np.allclose(np.abs(t[0]), [1, 0, 0]) and np.allclose(np.abs(t[1]), [1, 0, 0])
  assert np.allclose(np.abs(t[2]), np.array([1, 1, 0]) / np.sqrt(2), atol=1e-6)
  assert np.allclose(np.abs(t[3]), [0, 1, 0]) and np.allclose(t[4], [0, 0, 1])
  assert np.allclose(np.linalg.norm(t, axis=1), 1, atol=1e-6)


def test_slice_batch_sequence_is_lazy_views():
  """Volume.slice_batch returns a SliceBatch: a Sequence of Fortran-ordered views into one [n, pitch] buffer."""
  import numpy as np
  import xs3d
  buf = np.arange(36, dtype=np.uint16).reshape(3, 12)
  b = xs3d.SliceBatch(buf, np.array([[2, 3], [3, 4], [1, 1]]))
  assert len(b) == 3 and b[0].shape == (2, 3) and b[-1].shape == (1, 1)
  assert np.array_equal(b[0], np.array([[0, 2, 4], [1, 3, 5]], np.uint16)) and b[1][2, 3] == 12 + 11
  assert [x.shape for x in b] == [(2, 3), (3, 4), (1, 1)] and [x.shape for x in b[1:]] == [(3, 4), (1, 1)]
  assert np.shares_memory(b[1], buf)
  import pytest
  with pytest.raises(IndexError):
    b[3]


def test_host_bit_packing_matches_numpy():
  """xs_pack_bits (csrc/xs_hostpack.cpp; the host half of the bit-packed volume upload): bit i of the stream is
  byte i != 0, for single-threaded and pooled sizes, ragged tails, chunked calls and byte values other than 0/1."""
  from xs3d import _abi
  lib = C.CDLL(_abi.LIB_PATH)
  lib.xs_pack_bits.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
  lib.xs_pack_bits.restype = C.c_int
  rng = np.random.default_rng(5)
  for n in (1, 7, 8, 63, 64, 65, 1000, 4099, (1 << 20) + 13, (3 << 20) + 5, 9 << 20):
    img = (rng.random(n) < 0.3).astype(np.uint8) * rng.integers(1, 256, n, dtype=np.uint8)
    exp = np.packbits(img != 0, bitorder="little")
    out = np.full((n + 7) // 8 + 8, 0xAA, np.uint8)
    threads = lib.xs_pack_bits(img.ctypes.data, 0, n, out.ctypes.data)
    assert threads >= 1 and np.array_equal(out[: len(exp)], exp), n
    assert np.all(out[len(exp):] == 0xAA), n            # nothing written past the stream
    # the same stream built a quarter at a time (how the upload overlaps packing with the transfers)
    out2 = np.zeros_like(out)
    chunk = ((n // 4 + 63) // 64) * 64
    for lo in range(0, n, max(chunk, 64)):
      lib.xs_pack_bits(img.ctypes.data, lo, min(n, lo + max(chunk, 64)), out2.ctypes.data)
    assert np.array_equal(out2[: len(exp)], exp), n


def test_host_bit_packing_reports_parts_in_order():
  """xs_pack_bits_parts: one pooled job whose parts are reported in order on the calling thread, each one complete
  (and byte aligned in the bit stream) when its callback runs — the upload enqueues the part's copy there."""
  from xs3d import _abi
  lib = C.CDLL(_abi.LIB_PATH)
  CB = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_size_t)
  lib.xs_pack_bits_parts.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, CB, C.c_void_p]
  lib.xs_pack_bits_parts.restype = C.c_int
  rng = np.random.default_rng(6)
  for n, parts in (((5 << 20) + 77, 8), (3 << 20, 3), (1000, 8), ((2 << 20) + 1, 64)):
    img = (rng.random(n) < 0.5).astype(np.uint8)
    exp = np.packbits(img, bitorder="little")
    out = np.zeros((n + 7) // 8, np.uint8)
    seen = []

    def done(ctx, lo, hi):
      assert lo % 8 == 0 and np.array_equal(out[lo // 8:(hi + 7) // 8], exp[lo // 8:(hi + 7) // 8])
      seen.append((lo, hi))

    assert lib.xs_pack_bits_parts(img.ctypes.data, 0, n, out.ctypes.data, parts, CB(done), None) >= 1
    assert seen[0][0] == 0 and seen[-1][1] == n and all(a[1] == b[0] for a, b in zip(seen, seen[1:])), seen
    assert 1 <= len(seen) <= min(parts, 16) and np.array_equal(out, exp)


def test_pack_bits_host_api():
  """xs3d.pack_bits (C ABI xs3d_pack_bits): the bit stream of an image in its own memory order, for Volume.load_packed."""
  import xs3d
  rng = np.random.default_rng(8)
  vol = rng.random((37, 41, 53)) < 0.3
  for arr, order in ((np.asfortranarray(vol), "F"), (np.ascontiguousarray(vol), "C"), (np.asfortranarray(vol).view(np.uint8) * 7, "F")):
    bits, c_order = xs3d.pack_bits(arr)
    assert c_order == (order == "C") and bits.dtype == np.uint8
    assert np.array_equal(bits, np.packbits(arr.ravel(order=order) != 0, bitorder="little"))
  out = np.zeros((vol.size + 7) // 8 + 3, np.uint8)
  bits, _ = xs3d.pack_bits(np.asfortranarray(vol), out=out)
  assert bits is out and np.array_equal(out[:-3], np.packbits(vol.ravel(order="F"), bitorder="little"))
  with pytest.raises(ValueError):
    xs3d.pack_bits(vol.astype(np.float32))
  with pytest.raises(ValueError):
    xs3d.pack_bits(vol, out=np.zeros(5, np.uint8))
