"""CPU, world_size 2, gloo: the query-sharding host logic of xs3d.dist with the per-rank compute function
replaced by the oracle (no GPU here).  Checks partitioning, padding and the all-gather reassembly."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
  import torch.distributed as dist
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  from xs3d import dist as xdist
  from xs3d import synthetic
  import refapi
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    orc = refapi.Oracle()
    vol, verts, tans = synthetic.make_neurite(128, seed=0, n_branches=4)
    pos, nrm = synthetic.draw_queries(verts, tans, n, seed=1)
    compute = lambda p, m: orc.area_batch(vol, p, m, [32, 32, 40])
    a, c, (lo, hi) = xdist.sharded_area_batch(compute, pos, nrm)
    fa, fc = orc.area_batch(vol, pos, nrm, [32, 32, 40])
    ok = np.array_equal(a.view(np.uint32), fa.view(np.uint32)) and np.array_equal(c, fc)
    q.put((rank, ok, lo, hi))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 64])
def test_sharded_area_batch_gloo(n):
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import refapi
  if not refapi.have_oracle():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(60)
  res.sort()
  assert all(ok for _, ok, _, _ in res)
  assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == n


def _slice_worker(rank, world, port, q):
  import torch.distributed as dist
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  from xs3d import dist as xdist
  import refapi
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    orc = refapi.Oracle()
    rng = np.random.default_rng(8)
    lab = rng.integers(1, 1 << 40, size=(24, 20, 18)).astype(np.uint64)
    n = 11
    pos = rng.uniform(0, lab.shape, size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    slicer = lambda p, m: [orc.projection(lab, p[i], m[i], [4, 4, 40], True, 30.0) for i in range(len(p))]
    mine, (lo, hi) = xdist.sharded_slice_batch(slicer, pos, nrm)
    shapes = xdist.gather_slice_shapes(np.array([s.shape for s in mine], np.int64).reshape(-1, 2), n)
    full = slicer(pos, nrm)
    ok = (lo, hi) == xdist.shard_bounds(n, world, rank) and len(mine) == hi - lo
    ok = ok and all(np.array_equal(a, b) for a, b in zip(mine, full[lo:hi]))
    ok = ok and [tuple(s) for s in shapes] == [f.shape for f in full]
    q.put((rank, ok))
  finally:
    dist.destroy_process_group()


def test_sharded_slice_batch_gloo():
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import refapi
  if not refapi.have_oracle():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 33500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_slice_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(60)
  assert sorted(res) == [(0, True), (1, True)]


def _stream_worker(rank, world, port, q):
  """stream_volumes on gloo: rank 0 really bit-packs (xs3d.pack_bits is host code), a CPU stand-in plays the upload /
  ingest, the occupancy bytes travel by gloo broadcast, every rank checks what it holds at every step — both the
  plain loop (broadcast on the step's path) and the several-GPU loop (a broadcaster thread moves the staged buffers)."""
  import torch
  import torch.distributed as dist
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  import xs3d
  from xs3d import dist as xdist
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    shape = (20, 13, 9)
    h = [(s + 1) // 2 for s in shape]

    def occupancy(vol):
      pad = np.zeros([2 * x for x in h], bool)
      pad[: shape[0], : shape[1], : shape[2]] = vol
      out = np.zeros(h, np.uint8)
      for b in range(8):
        out |= pad[(b & 1)::2, ((b >> 1) & 1)::2, (b >> 2)::2].astype(np.uint8) << b
      return out.ravel(order="F")

    rng = np.random.default_rng(3)
    images = [rng.random(shape) < 0.4 for _ in range(5)]
    images = [np.asfortranarray(v) if i % 2 == 0 else np.ascontiguousarray(v) for i, v in enumerate(images)]   # both memory orders
    expected = [occupancy(v) for v in images]

    class CpuFeed:
      def __init__(self):
        self.staged = [None, None]
        self.cubes = torch.zeros(int(np.prod(h)), dtype=torch.uint8)
        self.loaded = 0
      def pack_and_stage(self, image, slot):
        bits, c_order = xs3d.pack_bits(image)
        self.staged[slot] = (bits.copy(), c_order)
      def commit(self, slot):
        bits, c_order = self.staged[slot]
        vox = np.unpackbits(bits, bitorder="little")[: int(np.prod(shape))].astype(bool)
        self.cubes[:] = torch.from_numpy(occupancy(vox.reshape(shape, order="C" if c_order else "F")))
        self.staged[slot] = None      # a second commit of the same slot without a new stage would fail
      def cubes_tensor(self):
        return self.cubes
      def mark_loaded(self):
        self.loaded += 1

    class CpuStagedFeed(CpuFeed):
      """adds what the several-GPU path of stream_volumes needs: spare occupancy buffers per slot, filled by the
      'ingest' on the source rank and by the broadcast elsewhere, swapped in at commit"""
      def __init__(self):
        super().__init__()
        self.spare = [torch.zeros_like(self.cubes), torch.zeros_like(self.cubes)]
        self.marked = [False, False]
        self.bound = 0
      def pack_and_stage(self, image, slot):
        super().pack_and_stage(image, slot)
        bits, c_order = self.staged[slot]
        vox = np.unpackbits(bits, bitorder="little")[: int(np.prod(shape))].astype(bool)
        self.spare[slot][:] = torch.from_numpy(occupancy(vox.reshape(shape, order="C" if c_order else "F")))
      def bind_thread(self):
        self.bound += 1
      def bcast_ctx(self):
        import contextlib
        return contextlib.nullcontext()
      def staged_tensor(self, slot):
        return self.spare[slot]
      def wait_staged(self, slot):
        assert self.staged[slot] is not None
      def mark_staged(self, slot):
        self.marked[slot] = True
      def commit_swap(self, slot):
        assert self.marked[slot]
        self.marked[slot] = False
        self.cubes, self.spare[slot] = self.spare[slot], self.cubes

    sfeed = CpuStagedFeed()
    res2 = xdist.stream_volumes(None, images if rank == 0 else None, 5, lambda k: sfeed.cubes.numpy().copy(), feed=sfeed)
    ok2 = len(res2) == 5 and all(np.array_equal(r, e) for r, e in zip(res2, expected)) and sfeed.bound == 1
    # the sweep may run collectives on the caller's group while the broadcaster thread works (on its own communicator)
    sfeed2 = CpuStagedFeed()

    def sweep_with_collective(k):
      t = torch.tensor([float(sfeed2.cubes.sum())], dtype=torch.float64)
      dist.all_reduce(t)
      return float(t.item())
    res3 = xdist.stream_volumes(None, images if rank == 0 else None, 5, sweep_with_collective, feed=sfeed2)
    ok2 = ok2 and all(r == 2.0 * float(e.sum()) for r, e in zip(res3, expected))
    # a failure of the source rank while producing image 2 reaches the other rank as an exception at that step
    sfeed3 = CpuStagedFeed()
    bad = [images[0], images[1], np.zeros(shape, np.float32), images[3]]
    done = []
    try:
      xdist.stream_volumes(None, bad if rank == 0 else None, 4, done.append, feed=sfeed3)
      ok2 = False
    except ValueError:
      ok2 = ok2 and rank == 0 and done == [0, 1]
    except RuntimeError:
      ok2 = ok2 and rank == 1 and done in ([0, 1], [0, 1, 2])
    feed = CpuFeed()
    seen = []
    res = xdist.stream_volumes(None, images if rank == 0 else None, 5, lambda k: feed.cubes.numpy().copy(),
                               before_step=seen.append, feed=feed)
    ok = len(res) == 5 and all(np.array_equal(r, e) for r, e in zip(res, expected)) and seen == list(range(5))
    ok = ok and ok2 and feed.loaded == (0 if rank == 0 else 5)
    # a failure while packing surfaces on the source rank at that step instead of hanging the helper thread
    if rank == 0:
      try:
        xdist.stream_volumes(None, [images[0], np.zeros(shape, np.float32)], 2, lambda k: k, feed=CpuFeed(), group=dist.new_group([0]))
        ok = False
      except ValueError:
        pass
    else:
      dist.new_group([0])
    q.put((rank, ok))
  finally:
    dist.destroy_process_group()


def test_stream_volumes_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 31500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(60)
  assert sorted(res) == [(0, True), (1, True)]


def _shared_worker(rank, world, port, q):
  """stream_shared_volumes on gloo: every rank really bit-packs ITS part of every image (xs3d_pack_bits is host code),
  the parts travel by gloo all-gather, a CPU stand-in plays upload / ingest; every rank checks what it holds at every
  step, a sweep runs a collective beside the helper thread, and a rank that cannot pack its part makes every rank raise."""
  import ctypes as C
  import torch
  import torch.distributed as dist
  sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
  from xs3d import _abi
  from xs3d import dist as xdist
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  try:
    shape = (64, 40, 33)          # 84 480 voxels: parts of 65 536 and 18 944
    voxels = int(np.prod(shape))
    h = [(s + 1) // 2 for s in shape]

    def occupancy(vol):
      pad = np.zeros([2 * x for x in h], bool)
      pad[: shape[0], : shape[1], : shape[2]] = vol
      out = np.zeros(h, np.uint8)
      for b in range(8):
        out |= pad[(b & 1)::2, ((b >> 1) & 1)::2, (b >> 2)::2].astype(np.uint8) << b
      return out.ravel(order="F")

    rng = np.random.default_rng(5)
    images = [rng.random(shape) < 0.4 for _ in range(5)]
    images = [np.asfortranarray(v) if i % 2 == 0 else np.ascontiguousarray(v) for i, v in enumerate(images)]
    expected = [occupancy(v) for v in images]
    part, lo, hi = xdist.part_bounds(voxels, world, rank)
    assert part % 32768 == 0 and (lo, hi) == ((0, 65536) if rank == 0 else (65536, voxels))
    lib = _abi.lib()

    class CpuPartFeed:
      def __init__(self):
        self.cubes = torch.zeros(int(np.prod(h)), dtype=torch.uint8)
        self.spare = [torch.zeros_like(self.cubes), torch.zeros_like(self.cubes)]
        self.bits = None
        self.status = [torch.zeros(world, dtype=torch.uint8) for _ in range(4)]
        self.ready = [False, False]
        self.bound = 0
      def prepare_parts(self, world_size):
        self.bits = [torch.zeros(world_size * part // 8, dtype=torch.uint8) for _ in range(2)]
      def bind_thread(self):
        self.bound += 1
      def bcast_ctx(self):
        import contextlib
        return contextlib.nullcontext()
      def pack_part(self, image, slot, plo, phi):
        if image.dtype != bool:
          raise ValueError("A boolean image is required")
        flat = image.view(np.uint8).ravel(order="K")
        dst = self.bits[slot].numpy()
        if phi > plo:
          rc = lib.xs3d_pack_bits(C.c_void_p(flat.ctypes.data + plo), phi - plo, C.c_void_p(dst.ctypes.data + plo // 8))
          assert rc == 0
        return bool(image.flags.c_contiguous and not image.flags.f_contiguous)
      def wait_bits(self, slot):
        pass
      def bits_tensor(self, slot):
        return self.bits[slot]
      def gather_status(self, slot, value, r, group):
        t = self.status[slot]
        t[r] = int(value)
        dist.all_gather_into_tensor(t, t[r: r + 1].clone(), group=group)
      def read_gathered_status(self, slot, wait=True):
        return int(self.status[slot].max())
      def ingest(self, slot, c_order):
        vox = np.unpackbits(self.bits[slot].numpy(), bitorder="little")[:voxels].astype(bool)
        self.spare[slot][:] = torch.from_numpy(occupancy(vox.reshape(shape, order="C" if c_order else "F")))
        self.ready[slot] = True
      def commit_swap(self, slot):
        assert self.ready[slot]
        self.ready[slot] = False
        self.cubes, self.spare[slot] = self.spare[slot], self.cubes

    feed = CpuPartFeed()
    seen = []
    res = xdist.stream_shared_volumes(None, images, 5, lambda k: feed.cubes.numpy().copy(), before_step=seen.append, feed=feed)
    ok = len(res) == 5 and all(np.array_equal(r, e) for r, e in zip(res, expected)) and seen == list(range(5)) and feed.bound == 1
    feed2 = CpuPartFeed()

    def sweep_with_collective(k):
      t = torch.tensor([float(feed2.cubes.sum())], dtype=torch.float64)
      dist.all_reduce(t)
      return float(t.item())
    res2 = xdist.stream_shared_volumes(None, images, 5, sweep_with_collective, feed=feed2)
    ok = ok and all(r == 2.0 * float(e.sum()) for r, e in zip(res2, expected))
    # rank 1 cannot pack its part of image 2: it raises its own error, rank 0 a RuntimeError, both at step 2
    feed3 = CpuPartFeed()
    bad = list(images[:4])
    if rank == 1:
      bad[2] = np.zeros(shape, np.float32)
    done = []
    try:
      xdist.stream_shared_volumes(None, bad, 4, done.append, feed=feed3)
      ok = False
    except ValueError:
      ok = ok and rank == 1 and done == [0, 1]
    except RuntimeError:
      ok = ok and rank == 0 and done == [0, 1]
    q.put((rank, ok))
  finally:
    dist.destroy_process_group()


def test_stream_shared_volumes_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 35500 + (os.getpid() % 2000)
  procs = [ctx.Process(target=_shared_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(60)
  assert sorted(res) == [(0, True), (1, True)]
