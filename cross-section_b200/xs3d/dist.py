"""Query sharding over ranks (one process per GPU) — SURVEY.md §8(e).

Cross-section queries are independent and read-only on the volume, so the multi-GPU path has no
data-path collective: every rank holds the ingested volume and works on a contiguous slice of the
query list.  torch.distributed is used only as plumbing:

  * broadcast_volume(): rank `src` ingests the image, the 2x2x2 occupancy bytes (V/8 bytes, e.g.
    16 MiB for 512^3) are broadcast once over NCCL/NVLink into every rank's resident Volume;
  * shard_bounds() / gather_results(): contiguous partition of n queries and the (small: 5 B/query)
    all-gather of areas and contact bits;
  * stream_volumes(): a sequence of host volumes through the ranks' resident Volumes — rank `src` bit-packs and
    uploads volume k+1 on a helper thread while everybody sweeps volume k; the occupancy bytes are broadcast
    every step by a helper thread on a communicator of its own (broadcast_group), followed by a status byte
    that turns a failure of the source rank into an exception on every rank (the end-to-end loop of bench.py);
  * stream_shared_volumes(): the same sequence when every rank can read the host images (shared memory, or a replica
    each): every rank bit-packs and uploads 1/world of each image, the bit stream is all-gathered over NVLink and each
    GPU ingests it — the host-memory-bound packing is spread over the cores of all ranks, not the source rank's;
  * broadcast_labels() / sharded_slice_batch(): the same for xs3d.slice — the label volume is broadcast once,
    the planes are partitioned (BASELINE config 5).

The same code runs on the gloo backend with CPU tensors (tests/test_dist_gloo.py), where the
per-rank compute function is injected by the test.
"""
import numpy as np


def shard_bounds(n, world_size, rank):
  """[lo, hi) of rank's contiguous chunk; chunks differ in size by at most one query."""
  base, rem = divmod(int(n), int(world_size))
  lo = rank * base + min(rank, rem)
  hi = lo + base + (1 if rank < rem else 0)
  return lo, hi


def broadcast_volume(volume, binimg=None, src=0, group=None):
  """Make `volume` (an xs3d.Volume of the right shape on this rank's GPU) hold the image that rank
  `src` passes as `binimg`: src ingests, everyone receives the occupancy bytes by NCCL broadcast."""
  import ctypes
  import torch
  import torch.distributed as dist
  rank = dist.get_rank(group)
  if rank == src:
    if binimg is None:
      raise ValueError("the source rank must pass the image")
    volume.load(binimg)
  ptr, nbytes = volume.cubes()
  if nbytes == 0:
    volume.mark_loaded()
    return volume
  # zero-copy torch view of the library's device buffer (torch is plumbing only)
  dev = torch.device("cuda", volume.device)

  class _Holder:
    pass

  holder = _Holder()
  holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
  t = torch.as_tensor(holder, device=dev)
  dist.broadcast(t, src=src, group=group)
  torch.cuda.synchronize(dev)
  volume.mark_loaded()
  return volume


def _device_bytes(ptr, nbytes, device):
  import torch
  holder = type("_Holder", (), {})()
  holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
  return torch.as_tensor(holder, device=torch.device("cuda", device))


def broadcast_labels(volume, labels=None, src=0, group=None, chunk_bytes=1 << 30):
  """Make every rank's `volume` hold the label volume rank `src` passes (BASELINE config 5: an 8 GiB uint64
  segmentation is uploaded over PCIe once, by one rank, and travels to the others over NVLink): src uploads, then the
  device buffer is NCCL-broadcast in `chunk_bytes` pieces.  dtype and memory order travel with it."""
  import torch
  import torch.distributed as dist
  rank = dist.get_rank(group)
  gsrc = dist.get_global_rank(group, src) if group is not None else src
  meta = [None]
  if rank == src:
    if labels is None:
      raise ValueError("the source rank must pass the labels")
    if not labels.flags.f_contiguous:
      labels = np.ascontiguousarray(labels)
    volume.load_labels(labels)
    meta = [(labels.dtype.str, bool(labels.flags.c_contiguous))]
  dist.broadcast_object_list(meta, src=gsrc, group=group)
  dtype, c_order = np.dtype(meta[0][0]), meta[0][1]
  ptr, nbytes = volume.labels_buffer(dtype, c_order)          # (on src: the buffer load_labels just filled)
  if dist.get_backend(group) != "nccl":
    raise RuntimeError("broadcast_labels moves device buffers: it needs the nccl backend")
  t = _device_bytes(ptr, nbytes, volume.device)
  torch.cuda.synchronize(volume.device)
  for lo in range(0, nbytes, int(chunk_bytes)):
    dist.broadcast(t[lo: lo + int(chunk_bytes)], src=gsrc, group=group)
  torch.cuda.synchronize(volume.device)
  return volume


def sharded_slice_batch(slicer, positions, normals, group=None):
  """Run `slicer(pos_chunk, normal_chunk) -> sequence of 2-D arrays` (normally Volume.slice_batch with the anisotropy /
  crop bound) on this rank's contiguous share of the planes.  Returns (cut-outs of the share, (lo, hi)): slices are
  consumed where they are made — gathering 10^4 ragged images on every rank would cost more than making them — so there
  is no collective here at all; `gather_slice_shapes` exchanges just the shapes when a rank needs the global picture."""
  import torch.distributed as dist
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
  nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
  if pos.shape != nrm.shape:
    raise ValueError("positions and normals must have the same length")
  lo, hi = shard_bounds(pos.shape[0], world, rank)
  return slicer(pos[lo:hi], nrm[lo:hi]), (lo, hi)


def gather_slice_shapes(shapes, n, group=None):
  """All-gather of the [k,2] int64 shapes of every rank's share into the [n,2] array of the whole plane list."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  backend = dist.get_backend(group)
  dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
  kmax = shard_bounds(n, world, 0)[1]
  buf = torch.zeros((kmax, 2), dtype=torch.int64, device=dev)
  lo, hi = shard_bounds(n, world, rank)
  if hi > lo:
    buf[: hi - lo] = torch.from_numpy(np.ascontiguousarray(shapes, dtype=np.int64).reshape(-1, 2)).to(dev)
  parts = [torch.empty_like(buf) for _ in range(world)]
  dist.all_gather(parts, buf, group=group)
  out = np.empty((n, 2), dtype=np.int64)
  for r in range(world):
    rlo, rhi = shard_bounds(n, world, r)
    out[rlo:rhi] = parts[r][: rhi - rlo].cpu().numpy()
  return out


def sharded_area_batch(compute, positions, normals, group=None, gather=True):
  """Run `compute(pos_chunk, normal_chunk) -> (area float32 [k], contact uint8 [k])` on this rank's
  contiguous slice of the queries and (optionally) all-gather the results so every rank returns the
  full (area [n], contact [n]).  `compute` is normally Volume.cross_sectional_area_batch with
  return_contact=True."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
  nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
  n = pos.shape[0]
  lo, hi = shard_bounds(n, world, rank)
  area, contact = compute(pos[lo:hi], nrm[lo:hi])
  area = np.ascontiguousarray(area, dtype=np.float32)
  contact = np.ascontiguousarray(contact, dtype=np.uint8)
  if not gather:
    return area, contact, (lo, hi)
  backend = dist.get_backend(group)
  dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
  # pad to the largest chunk so a plain all_gather works on every backend
  kmax = shard_bounds(n, world, 0)[1]
  buf = torch.zeros(kmax * 5, dtype=torch.uint8, device=dev)
  k = hi - lo
  if k:
    buf[: 4 * k] = torch.from_numpy(area.view(np.uint8)).to(dev)
    buf[4 * kmax: 4 * kmax + k] = torch.from_numpy(contact).to(dev)
  parts = [torch.empty_like(buf) for _ in range(world)]
  dist.all_gather(parts, buf, group=group)
  out_a = np.empty(n, dtype=np.float32)
  out_c = np.empty(n, dtype=np.uint8)
  for r in range(world):
    rlo, rhi = shard_bounds(n, world, r)
    rk = rhi - rlo
    p = parts[r].cpu().numpy()
    out_a[rlo:rhi] = p[: 4 * rk].view(np.float32)
    out_c[rlo:rhi] = p[4 * kmax: 4 * kmax + rk]
  return out_a, out_c, (lo, hi)


class VolumeFeed:
  """What stream_volumes does to a real (CUDA) Volume: bit-pack the image with the library's host thread pool while the
  packed parts go up and are ingested on the volume's upload stream (Volume.pack_and_stage); later swap the result in."""

  def __init__(self, volume):
    self.volume = volume

  def pack_and_stage(self, image, slot):
    self.volume.pack_and_stage(image, slot)

  def commit(self, slot):
    import torch
    self.volume.commit_staged(slot)
    # the swap is ordered on the volume's launching stream; a broadcast issued on torch's current stream must come after
    # it, so unless the two are the same stream (Volume.set_stream) wait for the staged work here
    if getattr(self.volume, "_stream", None) != torch.cuda.current_stream(self.volume.device).cuda_stream:
      torch.cuda.synchronize(self.volume.device)

  def cubes_tensor(self):
    import torch
    ptr, nbytes = self.volume.cubes()
    holder = type("_Holder", (), {})()
    holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(holder, device=torch.device("cuda", self.volume.device))

  def mark_loaded(self):
    self.volume.mark_loaded()

  # -- broadcast of the STAGED occupancy bytes on a side stream (beside the sweep of the previous image)
  def bcast_stream(self):
    import torch
    if getattr(self, "_bstream", None) is None:
      self._bstream = torch.cuda.Stream(device=torch.device("cuda", self.volume.device))
    return self._bstream

  def bind_thread(self):
    """Called first thing on the broadcaster thread."""
    import torch
    torch.cuda.set_device(self.volume.device)

  def bcast_ctx(self):
    """Context in which the broadcast of a staged buffer is issued (the side stream)."""
    import torch
    return torch.cuda.stream(self.bcast_stream())

  def staged_tensor(self, slot):
    import torch
    ptr, nbytes = self.volume.staged_cubes(slot)
    holder = type("_Holder", (), {})()
    holder.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
    return torch.as_tensor(holder, device=torch.device("cuda", self.volume.device))

  def wait_staged(self, slot):
    self.volume.wait_staged(slot, self.bcast_stream().cuda_stream)

  def mark_staged(self, slot):
    self.volume.mark_staged(slot, self.bcast_stream().cuda_stream)

  def commit_swap(self, slot):
    self.volume.commit_staged(slot)

  # -- status byte behind every payload (stream_volumes): device byte broadcast on the side stream, copied to pinned host
  #    memory behind it, an event says when it can be read.  Four status slots (step & 3): the broadcast of step k+2 can
  #    be issued before the main thread has read the status of step k, that of step k+4 cannot
  def _status_bufs(self):
    import torch
    if getattr(self, "_st_dev", None) is None:
      dev = torch.device("cuda", self.volume.device)
      self._st_dev = [torch.zeros(1, dtype=torch.uint8, device=dev) for _ in range(4)]
      self._st_host = [torch.zeros(1, dtype=torch.uint8).pin_memory() for _ in range(4)]
      self._st_ev = [torch.cuda.Event() for _ in range(4)]
      self._st_issued = [False] * 4
    return self._st_dev, self._st_host, self._st_ev

  def bcast_status(self, slot, value, src, group):
    """Called inside bcast_ctx() (the side stream is current)."""
    import torch.distributed as dist
    d, h, ev = self._status_bufs()
    if value is not None:
      d[slot].fill_(int(value))
    dist.broadcast(d[slot], src=src, group=group)
    h[slot].copy_(d[slot], non_blocking=True)
    ev[slot].record()
    self._st_issued[slot] = True

  def read_status(self, slot, wait=True):
    d, h, ev = self._status_bufs()
    if not self._st_issued[slot]:
      return 0
    if wait:
      # wait without burning a core: the packing pool of the source rank needs them (eight broadcaster threads spinning
      # in a synchronize cost the 8-GPU end-to-end step 0.1 ms); nobody waits for this thread within a step
      import time
      while not ev[slot].query():
        time.sleep(20e-6)
    elif not ev[slot].query():
      return 0
    return int(h[slot][0])


def part_bounds(voxels, world_size, rank, align=32768):
  """(part, lo, hi): the image is cut into `world_size` parts of `part` voxels (a multiple of `align`, so that every part
  of the bit stream starts on a 4 KiB boundary); rank's part is [lo, hi) — empty, or shorter, at the end of the image."""
  voxels, world_size = int(voxels), int(world_size)
  per = -(-voxels // world_size)
  part = -(-per // align) * align
  lo = min(rank * part, voxels)
  return part, lo, min(lo + part, voxels)


class SharedVolumeFeed(VolumeFeed):
  """What stream_shared_volumes does to a real (CUDA) Volume: pack + upload this rank's part, all-gather the bit stream
  on the side stream, ingest it there."""

  def prepare_parts(self, world_size):
    voxels = int(np.prod(self.volume.shape))
    part, _, _ = part_bounds(voxels, world_size, 0)
    self._part_bytes = part // 8
    self._bits = []
    for slot in (0, 1):
      ptr, _ = self.volume.staged_bits(slot, world_size * self._part_bytes)
      self.volume.staged_cubes(slot)
      self._bits.append(_device_bytes(ptr, world_size * self._part_bytes, self.volume.device))

  def pack_part(self, image, slot, lo, hi):
    return self.volume.pack_and_stage_part(image, slot, lo, hi)

  def bits_tensor(self, slot):
    return self._bits[slot]

  def wait_bits(self, slot):
    self.volume.wait_staged_bits(slot, self.bcast_stream().cuda_stream)

  def ingest(self, slot, c_order):
    self.volume.ingest_staged(slot, c_order, self.bcast_stream().cuda_stream)

  # status bytes, one per rank, all-gathered behind every payload (same four-slot scheme as VolumeFeed's)
  def gather_status(self, slot, value, rank, group):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if getattr(self, "_gs_dev", None) is None or self._gs_dev[0].numel() != world:
      dev = torch.device("cuda", self.volume.device)
      self._gs_dev = [torch.zeros(world, dtype=torch.uint8, device=dev) for _ in range(4)]
      self._gs_host = [torch.zeros(world, dtype=torch.uint8).pin_memory() for _ in range(4)]
      self._gs_ev = [torch.cuda.Event() for _ in range(4)]
      self._gs_issued = [False] * 4
    d = self._gs_dev[slot]
    d[rank: rank + 1].fill_(int(value))
    dist.all_gather_into_tensor(d, d[rank: rank + 1], group=group)
    self._gs_host[slot].copy_(d, non_blocking=True)
    self._gs_ev[slot].record()
    self._gs_issued[slot] = True

  def read_gathered_status(self, slot, wait=True):
    if getattr(self, "_gs_dev", None) is None or not self._gs_issued[slot]:
      return 0
    if wait:
      import time
      while not self._gs_ev[slot].query():
        time.sleep(20e-6)
    elif not self._gs_ev[slot].query():
      return 0
    return int(self._gs_host[slot].max())


class _HostStatus:
  """Status byte that travels behind every staged payload (stream_volumes): generic form for feeds whose tensors live
  in host memory (the gloo tests).  VolumeFeed has the CUDA form (side stream, pinned copy, event)."""

  def __init__(self):
    import torch
    self.t = [torch.zeros(1, dtype=torch.uint8) for _ in range(4)]

  def bcast_status(self, slot, value, src, group):
    import torch.distributed as dist
    if value is not None:
      self.t[slot][0] = int(value)
    dist.broadcast(self.t[slot], src=src, group=group)

  def read_status(self, slot, wait=True):
    return int(self.t[slot][0])


_bcast_groups = {}


def broadcast_group(group=None):
  """The process group stream_volumes' broadcaster threads use: a communicator of its own over the ranks of `group`,
  so that the broadcasts issued from the helper thread never share a communicator with the collectives the caller's
  sweep runs on `group` (two threads enqueueing on one NCCL communicator in rank-dependent order deadlock).
  Collective: the first call for a given `group` must be made by every rank of the default group at the same point
  (torch.distributed.new_group's rule); later calls return the cached group."""
  import torch.distributed as dist
  key = id(group) if group is not None else None
  if key not in _bcast_groups:
    import os
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
    backend = dist.get_backend(group)
    opts = None
    if backend == "nccl" and hasattr(dist, "ProcessGroupNCCL"):
      # The broadcast of image k+1 is enqueued a step ahead and its kernel WAITS on the GPU for the source rank's data while
      # the sweep of image k runs: with NCCL's default channel count that is a dozen or more resident CTAs taking shared
      # memory and registers from the query tiers for most of the sweep.  16 MiB a step ahead needs no bandwidth to speak
      # of, so this communicator is limited to a few CTAs (ncclConfig max_ctas).
      try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(os.environ.get("XS3D_BCAST_CTAS", "2"))
        opts.config.min_ctas = 1
      except Exception:
        opts = None
    _bcast_groups[key] = dist.new_group(ranks=ranks, backend=backend, pg_options=opts) if opts is not None else dist.new_group(ranks=ranks, backend=backend)
  return _bcast_groups[key]


def stream_volumes(volume, images, n_steps, sweep, src=0, group=None, before_step=None, feed=None):
  """Run `sweep(k)` for k = 0 .. n_steps-1 with `volume` holding images[k] (only rank `src` passes `images`; the others
  may pass None).  Rank `src` packs + uploads image k+1 on a helper thread (double-buffered) while step k runs; per
  step the staged occupancy bytes are swapped in on `src`, broadcast when there is more than one rank, and
  every rank calls sweep(k) — normally its own shard of the queries against `volume`.  Returns the list of sweep
  results.  `before_step(k)` runs at the top of each step (bench.py evicts L2 there); `feed`: a VolumeFeed(volume) to
  reuse its pinned buffers across calls, or a stand-in for the CUDA plumbing (tests run this loop on the gloo backend).

  Several ranks: the broadcasts are issued by a helper thread per rank on a process group of their own
  (broadcast_group(group), created collectively on first use) — `sweep` is free to run collectives on `group`
  (sharded_area_batch's all-gather, barriers), but must not use that broadcast group.  A status byte follows every
  payload: when the source rank fails to produce an image, the other ranks raise RuntimeError at that step instead of
  waiting for a broadcast that never comes.  (A failure on a NON-source rank cannot be signalled this way: the others
  find out through the communicator's own timeout.)"""
  import threading
  import torch.distributed as dist
  multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
  rank = dist.get_rank(group) if multi else src
  feed = feed if feed is not None else VolumeFeed(volume)
  is_src = rank == src
  ready = [threading.Semaphore(0), threading.Semaphore(0)]
  free = [threading.Semaphore(1), threading.Semaphore(1)]
  failure = []
  stop = threading.Event()

  def producer():
    for k in range(n_steps):
      free[k & 1].acquire()
      if stop.is_set():
        return
      try:
        feed.pack_and_stage(images[k], k & 1)
      except BaseException as e:      # surfaced by the consumer at step k
        failure.append((k, e))
        ready[k & 1].release()
        return
      ready[k & 1].release()

  def failed(k):
    for kk, e in failure:
      if kk <= k:
        return e
    return None

  if is_src and (images is None or len(images) < n_steps):
    raise ValueError("the source rank must pass one image per step")
  if multi and hasattr(feed, "staged_tensor"):
    # More than one GPU: a third thread per rank broadcasts the STAGED occupancy bytes of image k+1 on a side stream while
    # the main thread sweeps image k, so that neither the transfer nor the rendezvous of the ranks is on the step's path;
    # the step itself is then a buffer swap + the sweep, as on one GPU.  Every rank issues the broadcasts in step order,
    # on the broadcasters' own communicator.
    bgroup = broadcast_group(group)
    gsrc = dist.get_global_rank(group, src) if group is not None else src
    status = feed if hasattr(feed, "bcast_status") else _HostStatus()
    bready = [threading.Semaphore(0), threading.Semaphore(0)]
    if hasattr(status, "_st_issued"):
      status._status_bufs()
      status._st_issued = [False] * 4      # (a feed is reused across calls)

    def broadcaster():
      feed.bind_thread()
      for k in range(n_steps):
        slot = k & 1
        bad = 0
        if is_src:
          ready[slot].acquire()
          # a source rank that gives up — its producer failed, or its main thread bailed out — still sends this step's
          # broadcast, flagged, so that the other ranks raise instead of waiting for it
          bad = 1 if (failed(k) is not None or stop.is_set()) else 0
          if not bad:
            feed.wait_staged(slot)
        else:
          free[slot].acquire()
          if stop.is_set():
            return
        try:
          with feed.bcast_ctx():
            dist.broadcast(feed.staged_tensor(slot), src=gsrc, group=bgroup)     # (the payload of a failed step is whatever the buffer holds)
            status.bcast_status(k & 3, bad if is_src else None, gsrc, bgroup)
          feed.mark_staged(slot)
        except BaseException as e:
          failure.append((k, e))
          bready[slot].release()
          return
        bready[slot].release()
        # the status of this step decides whether another broadcast will come: wait for it (on this thread; the main
        # thread is sweeping the previous image meanwhile) before issuing the next one
        if bad or status.read_status(k & 3, wait=True):
          return

    threads = [threading.Thread(target=broadcaster, name="xs3d-volume-bcast")]
    if is_src:
      threads.append(threading.Thread(target=producer, name="xs3d-volume-feed"))
    for t in threads:
      t.start()
    results = []
    try:
      for k in range(n_steps):
        if before_step is not None:
          before_step(k)
        bready[k & 1].acquire()
        if failed(k) is not None:
          raise failed(k)
        if status.read_status(k & 3, wait=False):
          raise RuntimeError(f"stream_volumes: the source rank failed to produce image {k}")
        feed.commit_swap(k & 1)
        free[k & 1].release()
        r = sweep(k)
        # the sweep waited (on the device) for broadcast k, so its status byte has arrived by now
        if status.read_status(k & 3, wait=True):
          raise RuntimeError(f"stream_volumes: the source rank failed to produce image {k}")
        results.append(r)
    finally:
      stop.set()
      for sem in free + ready:
        sem.release()
      for t in threads:
        t.join()
    return results
  th = None
  if is_src:
    th = threading.Thread(target=producer, name="xs3d-volume-feed")
    th.start()
  results = []
  try:
    for k in range(n_steps):
      if before_step is not None:
        before_step(k)
      if is_src:
        ready[k & 1].acquire()
        if failed(k) is not None:
          raise failed(k)
        feed.commit(k & 1)
        free[k & 1].release()
      if multi:
        dist.broadcast(feed.cubes_tensor(), src=src, group=group)
        if not is_src:
          feed.mark_loaded()
      results.append(sweep(k))
  finally:
    if th is not None:
      stop.set()                      # a consumer that bailed out must not leave the helper thread waiting
      for sem in free:
        sem.release()
      th.join()
  return results


def stream_shared_volumes(volume, images, n_steps, sweep, group=None, before_step=None, feed=None):
  """stream_volumes for host images EVERY rank can read (one copy in shared memory mapped by all the processes of the
  node, or a replica per rank): run `sweep(k)` for k = 0 .. n_steps-1 with `volume` holding images[k].  Every rank
  bit-packs and uploads only its 1/world part of image k+1 (part_bounds) on a helper thread while step k runs; the
  helper all-gathers the parts of the bit stream on a side stream and a communicator of its own (broadcast_group) and
  builds the occupancy bytes from the whole stream behind it; the step itself is a buffer swap + the sweep.  Compared
  with stream_volumes the host work per image — reading it once, the part of the end-to-end step that is bound by host
  memory — is spread over the cores of all ranks, and 1/world of the bits crosses each GPU's PCIe link.

  Every rank must pass the same `n_steps` images (same content, shape and memory order).  `sweep` may run collectives
  on `group` but must not use broadcast_group(group).  A status byte per rank is all-gathered behind every payload: if
  any rank fails to pack its part, every rank raises at that step.  `feed`: a SharedVolumeFeed(volume), or a stand-in
  for the CUDA plumbing (the gloo tests).  One rank (or no process group): plain stream_volumes."""
  import threading
  import torch.distributed as dist
  multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
  if not multi:
    return stream_volumes(volume, images, n_steps, sweep, src=0, group=group, before_step=before_step,
                          feed=feed if feed is not None and hasattr(feed, "pack_and_stage") else None)
  if images is None or len(images) < n_steps:
    raise ValueError("every rank must pass one image per step")
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  feed = feed if feed is not None else SharedVolumeFeed(volume)
  bgroup = broadcast_group(group)
  feed.prepare_parts(world)
  voxels = int(np.prod(images[0].shape))
  part, lo, hi = part_bounds(voxels, world, rank)
  pbytes = part // 8
  free = [threading.Semaphore(1), threading.Semaphore(1)]
  bready = [threading.Semaphore(0), threading.Semaphore(0)]
  failure = []
  stop = threading.Event()
  if hasattr(feed, "_gs_issued"):
    feed._gs_issued = [False] * 4

  def failed(k):
    for kk, e in failure:
      if kk <= k:
        return e
    return None

  def helper():
    feed.bind_thread()
    for k in range(n_steps):
      slot = k & 1
      free[slot].acquire()
      bad = 1 if stop.is_set() else 0
      c_order = False
      if not bad:
        try:
          if tuple(images[k].shape) != tuple(images[0].shape):
            raise ValueError("image shape does not match the volume")
          c_order = feed.pack_part(images[k], slot, lo, hi)
        except BaseException as e:      # flagged to every rank below; raised here by the main thread at step k
          failure.append((k, e))
          bad = 1
      try:
        with feed.bcast_ctx():
          feed.wait_bits(slot)
          full = feed.bits_tensor(slot)
          dist.all_gather_into_tensor(full, full[rank * pbytes: (rank + 1) * pbytes], group=bgroup)
          feed.gather_status(k & 3, bad, rank, bgroup)
          feed.ingest(slot, c_order)
      except BaseException as e:
        failure.append((k, e))
        bready[slot].release()
        return
      bready[slot].release()
      if bad or feed.read_gathered_status(k & 3, wait=True):
        return

  th = threading.Thread(target=helper, name="xs3d-volume-parts")
  th.start()
  results = []
  try:
    for k in range(n_steps):
      if before_step is not None:
        before_step(k)
      bready[k & 1].acquire()
      if failed(k) is not None:
        raise failed(k)
      if feed.read_gathered_status(k & 3, wait=False):
        raise RuntimeError(f"stream_shared_volumes: a rank failed to produce its part of image {k}")
      feed.commit_swap(k & 1)
      free[k & 1].release()
      r = sweep(k)
      # the sweep waited (on the device) for the all-gather of step k, so its status bytes have arrived by now
      if feed.read_gathered_status(k & 3, wait=True):
        raise RuntimeError(f"stream_shared_volumes: a rank failed to produce its part of image {k}")
      results.append(r)
  finally:
    stop.set()
    for sem in free:
      sem.release()
    th.join()
  return results
