"""Query sharding over ranks (one process per GPU) — SURVEY.md §8(e).

Cross-section queries are independent and read-only on the volume, so the multi-GPU path has no
data-path collective: every rank holds the ingested volume and works on a contiguous slice of the
query list.  torch.distributed is used only as plumbing:

  * broadcast_volume(): rank `src` ingests the image, the 2x2x2 occupancy bytes (V/8 bytes, e.g.
    16 MiB for 512^3) are broadcast once over NCCL/NVLink into every rank's resident Volume;
  * shard_bounds() / gather_results(): contiguous partition of n queries and the (small: 5 B/query)
    all-gather of areas and contact bits.

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
