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
