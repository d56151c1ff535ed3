"""Dev script (torchrun, one process per GPU): the end-to-end loop of bench.py alone, in either mode.
  MODE=source : xs3d.dist.stream_volumes        (rank 0 packs the whole image, occupancy bytes NCCL-broadcast)
  MODE=shared : xs3d.dist.stream_shared_volumes (one image in POSIX shared memory, every rank packs 1/world of it,
                                                 the bit stream is all-gathered, every GPU ingests)
XS3D_HOST_THREADS is chosen per mode unless set."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
mode = os.environ.get("MODE", "shared")
cores = os.cpu_count() or 8
if mode == "shared":
  os.environ.setdefault("XS3D_HOST_THREADS", str(max(1, min(16, cores // world - 1))))
elif rank == 0:
  os.environ.setdefault("XS3D_HOST_THREADS", str(max(4, min(16, cores - world - 1))))
import numpy as np, torch, torch.distributed as dist
import xs3d
from xs3d import dist as xdist, synthetic as syn, _abi
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
  dist.init_process_group("nccl", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
vol, verts, tans = syn.make_neurite(512)
pos, nrm = syn.draw_queries(verts, tans, 10000, seed=1 + rank)
if mode == "shared" and world > 1:
  from multiprocessing import shared_memory
  name = "xs3d_e2e_%s" % os.environ.get("MASTER_PORT", "0")
  if rank == 0:
    shm = shared_memory.SharedMemory(name=name, create=True, size=vol.size)
    np.ndarray(vol.shape, np.uint8, buffer=shm.buf, order="F")[...] = vol.view(np.uint8)
  dist.barrier()
  if rank != 0:
    shm = shared_memory.SharedMemory(name=name)
  img = np.ndarray(vol.shape, np.uint8, buffer=shm.buf, order="F").view(bool)
else:
  img = vol
w = np.asarray([32, 32, 40], np.float32)
V = xs3d.Volume(shape=vol.shape, device=local)
V.set_stream(torch.cuda.current_stream().cuda_stream)
if world > 1:
  xdist.broadcast_volume(V, vol if rank == 0 else None, src=0)
else:
  V.load(vol)
L = _abi.lib()
hpos = torch.from_numpy(pos.copy()).pin_memory(); hnrm = torch.from_numpy(nrm.copy()).pin_memory()
harea = torch.empty(10000, dtype=torch.float32).pin_memory(); hct = torch.empty(10000, dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def sweep(k):
  _abi.check(L.xs3d_area_batch(V._h, 10000, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w), _abi.HOST, harea.data_ptr(), hct.data_ptr()))
ref_a = None
def barrier():
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier(); torch.cuda.synchronize()
sweep(0); ref_a = harea.numpy().copy(); ref_c = hct.numpy().copy()
TRACE = os.environ.get("XS_TRACE")
tlog = {"pack": [], "sweep": [], "gap": [], "bcast_enq": [], "status_wait": []}
_last_end = [None]
class TracedFeed(xdist.VolumeFeed):
  """host-side timeline (XS_TRACE=1): how long the packing, the enqueue of a broadcast, the wait for a broadcast's
  status and the sweeps take, and how long the main thread sits between two sweeps"""
  def pack_and_stage(self, image, slot):
    t0 = time.perf_counter(); super().pack_and_stage(image, slot); tlog["pack"].append(time.perf_counter() - t0)
  def bcast_status(self, slot, value, src, group):
    t0 = time.perf_counter(); super().bcast_status(slot, value, src, group); tlog["bcast_enq"].append(time.perf_counter() - t0)
  def read_status(self, slot, wait=True):
    t0 = time.perf_counter(); r = super().read_status(slot, wait)
    if wait: tlog["status_wait"].append(time.perf_counter() - t0)
    return r
feed = xdist.SharedVolumeFeed(V) if mode == "shared" else (TracedFeed(V) if TRACE else xdist.VolumeFeed(V))
if TRACE:
  _sweep = sweep
  def sweep(k):
    t0 = time.perf_counter()
    if _last_end[0] is not None: tlog["gap"].append(t0 - _last_end[0])
    _sweep(k)
    _last_end[0] = time.perf_counter(); tlog["sweep"].append(_last_end[0] - t0)
HOOK = os.environ.get("HOOK")
def run(n):
  if HOOK and mode != "shared":
    # the L2-evicting write of step k+1 enqueued by the library's after-enqueue hook during step k's call (as bench.py does)
    done = [0]
    def nxt():
      done[0] += 1
      if done[0] < n: flush.fill_(done[0] & 255)
    V.set_after_enqueue(nxt)
    try:
      xdist.stream_volumes(V, [img] * n if rank == 0 else None, n, sweep, before_step=lambda k: flush.fill_(0) if k == 0 else None, feed=feed)
    finally:
      V.set_after_enqueue(None)
    return
  if mode == "shared":
    xdist.stream_shared_volumes(V, [img] * n, n, sweep, before_step=lambda k: flush.fill_(k & 255), feed=feed)
  else:
    xdist.stream_volumes(V, [img] * n if rank == 0 else None, n, sweep, before_step=lambda k: flush.fill_(k & 255), feed=feed)
steps = int(os.environ.get("STEPS", "40"))
run(6)
res = []
for rep in range(int(os.environ.get("REPS", "3"))):
  barrier()
  s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
  s.record(); run(steps); torch.cuda.synchronize(); e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e)
  barrier()
  if world > 1:
    tt = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
  res.append(ms / steps)
ok = np.array_equal(harea.numpy().view(np.uint32), ref_a.view(np.uint32)) and np.array_equal(hct.numpy(), ref_c)
oks = torch.tensor([int(ok)], device=dev)
if world > 1:
  dist.all_reduce(oks, op=dist.ReduceOp.MIN)
if TRACE:
  tm = V.timing()
  for r in range(world):
    if world > 1: dist.barrier()
    if r == rank:
      print("rank %d host timeline (median / p90 ms): " % rank + " | ".join("%s %.3f / %.3f (n=%d)" % (k, np.median(v) * 1e3, np.percentile(v, 90) * 1e3, len(v)) for k, v in tlog.items() if v) + " || last sweep on the GPU: " + " ".join("%s %.3f" % (k.replace("_ms", ""), tm[k]) for k in ("prepass_ms", "warp_tier_ms", "mid_tier_ms", "mid_exposed_ms", "polygon_ms", "combine_ms", "p2_ms") if k in tm), flush=True)
if rank == 0:
  print("mode %-6s N=%d threads/rank %s: e2e ms/step %s -> %.1f M xs/s ; results identical on every rank: %s" % (
      mode, world, os.environ.get("XS3D_HOST_THREADS"), " ".join("%.3f" % r for r in res), world * 10000 / min(res) / 1e3, bool(int(oks[0]))), flush=True)
if mode == "shared" and world > 1:
  del img
  barrier()
  shm.close()
  if rank == 0:
    shm.unlink()
if world > 1:
  dist.destroy_process_group()
