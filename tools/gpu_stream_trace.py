"""Dev script: host-side timeline of the staged end-to-end loop (pack/stage on a helper thread, commit + sweep on the main one)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn, _abi, dist as xdist
vol, verts, tans = syn.make_neurite(512)
pos, nrm = syn.draw_queries(verts, tans, 10000)
w = _abi.f3([32, 32, 40]); L = _abi.lib()
V = xs3d.Volume(shape=vol.shape); V.set_stream(torch.cuda.current_stream().cuda_stream); V.load(vol)
hpos = torch.from_numpy(pos.copy()).pin_memory(); hnrm = torch.from_numpy(nrm.copy()).pin_memory()
harea = torch.empty(10000, dtype=torch.float32).pin_memory(); hct = torch.empty(10000, dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
feed = xdist.VolumeFeed(V)
log = []
T = time.perf_counter
class TracedFeed:
  def pack_and_stage(self, image, slot):
    t0 = T(); feed.pack_and_stage(image, slot); log.append(("pack+stage", slot, t0, T()))
  def commit(self, slot):
    t0 = T(); feed.commit(slot); log.append(("commit", slot, t0, T()))
  def cubes_tensor(self): return feed.cubes_tensor()
  def mark_loaded(self): pass
def sweep(k):
  t0 = T()
  _abi.check(L.xs3d_area_batch(V._h, 10000, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w), _abi.HOST, harea.data_ptr(), hct.data_ptr()))
  log.append(("sweep", k, t0, T()))
def before(k):
  t0 = T(); 
  if not os.environ.get("NO_FLUSH"): flush.fill_(k & 255)
  log.append(("flush-launch", k, t0, T()))
n = int(os.environ.get('XS_TRACE_STEPS', '12'))
xdist.stream_volumes(V, [vol] * 4, 4, sweep, before_step=before, feed=TracedFeed()); log.clear()
torch.cuda.synchronize(); t0 = T()
xdist.stream_volumes(V, [vol] * n, n, sweep, before_step=before, feed=TracedFeed())
torch.cuda.synchronize(); t1 = T()
print("total %.3f ms/step" % ((t1 - t0) * 1e3 / n))
if os.environ.get('XS_TRACE_QUIET'):
  pk = [b - a for name, k, a, b in log if name == 'pack+stage']; sw = [b - a for name, k, a, b in log if name == 'sweep']
  print("   pack+stage median %.3f ms, sweep median %.3f ms" % (np.median(pk) * 1e3, np.median(sw) * 1e3))
  sys.exit(0)
for name, k, a, b in sorted(log, key=lambda r: r[2]):
  print("%-12s %2d  %8.3f -> %8.3f  (%.3f)" % (name, k, (a - t0) * 1e3, (b - t0) * 1e3, (b - a) * 1e3))
