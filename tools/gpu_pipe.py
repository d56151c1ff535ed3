import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200"))
import numpy as np, torch
import xs3d
from xs3d import synthetic as syn, _abi
vol, verts, tans = syn.make_neurite(512)
pos, nrm = syn.draw_queries(verts, tans, 10000)
w = _abi.f3([32, 32, 40]); L = _abi.lib()
hvol = torch.from_numpy(np.ascontiguousarray(vol.view(np.uint8).ravel(order="F"))).pin_memory()
hpos = torch.from_numpy(pos.copy()).pin_memory(); hnrm = torch.from_numpy(nrm.copy()).pin_memory()
V2 = [xs3d.Volume(shape=vol.shape) for _ in range(2)]
outs = [(torch.empty(10000, dtype=torch.float32).pin_memory(), torch.empty(10000, dtype=torch.uint8).pin_memory()) for _ in range(2)]
log = []
def sweep(slot, i):
  t0 = time.perf_counter()
  t1 = time.perf_counter()
  _abi.check(L.xs3d_area_sweep_host(V2[slot]._h, hvol.data_ptr(), 0, 10000, hpos.data_ptr(), hnrm.data_ptr(), _abi.fptr(w), outs[slot][0].data_ptr(), outs[slot][1].data_ptr()))
  t2 = time.perf_counter()
  log.append((slot, i, t0, t1, t2))
def run(n):
  def worker(slot):
    for i in range(slot, n, 2): sweep(slot, i)
  ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
  [t.start() for t in ts]; [t.join() for t in ts]
run(4); log.clear()
torch.cuda.synchronize(); T0 = time.perf_counter(); run(8); torch.cuda.synchronize(); T1 = time.perf_counter()
print("total %.2f ms for 8 steps -> %.2f ms/step" % ((T1 - T0) * 1e3, (T1 - T0) * 1e3 / 8))
for slot, i, a, b, c in sorted(log, key=lambda x: x[2]):
  print("slot %d step %d: load %6.2f..%6.2f (%.2f)  batch ..%6.2f (%.2f)" % (slot, i, (a - T0) * 1e3, (b - T0) * 1e3, (b - a) * 1e3, (c - T0) * 1e3, (c - b) * 1e3))
# plain copy speed
d = torch.empty(vol.size, dtype=torch.uint8, device="cuda")
for _ in range(3):
  torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(hvol, non_blocking=True); torch.cuda.synchronize(); print("H2D 128 MiB: %.2f ms" % ((time.perf_counter() - t) * 1e3))
# host bit-packing rate alone (xs_pack_bits on the pinned image)
import ctypes as C
lib = C.CDLL(_abi.LIB_PATH)
lib.xs_pack_bits.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
hb = torch.empty(vol.size // 8 + 64, dtype=torch.uint8).pin_memory()
for _ in range(4):
  t = time.perf_counter(); nt = lib.xs_pack_bits(hvol.data_ptr(), 0, vol.size, hb.data_ptr()); dt = time.perf_counter() - t
  print("pack 128 MiB with %d threads: %.2f ms (%.1f GB/s)" % (nt, dt * 1e3, vol.size / dt / 1e9))
