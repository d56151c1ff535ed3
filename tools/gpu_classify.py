import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import xs3d
from xs3d import synthetic as syn, _abi
vol, verts, tans = syn.make_neurite(512)
pos, nrm = syn.draw_queries(verts, tans, 10000)
w = _abi.f3([32, 32, 40])
V = xs3d.Volume(vol)
L = _abi.lib()
L.xs3d_debug_classify.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
popc = np.zeros(10000, np.uint16)
for _ in range(2):
  _abi.check(L.xs3d_debug_classify(V._h, 10000, pos.ctypes.data, nrm.ctypes.data, _abi.fptr(w), popc.ctypes.data))
t = (C.c_float * 8)(); L.xs3d_volume_timing(V._h, t); print("classify kernel ms", t[7])
# which queries escalate? run singles through hostsim-free path: use stats of per-query batches is slow; instead use sample counts from oracle-free approach:
# run the batch in chunks of 1 query? too slow; use chunks of 50 and bisect is overkill. Use the reference fill size instead: count samples via batch stats per query (n=1 uses big tier only), so derive from the C oracle.
from refapi import Oracle
import ctypes
orc = Oracle()
# sample counts through the oracle are not exported; approximate escalation by running the GPU batch on subsets split by popcount threshold
for thr in (200, 300, 400, 500, 600, 700, 800):
  small = popc < thr
  V.cross_sectional_area_batch(pos[small], nrm[small], w)
  st = V.stats()
  V.cross_sectional_area_batch(pos[~small], nrm[~small], w)
  st2 = V.stats()
  print("thr", thr, "flagged big", int((~small).sum()), "| escalations among small:", st["escalated_mid"], "| among flagged:", st2["escalated_mid"])
print("popc percentiles", np.percentile(popc, [10, 50, 90, 95, 99, 100]))
