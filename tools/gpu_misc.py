"""Dev script: slice-batch throughput, single-call latencies, 500^3 ingest (unaligned path)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-section_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import xs3d
from xs3d import synthetic as syn
from refapi import Ref, Oracle, have_ref
ref = Ref() if have_ref() else Oracle()
vol, verts, tans = syn.make_neurite(512)
w = [32, 32, 40]
lab = syn.make_labels(vol, np.uint64, hashed=True)
qp, qn = syn.draw_queries(verts, tans, 10000, seed=2)
V = xs3d.Volume(labels=lab)
for crop in (100, 1000):
  n = 10000 if crop == 100 else 500
  V.slice_batch(qp[:n], qn[:n], w, crop=crop)
  t = time.time(); r = V.slice_batch(qp[:n], qn[:n], w, crop=crop); dt = time.time() - t
  print("slice_batch crop=%d: %d planes in %.4fs -> %.0f planes/s, shape[0]=%s" % (crop, n, dt, n / dt, r[0].shape))
t = time.time(); nref = 300
for i in range(nref): ref.projection(lab, qp[i], qn[i], w, True, 100.0)
print("reference slice crop=100: %.0f planes/s (1 core)" % (nref / (time.time() - t)))
t = time.time(); s = xs3d.slice(lab, qp[0], qn[0], w); print("single un-cropped slice", s.shape, "%.3fs (incl. 1 GiB label upload)" % (time.time() - t))
t = time.time(); s2 = ref.projection(lab, qp[0], qn[0], w, True, float("inf")); print("reference un-cropped slice %.3fs" % (time.time() - t), np.array_equal(s, s2))
V.close()
# single-call latency with a resident image
xs3d.set_shape(vol)
xs3d.cross_sectional_area(vol, qp[0], qn[0], w, use_persistent_data=True)
t = time.time()
for i in range(200): xs3d.cross_sectional_area(vol, qp[i], qn[i], w, use_persistent_data=True)
print("single-call latency (resident image): %.3f ms/query" % ((time.time() - t) / 200 * 1e3))
xs3d.clear_shape()
t = time.time(); xs3d.cross_sectional_area(vol, qp[0], qn[0], w)
print("single call, first after clear_shape (buffers re-created): %.2f ms" % ((time.time() - t) * 1e3))
t = time.time()
for i in range(20): xs3d.cross_sectional_area(vol, qp[i], qn[i], w)
print("single-call latency (upload every call, pageable numpy image): %.2f ms/query" % ((time.time() - t) / 20 * 1e3))
# unaligned ingest (500^3)
sph = syn.make_sphere()
Vs = xs3d.Volume(sph)
for _ in range(3): Vs.load(sph)
print("500^3 upload+ingest: ingest kernel %.3f ms" % Vs.timing()["ingest_ms"])
cs = np.ascontiguousarray(vol)
Vc = xs3d.Volume(cs)
Vc.load(cs); print("512^3 C-order ingest kernel %.3f ms" % Vc.timing()["ingest_ms"])
# perf.cpp workload of the reference: solid 500^3, normal (1,1,1), 20 planes
solid = np.ones((500, 500, 500), bool, order="F")
Vp = xs3d.Volume(solid)
pp = np.array([[i, 250, 250] for i in range(0, 500, 25)], np.float32); nn = np.tile(np.array([1, 1, 1], np.float32), (len(pp), 1))
Vp.cross_sectional_area_batch(pp, nn, [1, 1, 1])
t = time.time(); a = Vp.cross_sectional_area_batch(pp, nn, [1, 1, 1]); dt = time.time() - t
print("perf.cpp-like: %d planes through solid 500^3 in %.3fs -> %.1f ms/plane" % (len(pp), dt, dt / len(pp) * 1e3), Vp.stats())
t = time.time(); ra, _ = ref.area_batch(solid, pp[:4], nn[:4], [1, 1, 1]); print("reference: %.1f ms/plane" % ((time.time() - t) / 4 * 1e3), np.array_equal(ra.view(np.uint32), a[:4].view(np.uint32)))
